#!/bin/bash
# round 2 (session 3), GPU call Y: P2 cell kernel with its quadrature loop unrolled by two
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for v in default u2m6 u2m4 u1m4; do
  lib=$PWD/phifem_b200/libphifem_b200_$v.so
  [ $v = default ] && lib=$PWD/phifem_b200/libphifem_b200.so
  PHIFEM_B200_LIB=$lib python bench.py --config 2d-p2 --no-cpu --no-e2e --steps 10 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); k=d['roofline']['kernels_ms']; print('$v', round(d['ms_per_step'],4), round(k['assemble_cells'],4))"
done
