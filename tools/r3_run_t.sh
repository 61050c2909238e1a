#!/bin/bash
# round 2 (session 3), GPU call T: P2 cell kernel with plain stores for the single-contributor entries
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_assembly_pk.py -x -q -m gpu 2>&1 | tail -3
for v in default nos sb64 sb64pf; do
  lib=$PWD/phifem_b200/libphifem_b200_$v.so
  [ $v = default ] && lib=$PWD/phifem_b200/libphifem_b200.so
  PHIFEM_B200_LIB=$lib python bench.py --config 2d-p2 --no-cpu --no-e2e --steps 10 > gpurun_out/r3t_bench_$v.json 2> gpurun_out/r3t_bench_$v.err
  python -c "
import json; d=json.load(open('gpurun_out/r3t_bench_$v.json')); k=d['roofline']['kernels_ms']; print('$v', round(d['ms_per_step'],4), {n: round(t,4) for n,t in k.items()})" || tail -3 gpurun_out/r3t_bench_$v.err
done
