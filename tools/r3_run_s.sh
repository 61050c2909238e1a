#!/bin/bash
# round 2 (session 3), GPU call S: P2 cell kernel (config C) -- occupancy / block size / slot prefetch variants
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for v in default pf m4 m2 b64 pfm2; do
  lib=$PWD/phifem_b200/libphifem_b200_$v.so
  [ $v = default ] && lib=$PWD/phifem_b200/libphifem_b200.so
  PHIFEM_B200_LIB=$lib python bench.py --config 2d-p2 --no-cpu --no-e2e --steps 10 > gpurun_out/r3s_bench_$v.json 2> gpurun_out/r3s_bench_$v.err
  python -c "
import json; d=json.load(open('gpurun_out/r3s_bench_$v.json')); k=d['roofline']['kernels_ms']; print('$v', round(d['ms_per_step'],4), {n: round(t,4) for n,t in k.items()})" || tail -3 gpurun_out/r3s_bench_$v.err
done
