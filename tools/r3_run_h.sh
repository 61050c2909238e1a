#!/bin/bash
# round 2 (session 3), GPU call H: surface row pass with three work buffers / 5 CTAs per SM
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for v in default sd3 sm5 sd3b; do
  lib=$PWD/phifem_b200/libphifem_b200_$v.so
  [ $v = default ] && lib=$PWD/phifem_b200/libphifem_b200.so
  PHIFEM_B200_LIB=$lib python bench.py --no-cpu --no-e2e --no-unstructured --no-solve --no-replan --steps 20 > gpurun_out/r3h_bench_$v.json 2> gpurun_out/r3h_bench_$v.err
  python -c "
import json; d=json.load(open('gpurun_out/r3h_bench_$v.json')); k=d['roofline']['kernels_ms']; print('$v', round(d['ms_per_step'],4), {n: round(t,4) for n,t in k.items()})" || tail -3 gpurun_out/r3h_bench_$v.err
done
