#!/bin/bash
# round 2 (session 3), GPU call K: row-gather cell pass with padded coordinates (one 256-bit gather per vertex)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for v in 1 0; do
  PHIFEM_ROWS_X4=$v python bench.py --no-cpu --no-e2e --no-solve --no-replan --steps 20 > gpurun_out/r3k_bench_x4_$v.json 2> gpurun_out/r3k_bench_x4_$v.err
  python -c "
import json; d=json.load(open('gpurun_out/r3k_bench_x4_$v.json')); k=d['roofline']['kernels_ms']; print('x4=$v', round(d['ms_per_step'],4), {n: round(t,4) for n,t in k.items()}); u=d['unstructured']; print('  unstructured', round(u['ms_per_step'],4), {n: round(t,4) for n,t in u['kernels_ms'].items()})" || tail -3 gpurun_out/r3k_bench_x4_$v.err
done
timeout 900 python -m pytest tests/test_gpu_assembly.py tests/test_gpu_unstructured.py tests/test_gpu_full_size.py -x -q -m gpu 2>&1 | tail -3
