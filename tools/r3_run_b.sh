#!/bin/bash
# round 2 (session 3), GPU call B: where the staged facet kernel differs; staged (strided mapping) cell kernel; PCIe rates
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=64 python tools/r3_debug_facets.py 2>&1 | tail -12
python tools/r3_debug_facets.py 2>&1 | tail -12
timeout 600 python -m pytest tests/test_gpu_tags.py tests/test_gpu_edge_cases.py tests/test_gpu_full_size.py -x -q -m gpu 2>&1 | tail -5
for v in staged ldg; do
  if [ $v = ldg ]; then export PHIFEM_CELLS_KERNEL=ldg PHIFEM_FACETS_KERNEL=ldg; fi
  python bench.py --no-cpu --no-e2e --no-unstructured --no-solve --no-replan --steps 20 > gpurun_out/r3b_bench_$v.json 2> gpurun_out/r3b_bench_$v.err
  python -c "
import json; d=json.load(open('gpurun_out/r3b_bench_$v.json')); print('$v', d['ms_per_step'], d['roofline']['kernels_ms'])"
done
python tools/r3_pcie.py
