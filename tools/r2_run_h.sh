#!/bin/bash
# round 2, GPU call H: facet kernel grid (one tile per CTA beside the boundary kernel) against the persistent grid
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for g in full persistent; do
  PHIFEM_FACETS_GRID=$g python bench.py --no-cpu --no-e2e --no-unstructured --no-solve --no-replan --steps 20 > gpurun_out/h_bench_$g.json 2> gpurun_out/h_bench_$g.err
  python -c "
import json; d=json.load(open('gpurun_out/h_bench_$g.json')); print('$g', d['ms_per_step'], d['roofline']['kernels_ms'])"
done
timeout 600 python -m pytest tests/test_gpu_tags.py tests/test_gpu_edge_cases.py tests/test_gpu_full_size.py -x -q -m gpu 2>&1 | tail -3
