#!/bin/bash
# round 2 (session 3), GPU call D: staged facet kernel after the fix, boundary facets from per-mesh records
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python tools/r3_debug_facets.py 2>&1 | tail -8
python tools/r3_tagbench.py 2>&1 | tail -2
timeout 900 python -m pytest tests/test_gpu_tags.py tests/test_gpu_edge_cases.py tests/test_gpu_full_size.py tests/test_gpu_unstructured.py tests/test_gpu_reference_golden.py -x -q -m gpu 2>&1 | tail -5
python bench.py --no-cpu --no-e2e --no-unstructured --no-solve --no-replan --steps 20 > gpurun_out/r3d_bench.json 2> gpurun_out/r3d_bench.err
python -c "
import json; d=json.load(open('gpurun_out/r3d_bench.json')); print(d['ms_per_step'], d['roofline']['kernels_ms'])"
