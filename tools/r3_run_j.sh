#!/bin/bash
# round 2 (session 3), GPU call J: entity-record kernel with one atomic per warp (e2e), 5 CTAs per SM for the staged cell classifier
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tags.py tests/test_gpu_edge_cases.py tests/test_gpu_full_size.py tests/test_gpu_unstructured.py -x -q -m gpu 2>&1 | tail -3
python tools/r3_tagbench.py 2>&1 | tail -1
PHIFEM_B200_LIB=$PWD/phifem_b200/libphifem_b200_c5.so python tools/r3_tagbench.py 2>&1 | tail -1
python bench.py --no-cpu --no-unstructured --no-solve --no-replan --steps 20 > gpurun_out/r3j_bench.json 2> gpurun_out/r3j_bench.err
python -c "
import json; d=json.load(open('gpurun_out/r3j_bench.json')); print(d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], d['roofline']['kernels_ms'])"
