#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_elasticity.py -x -q -m gpu -k large_pattern 2>&1 | tail -30
timeout 1200 python -m pytest tests/test_gpu_tags.py tests/test_gpu_edge_cases.py tests/test_gpu_full_size.py tests/test_gpu_unstructured.py -x -q -m gpu 2>&1 | tail -5
