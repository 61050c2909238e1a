#!/bin/bash
# round 2 (session 4), GPU call X: phifem_tags_match (AssemblyPlan.matches on the device), default bench line with plan_check_ms
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_edge_cases.py tests/test_gpu_rows_plan_capi.py tests/test_gpu_assembly.py -x -q -m gpu > gpurun_out/r4x_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/r4x_pytest.log
tail -4 gpurun_out/r4x_pytest.log
python bench.py > gpurun_out/r4x_bench_default.json 2> gpurun_out/r4x_bench_default.err; echo "exit $?"; tail -2 gpurun_out/r4x_bench_default.err
python -c "
import json; d=json.load(open('gpurun_out/r4x_bench_default.json')); print(d['ms_per_step'], d['e2e']['ms_per_step'], 'check', d['plan_check_ms'], 'sym', d['symbolic_ms'], d['cold_step_ms'])"
