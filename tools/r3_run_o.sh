#!/bin/bash
# round 2 (session 3), GPU call O: the whole GPU suite, smoke(), the reference arm and the default bench line
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 2400 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -6 | tee gpurun_out/r3o_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r3o_bench_reference.json 2> gpurun_out/r3o_bench_reference.err; tail -c 400 gpurun_out/r3o_bench_reference.json; echo
python bench.py > gpurun_out/r3o_bench_default.json 2> gpurun_out/r3o_bench_default.err
python -c "
import json; d=json.load(open('gpurun_out/r3o_bench_default.json')); print(d['ms_per_step'], d['value'], 'e2e', d['e2e']['ms_per_step'], 'sym', d['symbolic_ms'], d['symbolic_first_call_ms'], 'launches', d['gpu_launches_per_step'], d['gpu_kernels'])"
