#!/bin/bash
# round 2 (session 4), GPU call Z: final evidence of the round -- smoke, the whole GPU suite, launch list and --set full of
# the dominant kernel of the final build, default bench line, reference arm
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r4z_smoke.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/r4z_smoke.log
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r4z_pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/r4z_pytest_gpu.log
tail -4 gpurun_out/r4z_pytest_gpu.log
CMD="python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-unstructured --no-solve --no-replan --no-others"
$CMD > gpurun_out/r4z_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -c 400 --csv --log-file gpurun_out/r4z_launches.csv $CMD > gpurun_out/r4z_ncu1.log 2>&1
echo "launch list exit $?"
$CMD > gpurun_out/r4z_plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:^k_assemble_rows_p1 -s 8 -c 2 -o gpurun_out/r4z_step_rows -f $CMD > gpurun_out/r4z_ncu_rows.log 2>&1
echo "rows full exit $?"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r4z_bench_reference.json 2> gpurun_out/r4z_bench_reference.err; echo "reference exit $?"
python bench.py > gpurun_out/r4z_bench_default.json 2> gpurun_out/r4z_bench_default.err; echo "bench exit $?"
python -c "
import json; d=json.load(open('gpurun_out/r4z_bench_default.json')); u=d['unstructured']; e=d['e2e']
print(d['ms_per_step'], d['value'], 'e2e', e['ms_per_step'], e.get('one_step_at_a_time_ms'), 'sym', d['symbolic_ms'], 'unstructured', u['ms_per_step'], d['roofline']['frac'], d['roofline']['step_frac'], 'clocks', d['clocks'])
r=json.load(open('gpurun_out/r4z_bench_reference.json')); print('reference', r['value'], r['cpu_baseline']['cores'])"
