#!/bin/bash
# round 2 (session 4), GPU call E: default bench line with the other configurations and the pipelined end-to-end leg
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
SECONDS=0; python bench.py > gpurun_out/r4e_bench_default.json 2> gpurun_out/r4e_bench_default.err; echo "bench exit $?"
echo "bench wall ${SECONDS}s"; tail -3 gpurun_out/r4e_bench_default.err
python -c "
import json
d=json.load(open('gpurun_out/r4e_bench_default.json')); e=d['e2e']
print('step', d['ms_per_step'], 'e2e', e['ms_per_step'], e.get('one_step_at_a_time_ms'), e.get('two_steps_in_flight_ms'), 'unstructured', d['unstructured']['ms_per_step'], 'tts', d['time_to_solution_ms'].get('total_ms'))
for k, v in d['other_configs'].items(): print(k, {a: v.get(a) for a in ('ms_per_step', 'value', 'step_frac', 'error')}, v.get('kernels_ms'))"
