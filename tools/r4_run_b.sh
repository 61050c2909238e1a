#!/bin/bash
# round 2 (session 4), GPU call B: counter block posted to pinned host memory by a kernel (phifem_post_to_host);
# end-to-end leg with two steps in flight
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tags.py tests/test_gpu_edge_cases.py tests/test_reference_tagging_golden.py -x -q -m gpu > gpurun_out/r4b_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/r4b_pytest.log
tail -4 gpurun_out/r4b_pytest.log
for i in 1 2; do
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu --no-unstructured --no-solve --no-replan > gpurun_out/r4b_bench$i.json 2> gpurun_out/r4b_bench$i.err; echo "bench exit $?"
tail -3 gpurun_out/r4b_bench$i.err
python -c "
import json
d=json.load(open('gpurun_out/r4b_bench$i.json')); e=d['e2e']; print('step', d['ms_per_step'], 'e2e', e['ms_per_step'], 'one', e.get('one_step_at_a_time_ms'), 'two', e.get('two_steps_in_flight_ms'))"
done
