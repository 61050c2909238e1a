#!/bin/bash
# round 2 (session 4), GPU call F (2 GPUs): multi-GPU bench line with the pipelined end-to-end leg, 2-GPU tests
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
SECONDS=0
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r4f_scale2.json 2> gpurun_out/r4f_scale2.err; echo "bench exit $? wall ${SECONDS}s"
tail -3 gpurun_out/r4f_scale2.err
python -c "
import json
d=json.load(open('gpurun_out/r4f_scale2.json')); e=d['e2e']
print('step', d['ms_per_step'], 'value', d['value'], 'parity', d.get('parity_ok'), 'strong', d.get('strong'))
print('e2e', e['ms_per_step'], e.get('one_step_at_a_time_ms'), e.get('two_steps_in_flight_ms'), e['value'])"
timeout 900 python -m pytest tests/test_gpu_dist.py -x -q -m gpu > gpurun_out/r4f_pytest_dist.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/r4f_pytest_dist.log
