#!/bin/bash
# round 2 (session 4), GPU call S8 (8 GPUs): the multi-GPU line of the final build
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
SECONDS=0
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r4s_scale8.json 2> gpurun_out/r4s_scale8.err; echo "bench exit $? wall ${SECONDS}s"
tail -2 gpurun_out/r4s_scale8.err
python -c "
import json
d=json.load(open('gpurun_out/r4s_scale8.json')); e=d['e2e']; s=d.get('strong') or {}
print('step', d['ms_per_step'], 'value', d['value'], 'parity', d.get('parity_ok'), 'strong', s.get('ms_per_step'))
print('e2e', e['ms_per_step'], e.get('one_step_at_a_time_ms'), e.get('two_steps_in_flight_ms'), e['value'])"
