#!/bin/bash
# round 2 (session 4), GPU call S4 (4 GPUs): the multi-GPU line of the final build (pipelined end-to-end leg at 4 ranks)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
SECONDS=0
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 4 --steps 10 --warmup 3 > gpurun_out/r4s_scale4.json 2> gpurun_out/r4s_scale4.err; echo "bench exit $? wall ${SECONDS}s"
tail -2 gpurun_out/r4s_scale4.err
python -c "
import json
d=json.load(open('gpurun_out/r4s_scale4.json')); e=d['e2e']; s=d.get('strong') or {}
print('step', d['ms_per_step'], 'value', d['value'], 'parity', d.get('parity_ok'), 'strong', s.get('ms_per_step'))
print('e2e', e['ms_per_step'], e.get('one_step_at_a_time_ms'), e.get('two_steps_in_flight_ms'), e['value'])"
