#!/bin/bash
# round 2 (session 4), GPU call S4b (4 GPUs): host topology of a multi-GPU box; the multi-GPU line with every rank bound to
# the host cores local to its GPU
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( nvidia-smi topo -m | head -14; for n in /sys/devices/system/node/node*; do echo "$n $(cat $n/cpulist)"; done; nproc ) > gpurun_out/r4s4b_topology.txt 2>&1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29529 bench.py --gpus 4 --steps 10 --warmup 3 --no-strong --no-parity > gpurun_out/r4s4b_scale4.json 2> gpurun_out/r4s4b_scale4.err; echo "bench exit $?"
grep "bound to\|binding" gpurun_out/r4s4b_scale4.err | head -4
cat gpurun_out/r4s4b_topology.txt | head -12
python -c "
import json
d=json.load(open('gpurun_out/r4s4b_scale4.json')); e=d['e2e']
print('step', d['ms_per_step'], 'e2e', e['ms_per_step'], e.get('one_step_at_a_time_ms'), e.get('two_steps_in_flight_ms'))"
