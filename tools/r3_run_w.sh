#!/bin/bash
# round 2 (session 3), 4-GPU call W: halo exchange over peer memory at 4 ranks (parity against the single-GPU operator),
# 2-GPU pytest file, bench line at 4 GPUs
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29681 tests/dist_gpu_worker.py 10 exchange-peer 2>&1 | grep -E "DIST-OK|Error|error|assert" | head -5
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29683 tests/dist_gpu_worker.py 10 rows 2>&1 | grep -E "DIST-OK|Error|error|assert" | head -5
timeout 900 python -m pytest tests/test_gpu_dist.py -x -q -m gpu 2>&1 | tail -3
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29685 bench.py --gpus 4 --steps 20 --warmup 3 > gpurun_out/r3w_scale_4.json 2> gpurun_out/r3w_scale_4.err
python -c "
import json; d=json.load(open('gpurun_out/r3w_scale_4.json')); print(d['ms_per_step'], d['value'], d.get('parity_ok'), 'e2e', d['e2e']['ms_per_step'], d['e2e']['value']); s=d['strong']; print('strong', s['ms_per_step'], s['value'])" || tail -20 gpurun_out/r3w_scale_4.err
