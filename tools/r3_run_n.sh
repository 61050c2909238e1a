#!/bin/bash
# round 2 (session 3), 8-GPU call N: weak + strong scaling and parity with the staged tag kernels / boundary records
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29671 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/r3n_scale_8.json 2> gpurun_out/r3n_scale_8.err
python -c "
import json; d=json.load(open('gpurun_out/r3n_scale_8.json')); print(d['ms_per_step'], d['value'], d.get('parity_ok')); print(json.dumps(d.get('strong'))[:1500])" || tail -20 gpurun_out/r3n_scale_8.err
