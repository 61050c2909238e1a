#!/bin/bash
# round 2 (session 3), 2-GPU call G: halo exchange over NVLink peer memory (phifem_halo_exchange) -- parity and timing
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dist.py -x -q -m gpu 2>&1 | tail -5
for v in peer nccl rows; do
  extra="--dist-mode exchange"
  [ $v = nccl ] && extra="--dist-mode exchange --no-peer"
  [ $v = rows ] && extra=""
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29655 bench.py --gpus 2 --steps 20 --warmup 3 --no-strong --no-parity --no-cpu $extra > gpurun_out/r3g_scale2_$v.json 2> gpurun_out/r3g_scale2_$v.err
  python -c "
import json; d=json.load(open('gpurun_out/r3g_scale2_$v.json')); print('$v', d['ms_per_step'], d['value'], d['roofline']['kernels_ms']); print(d['config']['partition'])" || tail -5 gpurun_out/r3g_scale2_$v.err
done
