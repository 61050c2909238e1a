#!/bin/bash
# round 2 (session 3), GPU call U: P2 configurations with 64-thread cell blocks + slot prefetch; ncu --set full of the P2 cell kernel
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_assembly_pk.py tests/test_gpu_convergence.py -x -q -m gpu 2>&1 | tail -3
for c in 2d-p2 3d-p2; do
  python bench.py --config $c --no-cpu --no-e2e --steps 10 > gpurun_out/r3u_bench_$c.json 2> gpurun_out/r3u_bench_$c.err
  python -c "
import json; d=json.load(open('gpurun_out/r3u_bench_$c.json')); k=d['roofline']['kernels_ms']; print('$c', round(d['ms_per_step'],4), {n: round(t,4) for n,t in k.items()})" || tail -3 gpurun_out/r3u_bench_$c.err
done
CMD="python bench.py --config 2d-p2 --steps 2 --warmup 3 --no-cpu --no-e2e"
ncu --set full --clock-control none --import-source on -k regex:^k_assemble_cells_pk -s 4 -c 1 -o gpurun_out/u_p2_cells -f $CMD > gpurun_out/u_ncu.log 2>&1
echo "ncu exit $?"
