#!/bin/bash
# round 2 (session 3), GPU call R: the other configurations with the current build
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for c in 2d-p1 2d-p2 3d-p2; do
  python bench.py --config $c --no-cpu --steps 20 > gpurun_out/r3r_bench_$c.json 2> gpurun_out/r3r_bench_$c.err
  python -c "
import json; d=json.load(open('gpurun_out/r3r_bench_$c.json')); k=d['roofline']['kernels_ms']; print('$c', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['ms_per_step'],3) if d.get('e2e') else None, {n: round(t,4) for n,t in k.items()})" || tail -3 gpurun_out/r3r_bench_$c.err
done
python tools/bench_operators.py --ops neumann > gpurun_out/r3r_operators_neumann.jsonl 2> gpurun_out/r3r_neumann.err; python -c "
import json
for l in open('gpurun_out/r3r_operators_neumann.jsonl'):
    d=json.loads(l); print(d['operator'], round(d['ms_per_step'],3), round(d['ms_tags'],3), round(d['ms_assembly'],3))"
