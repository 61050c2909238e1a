#!/bin/bash
# round 2 (session 3), GPU call P: pencil renumbering against the Morton curve -- SURVEY 8(d) variant and a Delaunay mesh
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for c in pencil morton; do
  python bench.py --no-cpu --no-e2e --no-solve --no-replan --steps 20 --curve $c > gpurun_out/r3p_bench_$c.json 2> gpurun_out/r3p_bench_$c.err
  python -c "
import json; d=json.load(open('gpurun_out/r3p_bench_$c.json')); u=d['unstructured']; print('$c', 'structured', round(d['ms_per_step'],4), 'unstructured', round(u['ms_per_step'],4), {n: round(t,4) for n,t in u['kernels_ms'].items()}, 'reorder_ms', round(u['reorder_ms'],1), 'symbolic', round(u['symbolic_ms'],1))" || tail -5 gpurun_out/r3p_bench_$c.err
done
NPTS=400000 timeout 900 python tools/r3_delaunay.py 2>&1 | tail -5
