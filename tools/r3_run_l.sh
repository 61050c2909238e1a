#!/bin/bash
# round 2 (session 3), GPU call L: facet-once kernel after the cell pass (records still in L2 for the surface rows) against the fork
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for v in fork after fork after; do
  PHIFEM_ONCE_ORDER=$v python bench.py --no-cpu --no-e2e --no-unstructured --no-solve --no-replan --steps 30 > gpurun_out/r3l_bench_$v.json 2> gpurun_out/r3l_bench_$v.err
  python -c "
import json; d=json.load(open('gpurun_out/r3l_bench_$v.json')); k=d['roofline']['kernels_ms']; print('$v', round(d['ms_per_step'],4), {n: round(t,4) for n,t in k.items()})" || tail -3 gpurun_out/r3l_bench_$v.err
done
