#!/bin/bash
# usage: tools/sweep_env.sh VAR v1 v2 ... -- prints ms_per_step and the per-kernel split of bench.py for each value
var=$1; shift
for v in "$@"; do
  env $var=$v python bench.py --steps 10 --no-cpu --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); k=d['roofline']['kernels_ms']
print('$var=$v', round(d['ms_per_step'],4), {n: round(t,4) for n,t in k.items()})"
done
