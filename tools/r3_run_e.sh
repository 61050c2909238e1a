#!/bin/bash
# round 2 (session 3), GPU call E: elasticity Dirichlet passes (list-driven default against the full-matrix pass)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_elasticity.py -x -q -m gpu 2>&1 | tail -4
python tools/bench_operators.py --ops elasticity-2d,elasticity-3d > gpurun_out/r3e_operators_list.jsonl 2> gpurun_out/r3e_list.err
python tools/bench_operators.py --ops elasticity-2d --full-bc > gpurun_out/r3e_operators_full.jsonl 2> gpurun_out/r3e_full.err
python - <<'PY'
import json
for f in ("gpurun_out/r3e_operators_list.jsonl", "gpurun_out/r3e_operators_full.jsonl"):
    for l in open(f):
        d = json.loads(l)
        print(d["operator"], d["dirichlet_pass"], "ms", round(d["ms_per_step"], 3), "tags", round(d["ms_tags"], 3), "asm", round(d["ms_assembly"], 3))
PY
