#!/bin/bash
# round 2, GPU call G (N GPUs): bench.py --gpus N only (weak + strong + parity)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-8}
EXTRA="${2:-}"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 \
  bench.py --gpus $N --steps 20 --warmup 3 $EXTRA > gpurun_out/g_bench_$N.json 2> gpurun_out/g_bench_$N.err
echo "bench exit $?"; grep -v "Warning\|_warn_once\|^$" gpurun_out/g_bench_$N.err | tail -5
python - <<PY
import json
d = json.load(open("gpurun_out/g_bench_$N.json"))
print("weak", d["ms_per_step"], d["value"], "parity_ok", d.get("parity_ok"))
s = d.get("strong")
print("strong", {k: s[k] for k in ("ms_per_step", "value", "cuda_graph", "scatter_s", "local_cells", "owned_cells", "symbolic_ms_max", "topology_ms_max", "tags_ms", "assembly_ms")} if s else None)
PY
