#!/bin/bash
# round 2, GPU call E (2 GPUs): sharded tests, then bench.py --gpus 2 (parity_ok, strong scaling of config E, weak slabs)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-2}
timeout 900 python -m pytest tests/test_gpu_dist.py -x -q -m gpu > gpurun_out/e_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/e_pytest.log
tail -4 gpurun_out/e_pytest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/e_bench_$N.json 2> gpurun_out/e_bench_$N.err
echo "bench exit $?"; tail -5 gpurun_out/e_bench_$N.err
python - <<PY
import json
d = json.load(open("gpurun_out/e_bench_$N.json"))
print("weak", d["ms_per_step"], d["value"], "parity_ok", d.get("parity_ok"))
s = d.get("strong")
print("strong", {k: s[k] for k in ("ms_per_step", "value", "cuda_graph", "scatter_s", "local_cells", "owned_cells", "symbolic_ms_max", "topology_ms_max", "tags_ms", "assembly_ms")} if s else None)
PY
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
  bench.py --gpus $N --steps 20 --warmup 3 --no-peer --no-parity > gpurun_out/e_bench_${N}_eager.json 2> gpurun_out/e_bench_${N}_eager.err
python - <<PY
import json
d = json.load(open("gpurun_out/e_bench_${N}_eager.json"))
s = d.get("strong")
print("strong nccl all-reduce", {k: s[k] for k in ("ms_per_step", "value", "cuda_graph")} if s else None)
PY
