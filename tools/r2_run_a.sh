#!/bin/bash
# round 2, GPU call A: tiles parity + first numbers (structured / unstructured, rows / tiles)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv > gpurun_out/a_gpu.txt
timeout 900 python -m pytest tests/test_gpu_assembly.py tests/test_gpu_unstructured.py tests/test_gpu_edge_cases.py -x -q -m gpu > gpurun_out/a_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/a_pytest.log
tail -5 gpurun_out/a_pytest.log
for cfg in "--cell-pass rows" "--cell-pass tiles --rows-per-tile 256" "--cell-pass tiles --rows-per-tile 128" \
           "--mesh unstructured --cell-pass rows" "--mesh unstructured --cell-pass tiles --rows-per-tile 256" \
           "--mesh unstructured --no-reorder --cell-pass rows"; do
  tag=$(echo "$cfg" | tr -d ' -' )
  timeout 600 python bench.py --no-cpu --no-e2e --steps 10 $cfg > gpurun_out/a_bench_$tag.json 2> gpurun_out/a_bench_$tag.err
  echo "$cfg -> exit $?"
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/a_bench_$tag.json"))
    print(d["ms_per_step"], d["roofline"]["kernels_ms"], d["symbolic_ms"], d.get("reorder_ms"), d["scatter"].get("recompute_factor"))
except Exception as e:
    print("no line:", e)
PY
done
