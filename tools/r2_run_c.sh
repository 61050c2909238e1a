#!/bin/bash
# round 2, GPU call C: the restructured bench.py (default line with the unstructured / time-to-solution keys,
# same-config CPU arm)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( time python bench.py > gpurun_out/c_bench.json 2> gpurun_out/c_bench.err ) 2> gpurun_out/c_time.txt
echo "bench exit $?"; tail -3 gpurun_out/c_time.txt; tail -5 gpurun_out/c_bench.err
python - <<PY
import json
d = json.load(open("gpurun_out/c_bench.json"))
print("step", d["ms_per_step"], d["roofline"]["kernels_ms"])
print("frac", d["roofline"]["frac"], "step_frac", d["roofline"]["step_frac"], d["roofline"]["step_frac_survey_int32_tags"])
print("unstructured", {k: d["unstructured"].get(k) for k in ("ms_per_step", "ratio_to_structured", "reorder_ms", "symbolic_ms", "error")})
print("tts", d.get("time_to_solution_ms"))
print("cpu", d["cpu_baseline"])
print("e2e", d["e2e"]["ms_per_step"], "launches", d["gpu_launches_per_step"], d["gpu_kernels"], "symbolic", d["symbolic_ms"], d["topology_s"])
PY
( time python bench.py --impl reference > gpurun_out/c_ref.json 2> gpurun_out/c_ref.err ) 2> gpurun_out/c_ref_time.txt
echo "ref exit $?"; tail -3 gpurun_out/c_ref_time.txt; cat gpurun_out/c_ref.json | cut -c1-600
