#!/bin/bash
# round 2, GPU call B: pipelined tile kernel (v2): parity, occupancy variants, ncu of the tile kernel
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_assembly.py tests/test_gpu_unstructured.py -x -q -m gpu > gpurun_out/b_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/b_pytest.log
tail -4 gpurun_out/b_pytest.log
run() {  # tag, lib, args...
  tag=$1; lib=$2; shift 2
  PHIFEM_B200_LIB=$PWD/phifem_b200/$lib timeout 600 python bench.py --no-cpu --no-e2e --steps 10 "$@" > gpurun_out/b_bench_$tag.json 2> gpurun_out/b_bench_$tag.err
  echo "$tag ($lib $*) -> exit $?"
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/b_bench_$tag.json"))
    k = d["roofline"]["kernels_ms"]
    print("  step %.3f cells %.3f surface %.3f symbolic %.0f recompute %s" % (d["ms_per_step"], k["assemble_cells"], k["assemble_surface"], d["symbolic_ms"], d["scatter"].get("recompute_factor")))
except Exception as e:
    print("  no line:", e)
PY
}
run t128 libphifem_b200.so --cell-pass tiles --rows-per-tile 128
run t256 libphifem_b200.so --cell-pass tiles --rows-per-tile 256
run t128_mb5 libphifem_b200_t5.so --cell-pass tiles --rows-per-tile 128
run t128_mb3 libphifem_b200_t3.so --cell-pass tiles --rows-per-tile 128
run t256_mb1 libphifem_b200_t3.so --cell-pass tiles --rows-per-tile 256
run u128 libphifem_b200.so --mesh unstructured --cell-pass tiles --rows-per-tile 128
CMD="python bench.py --no-cpu --no-e2e --steps 2 --warmup 3 --cell-pass tiles --rows-per-tile 128"
$CMD > gpurun_out/b_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_assemble_tiles -s 4 -c 2 -o gpurun_out/b_tiles128 $CMD > gpurun_out/b_ncu.log 2>&1
echo "ncu exit $?"
