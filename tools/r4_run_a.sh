#!/bin/bash
# round 2 (session 4), GPU call A: push form of the cell-once pass (k_assemble_push_p1) -- parity, then config E timings:
# R = 256 / 128, neighbours / interleaved cells per warp, batched / compiler-generated shared-memory atomics
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_assembly.py tests/test_gpu_unstructured.py -x -q -m gpu -k "push or tiles256" > gpurun_out/r4a_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/r4a_pytest.log
tail -4 gpurun_out/r4a_pytest.log
python -c "
from phifem_b200 import build
print(build.build_variant('nobatch', ['PHIFEM_PUSH_BATCH=0'], sources=('assemble_tiles.cu',)))" > gpurun_out/r4a_variant.log 2>&1
B="timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e --no-unstructured --no-solve --no-replan"
show() { python -c "
import json,sys
d=json.load(open(sys.argv[1])); k=d['roofline']['kernels_ms']; print(sys.argv[1], 'step %.3f cells %.3f surface %.3f' % (d['ms_per_step'], k['assemble_cells'], k['assemble_surface']), d['scatter'].get('recompute_factor'))" $1; }
$B --cell-pass push --rows-per-tile 256 > gpurun_out/r4a_push256.json 2> gpurun_out/r4a_push256.err; show gpurun_out/r4a_push256.json
$B --cell-pass push --rows-per-tile 128 > gpurun_out/r4a_push128.json 2> gpurun_out/r4a_push128.err; show gpurun_out/r4a_push128.json
PHIFEM_PUSH_INTERLEAVE=1 $B --cell-pass push --rows-per-tile 256 > gpurun_out/r4a_push256i.json 2> gpurun_out/r4a_push256i.err; show gpurun_out/r4a_push256i.json
PHIFEM_B200_LIB=phifem_b200/libphifem_b200_nobatch.so $B --cell-pass push --rows-per-tile 256 > gpurun_out/r4a_push256nb.json 2> gpurun_out/r4a_push256nb.err; show gpurun_out/r4a_push256nb.json
$B > gpurun_out/r4a_rows.json 2> gpurun_out/r4a_rows.err; show gpurun_out/r4a_rows.json
