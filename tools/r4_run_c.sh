#!/bin/bash
# round 2 (session 4), GPU call C: TIMING ONLY -- the push kernel with plain (racy) shared-memory updates: the upper bound of
# what a conflict-free (coloured) cell-once pass could reach; R = 256 / 128, neighbours / interleaved
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -c "
from phifem_b200 import build
print(build.build_variant('racy', ['PHIFEM_PUSH_RACY=1'], sources=('assemble_tiles.cu',)))
print(build.build_variant('racy3', ['PHIFEM_PUSH_RACY=1', 'PHIFEM_TILES_MINBLOCKS=3', 'PHIFEM_TILES_MINBLOCKS_128=6'], sources=('assemble_tiles.cu',)))" > gpurun_out/r4c_variant.log 2>&1
B="timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e --no-unstructured --no-solve --no-replan"
show() { python -c "
import json,sys
d=json.load(open(sys.argv[1])); k=d['roofline']['kernels_ms']; print(sys.argv[1], 'step %.3f cells %.3f surface %.3f' % (d['ms_per_step'], k['assemble_cells'], k['assemble_surface']), d['scatter'].get('recompute_factor'))" $1; }
export PHIFEM_B200_LIB=phifem_b200/libphifem_b200_racy.so
$B --cell-pass push --rows-per-tile 256 > gpurun_out/r4c_racy256.json 2> gpurun_out/r4c_racy256.err; show gpurun_out/r4c_racy256.json
$B --cell-pass push --rows-per-tile 128 > gpurun_out/r4c_racy128.json 2> gpurun_out/r4c_racy128.err; show gpurun_out/r4c_racy128.json
PHIFEM_PUSH_INTERLEAVE=1 $B --cell-pass push --rows-per-tile 256 > gpurun_out/r4c_racy256i.json 2> gpurun_out/r4c_racy256i.err; show gpurun_out/r4c_racy256i.json
export PHIFEM_B200_LIB=phifem_b200/libphifem_b200_racy3.so
$B --cell-pass push --rows-per-tile 256 > gpurun_out/r4c_racy3_256.json 2> gpurun_out/r4c_racy3_256.err; show gpurun_out/r4c_racy3_256.json
$B --cell-pass push --rows-per-tile 128 > gpurun_out/r4c_racy3_128.json 2> gpurun_out/r4c_racy3_128.err; show gpurun_out/r4c_racy3_128.json
