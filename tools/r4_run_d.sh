#!/bin/bash
# round 2 (session 4), GPU call D: persistent, cp.async-pipelined P_k cell kernel (k_assemble_cells_pk_pipe): parity,
# config C (2d-p2) against the one-cell-per-thread kernel
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_assembly_pk.py tests/test_gpu_convergence.py -x -q -m gpu > gpurun_out/r4d_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/r4d_pytest.log
tail -4 gpurun_out/r4d_pytest.log
python -c "
from phifem_b200 import build
print(build.build_variant('nopipe', ['PHIFEM_PK_PIPE=0'], sources=('assemble_pk.cu',)))
print(build.build_variant('pipe3', ['PHIFEM_PK_PIPE_MINBLOCKS=3'], sources=('assemble_pk.cu',)))
print(build.build_variant('pipe5', ['PHIFEM_PK_PIPE_MINBLOCKS=5'], sources=('assemble_pk.cu',)))" > gpurun_out/r4d_variant.log 2>&1
B="timeout 600 python bench.py --config 2d-p2 --steps 10 --warmup 3 --no-cpu --no-e2e --no-unstructured --no-solve --no-replan"
show() { python -c "
import json,sys
d=json.load(open(sys.argv[1])); k=d['roofline']['kernels_ms']; print(sys.argv[1], 'step %.3f' % d['ms_per_step'], {a: round(b,3) for a,b in k.items()})" $1; }
$B > gpurun_out/r4d_pipe.json 2> gpurun_out/r4d_pipe.err; show gpurun_out/r4d_pipe.json
for v in nopipe pipe3 pipe5; do
PHIFEM_B200_LIB=phifem_b200/libphifem_b200_$v.so $B > gpurun_out/r4d_$v.json 2> gpurun_out/r4d_$v.err; show gpurun_out/r4d_$v.json
done
