#!/bin/bash
# round 2 (session 4), GPU call H: ghost-penalty kernel of the P_k path with the facet's duplicate entries added up in
# shared memory before the reductions; parity, 3d-p2 / 2d-p2 against the plain scatter
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_assembly_pk.py tests/test_gpu_convergence.py tests/test_gpu_edge_cases.py -x -q -m gpu > gpurun_out/r4h_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/r4h_pytest.log
tail -4 gpurun_out/r4h_pytest.log
python -c "
from phifem_b200 import build
print(build.build_variant('nodedupe', ['PHIFEM_PK_GHOST_DEDUPE=0'], sources=('assemble_pk.cu',)))" > gpurun_out/r4h_variant.log 2>&1
show() { python -c "
import json,sys
d=json.load(open(sys.argv[1])); k=d['roofline']['kernels_ms']; print(sys.argv[1], 'step %.3f' % d['ms_per_step'], {a: round(b,3) for a,b in k.items()})" $1; }
for c in 3d-p2 2d-p2; do
B="timeout 600 python bench.py --config $c --steps 10 --warmup 3 --no-cpu --no-e2e --no-unstructured --no-solve --no-replan"
$B > gpurun_out/r4h_$c.json 2> gpurun_out/r4h_$c.err; show gpurun_out/r4h_$c.json
PHIFEM_B200_LIB=phifem_b200/libphifem_b200_nodedupe.so $B > gpurun_out/r4h_${c}_nodedupe.json 2> gpurun_out/r4h_${c}_nodedupe.err; show gpurun_out/r4h_${c}_nodedupe.json
done
