#!/bin/bash
# round 2 (session 4), GPU call W: SpMV of the fused BiCGStab iteration with its loads batched per lane (PHIFEM_SPMV_BATCH)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_solve.py -x -q -m gpu > gpurun_out/r4w_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/r4w_pytest.log
tail -3 gpurun_out/r4w_pytest.log
python -c "
from phifem_b200 import build
print(build.build_variant('spmv1', ['PHIFEM_SPMV_BATCH=1'], sources=('solve.cu',)))
print(build.build_variant('spmv2', ['PHIFEM_SPMV_BATCH=2'], sources=('solve.cu',)))" > gpurun_out/r4w_variant.log 2>&1
for v in "" _spmv1 _spmv2; do
echo "== libphifem_b200$v.so"
PHIFEM_B200_LIB=$PWD/phifem_b200/libphifem_b200$v.so python tools/r3_solve_time.py 2>&1 | tail -3
done > gpurun_out/r4w_solve.txt 2>&1
cat gpurun_out/r4w_solve.txt
