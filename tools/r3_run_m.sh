#!/bin/bash
# round 2 (session 3), GPU call M: cell pass with the next record staged in shared memory by cp.async (5 / 4 / 6 CTAs per SM)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
PHIFEM_ROWS_ASYNC=1 timeout 900 python -m pytest tests/test_gpu_assembly.py tests/test_gpu_unstructured.py -x -q -m gpu 2>&1 | tail -3
for v in off default a4 a6; do
  lib=$PWD/phifem_b200/libphifem_b200.so; as=1
  [ $v = a4 ] && lib=$PWD/phifem_b200/libphifem_b200_a4.so
  [ $v = a6 ] && lib=$PWD/phifem_b200/libphifem_b200_a6.so
  [ $v = off ] && as=0
  PHIFEM_ROWS_ASYNC=$as PHIFEM_B200_LIB=$lib python bench.py --no-cpu --no-e2e --no-solve --no-replan --steps 20 > gpurun_out/r3m_bench_$v.json 2> gpurun_out/r3m_bench_$v.err
  python -c "
import json; d=json.load(open('gpurun_out/r3m_bench_$v.json')); k=d['roofline']['kernels_ms']; u=d['unstructured']['kernels_ms']; print('$v', round(d['ms_per_step'],4), 'cells', round(k['assemble_cells'],4), 'unstructured cells', round(u['assemble_cells'],4))" || tail -3 gpurun_out/r3m_bench_$v.err
done
