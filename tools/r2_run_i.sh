#!/bin/bash
# round 2, GPU call I: occupancy / block-size variants of the row-gather kernels (build variants, one bench line each)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for v in "" _once8 _once6 _surf6 _surf8 _b64m8 _b256m2 _b96m5 _b64m9; do
  PHIFEM_B200_LIB=$PWD/phifem_b200/libphifem_b200$v.so timeout 300 python bench.py --no-cpu --no-e2e --no-unstructured --no-solve --no-replan --steps 10 > gpurun_out/i_bench$v.json 2> gpurun_out/i_bench$v.err
  python -c "
import json
try:
    d=json.load(open('gpurun_out/i_bench$v.json')); k=d['roofline']['kernels_ms']
    print('variant[$v]', round(d['ms_per_step'],4), 'cells', round(k['assemble_cells'],4), 'surface', round(k['assemble_surface'],4), 'assembly', round(k['assembly'],4))
except Exception as e: print('variant[$v] failed', e)
"
done
PHIFEM_B200_LIB=$PWD/phifem_b200/libphifem_b200.so python bench.py --no-cpu --no-e2e --no-unstructured --no-solve --no-replan --steps 10 --order morton 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); k=d['roofline']['kernels_ms']; print('order morton', d['ms_per_step'], k['assemble_cells'])"
