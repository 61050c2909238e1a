#!/bin/bash
# round 2 (session 3), GPU call C: staged facet kernel mismatch; tag kernels one by one, occupancy / stage variants
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=64 python tools/r3_debug_facets.py 2>&1 | tail -12
python tools/r3_debug_facets.py 2>&1 | tail -12
python tools/r3_tagbench.py 2>&1 | tail -2
for v in c4 s4; do
  PHIFEM_B200_LIB=$PWD/phifem_b200/libphifem_b200_$v.so python tools/r3_tagbench.py 2>&1 | tail -2
done
