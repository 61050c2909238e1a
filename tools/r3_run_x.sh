#!/bin/bash
# round 2 (session 3), GPU call X: persistent grid of the interior-facet kernel beside the forked boundary kernel
cd "$(dirname "$0")/.."
for c in 6 5 4 3; do
  PHIFEM_FACETS_CTAS_PER_SM=$c python tools/r3_tagbench.py 2>&1 | tail -1 | sed "s/^/ctas=$c /"
done
