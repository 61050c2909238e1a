#!/bin/bash
# round 2 (session 3), GPU call I: ncu evidence after the staged tag kernels -- launch list of the default step, --set full
# of every kernel of the step (with source counters for the cell row pass)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-unstructured --no-solve --no-replan"
$CMD > gpurun_out/i_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -c 400 --csv --log-file gpurun_out/i_launches.csv $CMD > gpurun_out/i_ncu1.log 2>&1
echo "launch list exit $?"
for spec in "k_assemble_rows_p1 8 2 rows" "k_tag_cells_p1_staged 4 1 tagcells" "k_tag_facets_staged 4 1 tagfacets" "k_tag_boundary_facets_rec 4 1 tagbnd" "k_surface_once_p1 4 1 once"; do
  set -- $spec
  $CMD > gpurun_out/i_plain2.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:^$1 -s $2 -c $3 -o gpurun_out/i_step_$4 -f $CMD > gpurun_out/i_ncu_$4.log 2>&1
  echo "$1 full exit $?"
done
ls -la gpurun_out/i_* | head -30
