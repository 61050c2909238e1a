#!/bin/bash
# round 2, GPU call F: ncu evidence -- launch list of the default step, --set full of each kernel of the step (DRAM
# traffic), --set full of the row-gather cell pass on the unstructured (Morton-renumbered) mesh
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-unstructured --no-solve --no-replan"
$CMD > gpurun_out/f_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -c 400 --csv --log-file gpurun_out/f_launches.csv $CMD > gpurun_out/f_ncu1.log 2>&1
echo "launch list exit $?"
for spec in "k_assemble_rows_p1 8 2 rows" "k_tag_cells_p1 4 1 tagcells" "k_tag_facets 4 1 tagfacets" "k_tag_boundary_facets 4 1 tagbnd" "k_surface_once_p1 4 1 once" "k_vertex_class 4 1 vclass"; do
  set -- $spec
  $CMD > gpurun_out/f_plain2.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:^$1 -s $2 -c $3 -o gpurun_out/f_step_$4 -f $CMD > gpurun_out/f_ncu_$4.log 2>&1
  echo "$1 full exit $?"
done
CMDU="python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-solve --no-replan --mesh unstructured"
$CMDU > gpurun_out/f_plain3.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:^k_assemble_rows_p1 -s 8 -c 2 -o gpurun_out/f_unstructured_rows -f $CMDU > gpurun_out/f_ncu3.log 2>&1
echo "unstructured full exit $?"
ls -la gpurun_out/f_* | head -30
