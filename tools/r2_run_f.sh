#!/bin/bash
# round 2, GPU call F: ncu evidence -- launch list of the default step, --set full of the step's kernels (traffic),
# --set full of the row-gather cell pass on the unstructured (Morton-renumbered) mesh
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-unstructured --no-solve"
$CMD > gpurun_out/f_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 300 --csv --log-file gpurun_out/f_launches.csv $CMD > gpurun_out/f_ncu1.log 2>&1
echo "launch list exit $?"
$CMD > gpurun_out/f_plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"k_tag_cells_p1|k_tag_facets|k_assemble_rows_p1|k_surface_once_p1|k_tag_boundary_facets|k_vertex_class" -s 60 -c 7 -o gpurun_out/f_step $CMD > gpurun_out/f_ncu2.log 2>&1
echo "step full exit $?"
CMDU="python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-solve --mesh unstructured"
$CMDU > gpurun_out/f_plain3.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"k_assemble_rows_p1|k_tag_cells_p1|k_tag_facets" -s 30 -c 4 -o gpurun_out/f_unstructured $CMDU > gpurun_out/f_ncu3.log 2>&1
echo "unstructured full exit $?"
ls -la gpurun_out/f_*
