#!/bin/bash
# round 2 (session 3), GPU call Z: final evidence -- launch list and --set full of the step kernels of the final build,
# default bench line, reference arm
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-unstructured --no-solve --no-replan"
$CMD > gpurun_out/z_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -c 400 --csv --log-file gpurun_out/z_launches.csv $CMD > gpurun_out/z_ncu1.log 2>&1
echo "launch list exit $?"
for spec in "k_assemble_rows_p1 8 2 rows" "k_tag_cells_p1_staged 4 1 tagcells" "k_tag_facets_staged 4 1 tagfacets" "k_tag_boundary_facets_rec 4 1 tagbnd" "k_surface_fill_p1 4 1 fill"; do
  set -- $spec
  $CMD > gpurun_out/z_plain2.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:^$1 -s $2 -c $3 -o gpurun_out/z_step_$4 -f $CMD > gpurun_out/z_ncu_$4.log 2>&1
  echo "$1 full exit $?"
done
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/z_bench_reference.json 2> gpurun_out/z_bench_reference.err
python bench.py > gpurun_out/z_bench_default.json 2> gpurun_out/z_bench_default.err
python -c "
import json; d=json.load(open('gpurun_out/z_bench_default.json')); u=d['unstructured']; print(d['ms_per_step'], d['value'], 'e2e', d['e2e']['ms_per_step'], 'sym', d['symbolic_ms'], 'unstructured', u['ms_per_step'], d['roofline']['frac'], d['roofline']['step_frac'])"
