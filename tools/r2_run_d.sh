#!/bin/bash
# round 2, GPU call D: the whole GPU suite with the native symbolic phase as the default, symbolic timing
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/d_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/d_pytest.log
tail -6 gpurun_out/d_pytest.log
python - > gpurun_out/d_symbolic.txt 2>&1 <<'PY'
import time, torch, warnings
from phifem_b200 import assemble, fem, mesh_scripts, synthetic
from phifem_b200.mesh import MeshTags
mesh = synthetic.box_mesh(204, device="cuda")
phi = synthetic.sphere_levelset(mesh.x)
mesh.c2f; mesh.detj_bounds()
dls = mesh_scripts._DeviceLevelset(mesh, fem.Function(fem.functionspace_p1_device(mesh), phi), 1)
ws = mesh_scripts.classify(mesh, dls)
ents = mesh_scripts._integration_entities_dev(mesh, ws.cell_tags8, ws.facet_tags8, 4, (1, 2))
ct, ft = MeshTags(mesh, 3, None, tags8=ws.cell_tags8), MeshTags(mesh, 2, None, tags8=ws.facet_tags8)
for sym in ("native", "native", "native", "torch", "torch", "native"):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    plan = assemble.build_plan(mesh, ct, ft, ents, symbolic=sym)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) * 1e3
    print(sym, "%.1f ms" % dt, "nnz", plan.nnz, "plan bytes", plan.rowsplan.index_bytes())
    del plan
PY
cat gpurun_out/d_symbolic.txt
python bench.py --no-cpu --no-e2e --no-unstructured --no-solve > gpurun_out/d_bench.json 2> gpurun_out/d_bench.err; echo "bench exit $?"
python -c "
import json; d=json.load(open('gpurun_out/d_bench.json')); print(d['ms_per_step'], d['symbolic_ms'], d['gpu_launches_per_step'], d['gpu_kernels'])"
