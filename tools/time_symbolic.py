#!/usr/bin/env python
"""Time the symbolic phase at the headline size (config E, n = 204): the C-ABI pattern builder (csrc/symbolic.cu) against
the torch-built plan of phifem_b200/assemble.py (method="atomic": pattern + slot maps only), and check they agree."""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from phifem_b200 import assemble, fem, mesh_scripts, symbolic, synthetic  # noqa: E402
from phifem_b200.mesh import MeshTags  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 204
mesh = synthetic.box_mesh(n, device="cuda")
mesh.c2f
mesh.detj_bounds()
phi = synthetic.sphere_levelset(mesh.x)
dls = mesh_scripts._DeviceLevelset(mesh, fem.Function(fem.functionspace_p1_device(mesh), phi), 1)
ws = mesh_scripts.TagWorkspace(mesh)
mesh_scripts.classify(mesh, dls, ws=ws)
ctags, ftags = MeshTags(mesh, 3, ws.cell_tags), MeshTags(mesh, 2, ws.facet_tags)
ctags.tags8, ftags.tags8 = ws.cell_tags8, ws.facet_tags8
ents = mesh_scripts._integration_entities_dev(mesh, ws.cell_tags8, ws.facet_tags8, 4, (1, 2))
out = {"n": n, "cells": mesh.num_cells}
for name, fn in (("capi_ms", lambda: symbolic.DevicePattern(mesh, ctags, ftags, ents)),
                 ("torch_ms", lambda: assemble.build_plan(mesh, ctags, ftags, ents, method="atomic"))):
    res = None
    best = 1e30
    for _ in range(3):
        res = None
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        res = fn()
        torch.cuda.synchronize()
        best = min(best, (time.perf_counter() - t0) * 1e3)
    out[name] = best
    out[name.replace("_ms", "_nnz")] = res.nnz
    if name == "capi_ms":
        pat = res
    else:
        plan = res
out["identical"] = all(torch.equal(getattr(pat, k), getattr(plan, k))
                       for k in ("indptr", "indices", "active", "ghost", "slots_cells", "slots_ghost", "slots_boundary"))
print(json.dumps(out))
