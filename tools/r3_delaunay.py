"""Row-gather passes on a genuinely unstructured mesh: Delaunay tetrahedra of random points in the unit cube, renumbered
along the Morton curve or in count-balanced pencils (Mesh.reordered).  Tags + plan + cell / surface passes, CUDA events."""
import os, sys, time, warnings
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import torch
from scipy.spatial import Delaunay
from phifem_b200 import assemble, fem, mesh_scripts, synthetic
from phifem_b200.mesh import Mesh

npts = int(os.environ.get("NPTS", "400000"))
rng = np.random.default_rng(0)
t0 = time.perf_counter()
pts = rng.uniform(0.0, 1.0, (npts, 3))
tri = Delaunay(pts)
cells = tri.simplices.astype(np.int64)
# drop slivers (|det| below 1e-3 of the mean cell volume * 6): they are legal but make every fp64 tolerance moot
e = pts[cells[:, 1:]] - pts[cells[:, :1]]
det = np.abs(np.linalg.det(e))
keep = det > 1e-3 * det.mean()
cells = cells[keep]
print("delaunay: %d points, %d tetrahedra (%d slivers dropped) in %.1f s" % (npts, len(cells), int((~keep).sum()), time.perf_counter() - t0), flush=True)
base = Mesh(torch.from_numpy(pts), torch.from_numpy(cells.astype(np.int32)), "tetrahedron", "cuda")


def timed(fn_, reps=20):
    for _ in range(3):
        fn_()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
    ev[0].record()
    for i in range(reps):
        fn_()
        ev[i + 1].record()
    torch.cuda.synchronize()
    ts = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(reps))
    return ts[len(ts) // 2]


for curve in ("none", "morton", "pencil"):
    mesh = base if curve == "none" else base.reordered(curve)
    phi = synthetic.sphere_levelset(mesh.x)
    f = synthetic.ball_source(mesh.x)
    fn = fem.Function(fem.functionspace_p1_device(mesh), phi)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", RuntimeWarning)
        ct, ft, _, ds, _ = mesh_scripts.compute_tags_measures(mesh, fn, 1, box_mode=True)
    plan = assemble.build_plan(mesh, ct, ft, ds(100))
    data, b = plan.new_outputs()
    dls = mesh_scripts._DeviceLevelset(mesh, fn, 1)
    ws = mesh_scripts.TagWorkspace(mesh)
    from phifem_b200 import rows as rows_mod
    res = {
        "tag_cells": timed(lambda: mesh_scripts.classify_cells(mesh, dls, ws)),
        "tag_facets": timed(lambda: mesh_scripts.classify_facets(mesh, dls, ws)),
        "cells": timed(lambda: rows_mod.assemble_rows_into(plan.rowsplan, phi, f, 1.0, data, b, passes=("cells",))),
        "surface": timed(lambda: rows_mod.assemble_rows_into(plan.rowsplan, phi, f, 1.0, data, b, passes=("surface",))),
    }
    na = int(((ct.tags8 == 1) | (ct.tags8 == 2)).sum())
    print(curve, "cells", mesh.num_cells, "active", na, "nnz", plan.nnz, {k: round(v, 4) for k, v in res.items()},
          "ns per active cell (cell pass): %.3f" % (res["cells"] * 1e6 / na), flush=True)
