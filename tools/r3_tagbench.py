"""Tag kernels of config E timed one by one (CUDA events, 30 repetitions after 5 warm-ups; the mesh streams > L2 per pass)."""
import os, sys, warnings
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from phifem_b200 import synthetic, fem, mesh_scripts

N = int(os.environ.get("N", "204"))
mesh = synthetic.box_mesh(N, device="cuda")
phi = synthetic.sphere_levelset(mesh.x)
fn = fem.Function(fem.functionspace_p1_device(mesh), phi)
dls = mesh_scripts._DeviceLevelset(mesh, fn, 1)
ws = mesh_scripts.TagWorkspace(mesh)


def timed(fn_, reps=30):
    for _ in range(5):
        fn_()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
    ev[0].record()
    for i in range(reps):
        fn_()
        ev[i + 1].record()
    torch.cuda.synchronize()
    ts = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(reps))
    return ts[len(ts) // 2]


for kind in ("ldg", "staged"):
    os.environ["PHIFEM_FACETS_KERNEL"] = kind
    os.environ["PHIFEM_CELLS_KERNEL"] = kind
    res = {
        "cells": timed(lambda: mesh_scripts.classify_cells(mesh, dls, ws)),
        "facets_interior": timed(lambda: mesh_scripts.classify_facets(mesh, dls, ws, mesh_scripts.FACETS_INTERIOR)),
        "facets_boundary": timed(lambda: mesh_scripts.classify_facets(mesh, dls, ws, mesh_scripts.FACETS_BOUNDARY)),
        "facets_both": timed(lambda: mesh_scripts.classify_facets(mesh, dls, ws)),
    }
    print(os.environ.get("PHIFEM_B200_LIB", "default").split("_")[-1], kind, {k: round(v, 4) for k, v in res.items()})
