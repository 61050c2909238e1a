#!/usr/bin/env python
"""The reference's weak-Dirichlet (dual) demo on the GPU (BASELINE.json configs[0]):

    python demo/weak_dirichlet_flower.py [--n 200]

mirrors `python main.py bg` of reference demo/weak-dirichlet/flower/main.py: tags with single_layer_cut=True
(main.py:53-66), mixed P1 x P1 space (u, p) (:76-82), forms (:112-151), solve (:161-184), u = first component
(:186-188)."""
import argparse
import os
import sys
import time
import warnings

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from phifem_b200 import assemble, fem, solve, synthetic  # noqa: E402
from phifem.mesh_scripts import compute_tags_measures  # noqa: E402


def main(n=200, quiet=False):
    t0 = time.perf_counter()
    bg_mesh = synthetic.rectangle_mesh(n, lo=(-4.5, -4.5), hi=(4.5, 4.5))
    V = fem.functionspace(bg_mesh, ("Lagrange", 1))
    detection_levelset_h = fem.Function(V).interpolate(synthetic.flower_detection)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", RuntimeWarning)
        cells_tags, facets_tags, _, ds_bdy, _ = compute_tags_measures(
            bg_mesh, detection_levelset_h, 1, box_mode=True, single_layer_cut=True)
    phi_h = fem.Function(V).interpolate(synthetic.flower_levelset)
    f_h = fem.Function(V).interpolate(synthetic.flower_source)
    u_D = fem.Function(V).interpolate(lambda x: np.zeros_like(x[0]))
    plan = assemble.build_plan_weak_dirichlet(bg_mesh, cells_tags, facets_tags, ds_bdy(100), V=V)
    A, b = assemble.assemble_weak_dirichlet(plan, phi_h, f_h, u_D, pen_coef=1.0, stab_coef=1.0)
    sol, info = solve.bicgstab(A, b, rtol=1e-10)
    if not info.converged:
        raise RuntimeError("the linear solve did not converge: %r" % info)
    torch.cuda.synchronize()
    u_h = sol[0::2]
    if not quiet:
        print("cells %d (inside %d, cut %d), mixed dofs %d, nnz %d, %.1f ms in total" % (
            bg_mesh.num_cells, len(cells_tags.find(1)), len(cells_tags.find(2)), plan.n_rows, plan.nnz,
            (time.perf_counter() - t0) * 1e3))
        print(info)
        print("u_h: min %.4e max %.4e" % (float(u_h.min()), float(u_h.max())))
    return u_h, info


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=200)
    main(ap.parse_args().n)
