"""The step after the path on the GPU: CSR SpMV kernel and Jacobi-BiCGStab (phifem_b200/solve.py) against scipy,
and the two demos end to end."""
import os
import sys
import warnings

import numpy as np
import pytest
import torch

from phifem_b200 import assemble, fem, mesh_scripts, solve, synthetic

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _system(kind, n, form):
    if kind == "tri":
        mesh = synthetic.rectangle_mesh(n, device="cuda")
        center, radius = (0.013, -0.021), 0.62
    else:
        mesh = synthetic.unstructured_variant(synthetic.box_mesh(n, device="cuda"), jitter=0.15, seed=4)
        center, radius = synthetic.SPHERE_CENTER, 0.37
    V = fem.functionspace(mesh, 1)
    phi = synthetic.sphere_levelset(mesh.x, center=center, radius=radius)
    f = torch.cos(3.0 * mesh.x[:, 0]) + mesh.x[:, 1]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", RuntimeWarning)
        ct, ft, _, ds, _ = mesh_scripts.compute_tags_measures(mesh, fem.Function(V, phi), 1, box_mode=True)
    if form == "strong":
        A, b = assemble.assemble_strong_dirichlet(assemble.build_plan(mesh, ct, ft, ds(100)), phi, f)
    else:
        plan = assemble.build_plan_weak_dirichlet(mesh, ct, ft, ds(100), V=V)
        A, b = assemble.assemble_weak_dirichlet(plan, phi, f)
    return A, b


@pytest.mark.parametrize("kind,n,form", [("tri", 64, "strong"), ("tet", 14, "strong"), ("tri", 48, "weak")])
def test_spmv_and_bicgstab_match_scipy(kind, n, form):
    import scipy.sparse.linalg as spla
    A, b = _system(kind, n, form)
    M = A.to_scipy().tocsr()
    x = torch.from_numpy(np.random.default_rng(3).uniform(-1, 1, M.shape[0])).cuda()
    y = solve.spmv(A, x).cpu().numpy()
    want = M @ x.cpu().numpy()
    assert np.abs(y - want).max() <= 1e-13 * np.abs(want).max()
    sol, info = solve.bicgstab(A, b, rtol=1e-11)
    assert info.converged and info.residual < 1e-9, info
    d = M.diagonal()
    keep = np.nonzero(d != 0)[0]
    assert info.n_active == len(keep)
    ref = np.zeros(M.shape[0])
    ref[keep] = spla.spsolve(M[keep][:, keep].tocsc(), b.cpu().numpy()[keep])
    got = sol.cpu().numpy()
    assert np.all(got[d == 0] == 0.0)                       # null pivots: unknown set to zero
    assert np.linalg.norm(got - ref) <= 1e-7 * np.linalg.norm(ref)


def test_demos_run_end_to_end():
    sys.path.insert(0, os.path.join(ROOT, "demo"))
    import strong_dirichlet_flower
    import weak_dirichlet_flower
    u1, info1 = strong_dirichlet_flower.main(n=100, degree=1, quiet=True)
    u2, info2 = weak_dirichlet_flower.main(n=100, quiet=True)
    assert info1.converged and info2.converged
    # the two formulations discretise the same Poisson problem on the flower: both solutions are positive in the
    # petal holding the source and agree to discretisation accuracy
    assert float(u1.max()) > 0.05 and float(u2.max()) > 0.05
    assert abs(float(u1.max()) - float(u2.max())) < 0.25 * float(u1.max())
    u3, info3 = strong_dirichlet_flower.main(n=60, degree=2, quiet=True)
    assert info3.converged and abs(float(u3.max()) - float(u1.max())) < 0.25 * float(u1.max())
