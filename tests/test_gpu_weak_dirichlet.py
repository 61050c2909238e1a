"""GPU parity of the weak-Dirichlet (dual) phi-FEM operator (reference demo/weak-dirichlet/flower/main.py:112-154,
BASELINE.json configs[0]) through the C ABI against the oracle, including the demo's own configuration."""
import os
import warnings

import numpy as np
import pytest
import torch

import cases
from oracle import assembly as OA
from oracle import tags as OT
from phifem_b200 import assemble, fem, mesh_scripts, quadrature, synthetic

pytestmark = pytest.mark.gpu
RTOL = 1e-12
HERE = os.path.dirname(os.path.abspath(__file__))


def _row_scale(indptr, data):
    scale = np.zeros(len(indptr) - 1)
    rows = np.repeat(np.arange(len(indptr) - 1), np.diff(indptr))
    np.maximum.at(scale, rows, np.abs(data))
    return scale, rows


def _compare(A, b, ip, ix, data, bo):
    assert np.array_equal(A.indptr.cpu().numpy(), ip)
    assert np.array_equal(A.indices.cpu().numpy(), ix)
    scale, rows = _row_scale(ip, data)
    gscale = np.abs(data).max()
    # rows holding only structural zeros (p dofs away from the cut cells) must be exactly zero
    assert np.all(np.abs(A.data.cpu().numpy() - data) <= RTOL * np.maximum(scale[rows], 1e-300 * gscale))
    assert np.all(np.abs(b.cpu().numpy() - bo) <= RTOL * np.abs(bo).max())


@pytest.mark.parametrize("kw,kphi", [(1, 1), (2, 2), (2, 1), (1, 2)])
@pytest.mark.parametrize("kind,n", [("tri", 14), ("tri-unstructured", 10), ("tet", 5), ("tet-unstructured", 4)])
def test_weak_operator_matches_oracle(kind, n, kw, kphi):
    if kind.startswith("tri"):
        mesh = synthetic.rectangle_mesh(n, device="cuda")
        center, radius = (0.013, -0.021), 0.61
    else:
        mesh = synthetic.box_mesh(n, device="cuda")
        center, radius = synthetic.SPHERE_CENTER, 0.37
    if kind.endswith("unstructured"):
        mesh = synthetic.unstructured_variant(mesh, jitter=0.2, seed=7)
    V, Vp, V1 = fem.functionspace(mesh, kw), fem.functionspace(mesh, kphi), fem.functionspace(mesh, 1)
    det = fem.Function(V1, synthetic.sphere_levelset(mesh.x, center=center, radius=radius).cpu().numpy())
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", RuntimeWarning)
        ctags, ftags, _, ds, _ = mesh_scripts.compute_tags_measures(mesh, det, 1, box_mode=True)
    phi = synthetic.sphere_levelset(Vp.dof_coordinates_dev(), center=center, radius=radius)
    rng = np.random.default_rng(4321)
    f = torch.from_numpy(rng.uniform(-1, 1, V.num_dofs)).cuda()
    ud = torch.from_numpy(rng.uniform(-1, 1, V.num_dofs)).cuda()
    plan = assemble.build_plan_weak_dirichlet(mesh, ctags, ftags, ds(100), V=V, V_phi=Vp)
    A, b = assemble.assemble_weak_dirichlet(plan, phi, f, ud, pen_coef=1.3, stab_coef=0.7)
    assert A.shape == (2 * V.num_dofs, 2 * V.num_dofs) and plan.ghost.numel() > 0 and plan.entities.shape[0] > 0
    x = mesh.x.cpu().numpy()
    cells = mesh.cells.cpu().numpy().astype(np.int64)
    ref = OA.assemble_weak_dirichlet(
        x, cells, V.dofmap.astype(np.int64), V.num_dofs, phi.cpu().numpy(), f.cpu().numpy(), ud.cpu().numpy(),
        ctags.values_dev.cpu().numpy(), ftags.values_dev.cpu().numpy(), mesh.c2f.cpu().numpy(),
        mesh.f2c.cpu().numpy(), ds(100).integration_entities, gamma=1.3, sigma=0.7,
        method="closed_form" if (kw, kphi) == (1, 1) else "quadrature", kphi=kphi, kw=kw,
        phi_dofmap=Vp.dofmap.astype(np.int64))
    _compare(A, b, *ref)
    with pytest.raises(ValueError, match="assemble_weak_dirichlet"):
        assemble.assemble_strong_dirichlet(plan, phi, f)


def test_weak_dirichlet_demo_configuration():
    """BASELINE.json configs[0]: `python main.py bg` of demo/weak-dirichlet/flower -- 200 x 200 background mesh of
    [-4.5, 4.5]^2, flower level sets, P1 x P1, single_layer_cut=True, box mode (main.py:44-66,112-154).  Inputs
    are the reference's own functions evaluated here once (tests/golden/flower_demo.npz)."""
    gold = np.load(os.path.join(HERE, "golden", "flower_demo.npz"))
    n = int(gold["n"])
    mesh = synthetic.rectangle_mesh(n, lo=(-4.5, -4.5), hi=(4.5, 4.5), device="cuda")
    x = mesh.x.cpu().numpy()
    # our restatement of the demo's data functions agrees with the reference's values at the mesh vertices
    assert np.abs(cases.flower_detection(x.T) - gold["detection"]).max() < 1e-13
    assert np.abs(cases.flower_levelset(x.T) - gold["levelset"]).max() < 1e-13
    assert np.array_equal(cases.flower_source(x.T), gold["source"])
    V = fem.functionspace(mesh, 1)
    det = fem.Function(V, gold["detection"])
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", RuntimeWarning)
        ctags, ftags, _, ds, _ = mesh_scripts.compute_tags_measures(mesh, det, 1, box_mode=True,
                                                                   single_layer_cut=True)
    cells = mesh.cells.cpu().numpy().astype(np.int64)
    pts = OT.cell_detection_points("triangle", 1)
    fpts = OT.facet_points_in_cell("triangle", 1)
    ftab = np.asarray([OT.coordinate_basis("triangle", p)[0] for p in fpts])
    out = OT.compute_tags_measures(x, cells, "triangle", gold["detection"][cells],
                                   OT.point_values_function(gold["detection"], cells, ftab), box_mode=True,
                                   single_layer_cut=True, detection_points=pts)
    assert np.array_equal(ctags.values_dev.cpu().numpy(), out["cell_tags"])
    assert np.array_equal(ftags.values_dev.cpu().numpy(), out["facet_tags"])
    assert np.array_equal(ds(100).integration_entities, out["ds100"])
    hist = np.bincount(out["cell_tags"], minlength=4)
    assert hist[1] > 10000 and hist[2] > 500 and hist[3] > 10000 and hist.sum() == 80000
    phi, f, ud = gold["levelset"], gold["source"], gold["dirichlet"]
    plan = assemble.build_plan_weak_dirichlet(mesh, ctags, ftags, ds(100), V=V)
    A, b = assemble.assemble_weak_dirichlet(plan, phi, f, ud, pen_coef=1.0, stab_coef=1.0)
    ref = OA.assemble_weak_dirichlet(x, cells, cells, len(x), phi, f, ud, out["cell_tags"], out["facet_tags"],
                                     out["c2f"], out["f2c"], out["ds100"], gamma=1.0, sigma=1.0)
    _compare(A, b, *ref)
    # the demo's solve (main.py:161-184; MUMPS with null-pivot detection): rows that hold only structural zeros
    # are dropped; the discrete solution must vanish near the boundary and be positive in the petal with the source
    import scipy.sparse.linalg as spla
    M = A.to_scipy().tocsr()
    keep = np.nonzero(np.asarray(abs(M).sum(axis=1)).ravel() > 0)[0]
    sol = np.zeros(M.shape[0])
    sol[keep] = spla.spsolve(M[keep][:, keep].tocsc(), b.cpu().numpy()[keep])
    u = sol[0::2]
    assert np.isfinite(u).all() and u.max() > 0.05
    inside = gold["detection"] < -0.5
    assert u[inside].min() > -1e-3
    near = np.abs(gold["levelset"]) < 0.02
    assert near.sum() > 50 and np.abs(u[near]).max() < 0.1 * u.max()


def _weak_error(n):
    import scipy.sparse.linalg as spla
    R, C = 0.62, (0.013, -0.021)
    mesh = synthetic.rectangle_mesh(n, device="cuda")
    V = fem.functionspace(mesh, 1)
    X = mesh.x
    phi = synthetic.sphere_levelset(X, center=C, radius=R)
    x0, x1 = X[:, 0] - C[0], X[:, 1] - C[1]
    w = -torch.exp(x0) * torch.cos(x1)
    u_exact = phi * w
    f = -(4.0 * w + 2.0 * (2.0 * x0 * w + 2.0 * x1 * torch.exp(x0) * torch.sin(x1)))
    det = fem.Function(V, phi)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", RuntimeWarning)
        ctags, ftags, _, ds, _ = mesh_scripts.compute_tags_measures(mesh, det, 1, box_mode=True)
    plan = assemble.build_plan_weak_dirichlet(mesh, ctags, ftags, ds(100), V=V)
    A, b = assemble.assemble_weak_dirichlet(plan, phi, f, None, pen_coef=1.0, stab_coef=1.0)
    M = A.to_scipy().tocsr()
    keep = np.nonzero(np.asarray(abs(M).sum(axis=1)).ravel() > 0)[0]
    sol = np.zeros(M.shape[0])
    sol[keep] = spla.spsolve(M[keep][:, keep].tocsc(), b.cpu().numpy()[keep])
    uh = sol[0::2]
    # L2 error over the interior cells with a degree-4 rule (u_h is P1, u is smooth)
    lam, wq = quadrature.simplex_rule(2, 4)
    cells = mesh.cells[torch.nonzero(ctags.values_dev == 1).reshape(-1)].long().cpu().numpy()
    xc = X.cpu().numpy()[cells]
    xq = np.einsum("qv,mvd->qmd", lam, xc)
    q0, q1 = xq[..., 0] - C[0], xq[..., 1] - C[1]
    ue = (q0 ** 2 + q1 ** 2 - R * R) * (-np.exp(q0) * np.cos(q1))
    e = xc[:, 1:] - xc[:, :1]
    area = 0.5 * np.abs(e[:, 0, 0] * e[:, 1, 1] - e[:, 0, 1] * e[:, 1, 0])
    err2 = float((wq[:, None] * (lam @ uh[cells].T - ue) ** 2 * area[None]).sum())
    nrm2 = float((wq[:, None] * ue ** 2 * area[None]).sum())
    del u_exact
    return (err2 / nrm2) ** 0.5


def test_weak_dirichlet_manufactured_solution_converges():
    errs = [_weak_error(n) for n in (24, 48, 96)]
    rates = [np.log2(errs[i] / errs[i + 1]) for i in range(2)]
    assert errs[-1] < 4e-2 and min(rates) > 1.8, (errs, rates)   # measured: 0.41, 0.099, 0.023
