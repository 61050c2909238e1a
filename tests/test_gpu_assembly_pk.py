"""GPU parity of the P1 / P2 quadrature kernels (csrc/assemble_pk.cu, through the C ABI) against the
oracle's brute-force quadrature (oracle/assembly.py *_quadrature, exact to degree 11).

Bar: identical CSR sparsity; entries within 1e-12 of the row scale; load vector within 1e-12 of max|b|."""
import warnings

import numpy as np
import pytest
import torch

from oracle import assembly as OA
from phifem_b200 import assemble, fem, mesh_scripts, synthetic

pytestmark = pytest.mark.gpu
RTOL = 1e-12


def _problem(kind, n, kw, kphi):
    if kind.startswith("tri"):
        mesh = synthetic.rectangle_mesh(n, device="cuda")
        center, radius = (0.013, -0.021), 0.61
    else:
        mesh = synthetic.box_mesh(n, device="cuda")
        center, radius = synthetic.SPHERE_CENTER, 0.37
    if kind.endswith("unstructured"):
        mesh = synthetic.unstructured_variant(mesh, jitter=0.2, seed=7)
    V, Vp, V1 = fem.functionspace(mesh, kw), fem.functionspace(mesh, kphi), fem.functionspace(mesh, 1)
    # tags from the P1 detection level set (detection_degree 1, as the demo does at main.py:59-61)
    det = fem.Function(V1, synthetic.sphere_levelset(mesh.x, center=center, radius=radius).cpu().numpy())
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", RuntimeWarning)
        ctags, ftags, _, ds, _ = mesh_scripts.compute_tags_measures(mesh, det, 1, box_mode=True)
    # phi_h = interpolant of the smooth level set in the level-set space (main.py:85-86)
    phi = synthetic.sphere_levelset(Vp.dof_coordinates_dev(), center=center, radius=radius)
    f = torch.from_numpy(np.random.default_rng(1234).uniform(-1, 1, V.num_dofs)).cuda()
    return mesh, V, Vp, ctags, ftags, ds, phi, f


def _row_scale(indptr, data):
    scale = np.zeros(len(indptr) - 1)
    rows = np.repeat(np.arange(len(indptr) - 1), np.diff(indptr))
    np.maximum.at(scale, rows, np.abs(data))
    return scale, rows


@pytest.mark.parametrize("kw,kphi", [(2, 2), (2, 1), (1, 2), (1, 1)])
@pytest.mark.parametrize("kind,n", [("tri", 14), ("tri-unstructured", 10), ("tet", 5), ("tet-unstructured", 4)])
def test_pk_operator_matches_oracle_quadrature(kind, n, kw, kphi):
    mesh, V, Vp, ctags, ftags, ds, phi, f = _problem(kind, n, kw, kphi)
    plan = assemble.build_plan(mesh, ctags, ftags, ds(100), V=V, V_phi=Vp, method="pk")
    assert plan.method == "pk-atomic"
    A, b = assemble.assemble_strong_dirichlet(plan, phi, f, stab_coef=0.9)
    x = mesh.x.cpu().numpy()
    cells = mesh.cells.cpu().numpy().astype(np.int64)
    ip, ix, data, bo = OA.assemble_strong_dirichlet(
        x, cells, V.dofmap.astype(np.int64), V.num_dofs, phi.cpu().numpy(), f.cpu().numpy(),
        ctags.values_dev.cpu().numpy(), ftags.values_dev.cpu().numpy(), mesh.c2f.cpu().numpy(),
        mesh.f2c.cpu().numpy(), ds(100).integration_entities, sigma=0.9, method="quadrature", kphi=kphi,
        kw=kw, phi_dofmap=Vp.dofmap.astype(np.int64))
    assert np.array_equal(A.indptr.cpu().numpy(), ip)
    assert np.array_equal(A.indices.cpu().numpy(), ix)
    assert plan.ghost.numel() > 0 and plan.entities.shape[0] > 0
    got = A.data.cpu().numpy()
    scale, rows = _row_scale(ip, data)
    assert np.all(np.abs(got - data) <= RTOL * scale[rows])
    assert np.all(np.abs(b.cpu().numpy() - bo) <= RTOL * np.abs(bo).max())


def test_pk_p1_agrees_with_closed_form_kernels():
    """kw = kphi = 1 through the quadrature kernels == the closed-form row-gather kernels."""
    mesh, V, Vp, ctags, ftags, ds, phi, f = _problem("tet-unstructured", 8, 1, 1)
    plan_q = assemble.build_plan(mesh, ctags, ftags, ds(100), V=V, V_phi=Vp, method="pk")
    plan_c = assemble.build_plan(mesh, ctags, ftags, ds(100))
    Aq, bq = assemble.assemble_strong_dirichlet(plan_q, phi, f, stab_coef=1.0)
    Ac, bc = assemble.assemble_strong_dirichlet(plan_c, phi, f, stab_coef=1.0)
    assert torch.equal(Aq.indptr, Ac.indptr) and torch.equal(Aq.indices, Ac.indices)
    scale, rows = _row_scale(Ac.indptr.cpu().numpy(), Ac.data.cpu().numpy())
    assert np.all(np.abs((Aq.data - Ac.data).cpu().numpy()) <= RTOL * scale[rows])
    assert float((bq - bc).abs().max()) <= RTOL * float(bc.abs().max())


@pytest.mark.parametrize("kind,n", [("tri", 700), ("tet", 40)])
def test_pk_p2_properties_at_scale(kind, n):
    """Size-independent properties of the P2 operator on ~1 M triangles / 0.4 M tetrahedra: with phi == 1
    the form is int grad w . grad v + ghost jumps (constants in its kernel, symmetric), sum(b) = |Omega_h|,
    and the P2 stiffness reproduces int |grad u|^2 exactly for a quadratic u."""
    mesh, V, Vp, ctags, ftags, ds, phi, f = _problem(kind, n, 2, 2)
    plan = assemble.build_plan(mesh, ctags, ftags, None, V=V, V_phi=Vp)
    one_phi = torch.ones(Vp.num_dofs, dtype=torch.float64, device="cuda")
    one_f = torch.ones(V.num_dofs, dtype=torch.float64, device="cuda")
    A1, _ = assemble.assemble_strong_dirichlet(plan, one_phi, one_f, stab_coef=1.0)
    M1 = A1.to_scipy()
    assert np.abs(M1 @ np.ones(M1.shape[0])).max() <= 1e-10 * abs(M1).max()
    assert abs(M1 - M1.T).max() <= 1e-12 * abs(M1).max()
    A, b = assemble.assemble_strong_dirichlet(plan, one_phi, one_f, stab_coef=0.0)
    M = A.to_scipy()
    tags = ctags.values_dev
    d = mesh.gdim
    vol_cell = (2.0 / n) ** 2 / 2.0 if d == 2 else (1.0 / n) ** 3 / 6.0
    want = float(((tags == 1) | (tags == 2)).sum()) * vol_cell
    assert abs(float(b.sum()) - want) <= 1e-10 * want
    # u = |x|^2 is in P2: u^T A u = int_{Omega_h} |grad u|^2 = 4 int |x|^2, summed cell by cell from the
    # exact integral of a quadratic (vertex + edge-midpoint rule is exact for degree 2 on simplices: use
    # the P2 mass lumping identity int q = |K| * mean of q at the edge midpoints in 2D)
    X = V.dof_coordinates_dev()
    u = (X * X).sum(dim=1).cpu().numpy()
    energy = float(u @ (M @ u))
    act = torch.nonzero((tags == 1) | (tags == 2)).reshape(-1)
    dm = V.dofmap_dev[act].long()
    nv = d + 1
    q = 4.0 * (X * X).sum(dim=1)
    if d == 2:
        exact = float((q[dm[:, nv:]].mean(dim=1) * vol_cell).sum())
    else:
        # tetrahedron: int q = |K| (4/5 mean of edge midpoints - ... ) -> use the degree-2 rule
        # int q = |K| * (-1/20 sum_vertices q + 1/5 sum_edge_midpoints q)
        exact = float(((-q[dm[:, :nv]].sum(dim=1) / 20.0 + q[dm[:, nv:]].sum(dim=1) / 5.0) * vol_cell).sum())
    assert abs(energy - exact) <= 1e-10 * exact
