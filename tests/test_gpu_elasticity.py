"""GPU parity of the interface-elasticity phi-FEM operator (reference demo/interface-elasticity/main.py:107-275,
BASELINE.json configs[3]) through the C ABI against the oracle (entry formulas for a P1 level set, brute-force
quadrature for P2), with and without Dirichlet conditions, and its meaning: the solved system converges to the demo's
manufactured two-material solution (data.py:43-48)."""
import warnings

import numpy as np
import pytest
import torch

from oracle import elasticity as OE
from phifem_b200 import elasticity, fem, mesh_scripts, synthetic

pytestmark = pytest.mark.gpu
RTOL = 1e-12


def _row_scale(indptr, data):
    scale = np.zeros(len(indptr) - 1)
    rows = np.repeat(np.arange(len(indptr) - 1), np.diff(indptr))
    np.maximum.at(scale, rows, np.abs(data))
    return scale, rows


@pytest.mark.parametrize("kphi,with_bc", [(1, False), (1, True), (2, True)])
@pytest.mark.parametrize("kind,n", [("tri", 12), ("tri-unstructured", 9), ("tet", 4), ("tet-unstructured", 3)])
def test_elasticity_operator_matches_oracle(kind, n, kphi, with_bc):
    if kind.startswith("tri"):
        mesh = synthetic.rectangle_mesh(n, device="cuda")
        center, radius = (0.013, -0.021), 0.61
    else:
        mesh = synthetic.box_mesh(n, device="cuda")
        center, radius = synthetic.SPHERE_CENTER, 0.37
    if kind.endswith("unstructured"):
        mesh = synthetic.unstructured_variant(mesh, jitter=0.2, seed=7)
    d = mesh.gdim
    V1, Vp = fem.functionspace(mesh, 1), fem.functionspace(mesh, kphi)
    det = fem.Function(V1, synthetic.sphere_levelset(mesh.x, center=center, radius=radius).cpu().numpy())
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", RuntimeWarning)
        ctags, ftags, _, d_bdry, _ = mesh_scripts.compute_tags_measures(mesh, det, 1, box_mode=True)
    phi = synthetic.sphere_levelset(Vp.dof_coordinates_dev(), center=center, radius=radius)
    rng = np.random.default_rng(77)
    f = torch.from_numpy(rng.uniform(-1, 1, (mesh.num_vertices, d))).cuda()
    plan = elasticity.build_plan_interface_elasticity(mesh, ctags, ftags, d_bdry, V_phi=Vp)
    assert plan.facets_in.numel() > 0 and plan.facets_out.numel() > 0
    assert plan.entities_in.shape[0] > 0 and plan.entities_out.shape[0] > 0
    mat = elasticity.Material(1.0, 0.3, 0.05, 0.27)
    bcs = bc_dofs = bc_vals = None
    if with_bc:
        bc_dofs = plan.dofs("u_in", plan.boundary_vertices()).reshape(-1)
        bc_vals = torch.from_numpy(rng.uniform(-1, 1, bc_dofs.numel())).cuda()
        bcs = (bc_dofs, bc_vals)
        bc_dofs, bc_vals = bc_dofs.cpu().numpy(), bc_vals.cpu().numpy()
    A, b = elasticity.assemble_interface_elasticity(plan, phi, f, mat, pen_coef=1.3, stab_coef=0.7, bcs=bcs)
    if with_bc:   # the full-matrix Dirichlet pass gives the same system (the default, list-driven one sums its lifting
        # in another order)
        A2, b2 = elasticity.assemble_interface_elasticity(plan, phi, f, mat, pen_coef=1.3, stab_coef=0.7, bcs=bcs,
                                                          symmetric_bc=False)
        assert torch.allclose(A2.data, A.data, rtol=0, atol=1e-13 * float(A.data.abs().max()))
        assert torch.allclose(b2, b, rtol=0, atol=1e-12 * float(b.abs().max()))
    assert A.shape[0] == plan.nb * mesh.num_vertices
    omat = OE.Material(1.0, 0.3, 0.05, 0.27)
    ip, ix, data, bo = OE.assemble_interface_elasticity(
        mesh.x.cpu().numpy(), mesh.cells.cpu().numpy().astype(np.int64), phi.cpu().numpy(), f.cpu().numpy(),
        ctags.values_dev.cpu().numpy(), ftags.values_dev.cpu().numpy(), mesh.c2f.cpu().numpy(), mesh.f2c.cpu().numpy(),
        d_bdry(100).integration_entities, d_bdry(101).integration_entities, mat=omat, gamma=1.3, sigma_s=0.7,
        method="closed_form" if kphi == 1 else "quadrature", kphi=kphi, phi_dofmap=Vp.dofmap.astype(np.int64),
        bc_dofs=bc_dofs, bc_values=bc_vals, nquad=4)
    assert np.array_equal(A.indptr.cpu().numpy(), ip) and np.array_equal(A.indices.cpu().numpy(), ix)
    scale, rows = _row_scale(ip, data)
    gs = np.abs(data).max()
    assert np.all(np.abs(A.data.cpu().numpy() - data) <= RTOL * np.maximum(scale[rows], 1e-300 * gs))
    assert np.all(np.abs(b.cpu().numpy() - bo) <= RTOL * np.abs(bo).max())


def _exact(X, mat):
    """data.py:43-48 (both components equal)."""
    r = np.sqrt(X[:, 0] ** 2 + X[:, 1] ** 2)
    val = np.cos(r) - np.cos(1.0) / mat.E_in
    val = np.where(r < 1.0, val * (mat.E_in / mat.E_out), val)
    return np.stack([val, val], axis=1)


def _source(X, mat):
    """f = -div(sigma_in(cos_vec(x))) / E_in (main.py:148-150), u = (cos r, cos r):
    (div sigma)_i = (lmbda + mu) d_i(d_x c + d_y c) + mu lap c."""
    r = np.sqrt(X[:, 0] ** 2 + X[:, 1] ** 2)
    c1, c2 = -np.sin(r), -np.cos(r)                      # c', c''
    H = np.empty((len(X), 2, 2))
    for i in range(2):
        for j in range(2):
            H[:, i, j] = c2 * X[:, i] * X[:, j] / r ** 2 + c1 * ((i == j) / r - X[:, i] * X[:, j] / r ** 3)
    lap = c2 + c1 / r
    div = (mat.lmbda_in + mat.mu_in) * H.sum(axis=2) + mat.mu_in * lap[:, None]
    return -div / mat.E_in


def _interface_error(n, E_out):
    import scipy.sparse.linalg as spla
    from phifem_b200 import quadrature
    mat = elasticity.Material(1.0, 0.3, E_out, 0.3)
    mesh = synthetic.rectangle_mesh(n, lo=(-1.5, -1.5), hi=(1.5, 1.5), device="cuda")   # param1.yaml bbox
    X = mesh.x.cpu().numpy()
    V = fem.functionspace(mesh, 1)
    phi = 1.0 - (mesh.x[:, 0] ** 2 + mesh.x[:, 1] ** 2)                                  # data.py:39-40
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", RuntimeWarning)
        ctags, ftags, _, d_bdry, _ = mesh_scripts.compute_tags_measures(mesh, fem.Function(V, phi), 1, box_mode=True)
    plan = elasticity.build_plan_interface_elasticity(mesh, ctags, ftags, d_bdry)
    bv = plan.boundary_vertices()
    ue = _exact(X, mat)
    bcs = (plan.dofs("u_in", bv).reshape(-1), torch.from_numpy(ue[bv.cpu().numpy()].reshape(-1)).cuda())
    A, b = elasticity.assemble_interface_elasticity(plan, phi, _source(X, mat), mat, pen_coef=1.0, stab_coef=1.0, bcs=bcs)
    M = A.to_scipy().tocsr()
    keep = np.nonzero(np.asarray(abs(M).sum(axis=1)).ravel() > 0)[0]
    sol = np.zeros(M.shape[0])
    sol[keep] = spla.spsolve(M[keep][:, keep].tocsc(), b.cpu().numpy()[keep])
    u_in, u_out, *_ = plan.split(torch.from_numpy(sol))
    lam, wq = quadrature.simplex_rule(2, 4)
    tags = ctags.values_dev.cpu().numpy()
    err2 = nrm2 = 0.0
    for tag, uh in ((1, u_in.numpy()), (3, u_out.numpy())):      # phi < 0: material "in"; phi > 0: material "out"
        cells = mesh.cells.cpu().numpy()[tags == tag]
        xc = X[cells]
        xq = np.einsum("qv,mvd->qmd", lam, xc).reshape(-1, 2)
        ueq = _exact(xq, mat).reshape(len(wq), len(cells), 2)
        uhq = np.einsum("qv,mvd->qmd", lam, uh[cells])
        e = xc[:, 1:] - xc[:, :1]
        area = 0.5 * np.abs(e[:, 0, 0] * e[:, 1, 1] - e[:, 0, 1] * e[:, 1, 0])
        err2 += float((wq[:, None, None] * (uhq - ueq) ** 2 * area[None, :, None]).sum())
        nrm2 += float((wq[:, None, None] * ueq ** 2 * area[None, :, None]).sum())
    return (err2 / nrm2) ** 0.5


def test_interface_elasticity_manufactured_solution_converges():
    """The demo's own test case (param1.yaml: bbox [-1.5, 1.5]^2, unit-disc interface) with a milder contrast
    (E_out = 0.1) and the demo's (E_out = 0.001): relative L2 error of (u_in on interior cells, u_out on exterior cells)."""
    for E_out, bound in ((0.1, 2e-3), (0.001, 3e-3)):
        errs = [_interface_error(n, E_out) for n in (15, 31, 63)]
        rates = [np.log2(errs[i] / errs[i + 1]) for i in range(2)]
        print("interface-elasticity errors", E_out, errs, rates)
        assert errs[-1] < bound and min(rates) > 1.7, (E_out, errs, rates)


def test_interface_elasticity_demo_runs_end_to_end():
    """demo/interface_elasticity.py = `python main.py param1` of the reference demo (mesh size 0.2, uniform refinement;
    n = 30 has a vertex at the origin, where the analytic source needs its limit)."""
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "demo"))
    import interface_elasticity
    res = interface_elasticity.main(mesh_size=0.2, iterations=3, quiet=True)
    l2, h1 = res["L2 relative error"], res["H10 relative error"]
    assert res["dof"] == [2 * 16 * 16, 2 * 31 * 31, 2 * 61 * 61]
    assert l2[-1] < 3e-3 and l2[0] / l2[-1] > 10.0, l2
    assert h1[-1] < 6e-2 and h1[0] / h1[-1] > 3.0, h1


@pytest.mark.parametrize("where", ["outside", "inside"])
def test_elasticity_without_cut_cells(where):
    """The interface misses the mesh: every cell carries one material only (tag 3 -> u_out stiffness, tag 1 -> u_in), no
    cut cell, no interface facet, no one-sided entity; the kernels must cope with the empty lists and match the oracle."""
    mesh = synthetic.unstructured_variant(synthetic.rectangle_mesh(7, device="cuda"), jitter=0.2, seed=3)
    r = 0.5 if where == "outside" else 9.0
    c = (5.0, 5.0) if where == "outside" else (0.0, 0.0)
    phi = synthetic.sphere_levelset(mesh.x, center=c, radius=r)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", RuntimeWarning)
        ctags, ftags, _, d_bdry, _ = mesh_scripts.compute_tags_measures(
            mesh, fem.Function(fem.functionspace(mesh, 1), phi), 1, box_mode=True)
    tag = 3 if where == "outside" else 1
    assert bool((ctags.values_dev == tag).all())
    plan = elasticity.build_plan_interface_elasticity(mesh, ctags, ftags, d_bdry)
    assert plan.cut_cells.numel() == 0 and plan.facets_in.numel() == 0 and plan.facets_out.numel() == 0
    f = torch.from_numpy(np.random.default_rng(5).uniform(-1, 1, (mesh.num_vertices, 2))).cuda()
    A, b = elasticity.assemble_interface_elasticity(plan, phi, f, elasticity.Material(1.0, 0.3, 0.05, 0.27))
    ip, ix, data, bo = OE.assemble_interface_elasticity(
        mesh.x.cpu().numpy(), mesh.cells.cpu().numpy().astype(np.int64), phi.cpu().numpy(), f.cpu().numpy(),
        ctags.values_dev.cpu().numpy(), ftags.values_dev.cpu().numpy(), mesh.c2f.cpu().numpy(), mesh.f2c.cpu().numpy(),
        d_bdry(100).integration_entities, d_bdry(101).integration_entities, mat=OE.Material(1.0, 0.3, 0.05, 0.27))
    assert np.array_equal(A.indptr.cpu().numpy(), ip) and np.array_equal(A.indices.cpu().numpy(), ix)
    assert np.abs(A.data.cpu().numpy() - data).max() <= RTOL * np.abs(data).max()
    assert np.abs(b.cpu().numpy() - bo).max() <= RTOL * np.abs(bo).max()
    # only the stiffness block of the one material is populated
    blk = np.repeat(np.arange(len(ip) - 1), np.diff(ip)) % plan.nb
    off = plan.layout["u_out" if where == "outside" else "u_in"][0]
    assert np.all(data[(blk < off) | (blk >= off + 2)] == 0.0) and np.abs(data).max() > 0


def test_list_driven_dirichlet_pass_on_a_large_pattern():
    """The default (list-driven) Dirichlet pass against the pass over the whole matrix on 180 000 triangles
    (1.26 M mixed dofs, 1.2e8 pattern entries): same matrix, same lifted load vector."""
    mesh = synthetic.rectangle_mesh(300, lo=(-1.5, -1.5), hi=(1.5, 1.5), device="cuda")
    V1 = fem.functionspace(mesh, 1)
    phi = synthetic.sphere_levelset(mesh.x, center=(0.0, 0.0), radius=1.0)
    det = fem.Function(V1, phi.cpu().numpy())
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", RuntimeWarning)
        ctags, ftags, _, d_bdry, _ = mesh_scripts.compute_tags_measures(mesh, det, 1, box_mode=True)
    plan = elasticity.build_plan_interface_elasticity(mesh, ctags, ftags, d_bdry, V_phi=V1)
    rng = np.random.default_rng(5)
    f = torch.from_numpy(rng.uniform(-1, 1, (mesh.num_vertices, 2))).cuda()
    bc_dofs = plan.dofs("u_out", plan.boundary_vertices()).reshape(-1)
    bcs = (bc_dofs, torch.from_numpy(rng.uniform(-1, 1, bc_dofs.numel())).cuda())
    mat = elasticity.Material(1.0, 0.3, 0.05, 0.27)
    A1, b1 = elasticity.assemble_interface_elasticity(plan, phi, f, mat, bcs=bcs)
    A2, b2 = elasticity.assemble_interface_elasticity(plan, phi, f, mat, bcs=bcs, symmetric_bc=False)
    assert plan.nnz > 1.0e8
    # (the cut-cell kernels accumulate with fp64 reductions: two assemblies differ in the last bits)
    assert torch.allclose(A1.data, A2.data, rtol=0, atol=1e-13 * float(A2.data.abs().max()))
    free = torch.ones(plan.n_rows, dtype=torch.bool, device="cuda")
    free[bc_dofs.long()] = False
    rows = torch.repeat_interleave(torch.arange(plan.n_rows, device="cuda"), (plan.indptr[1:] - plan.indptr[:-1]).long())
    cols = plan.indices.long()
    assert bool((A1.data[~free[rows] | ~free[cols]] == (rows == cols)[~free[rows] | ~free[cols]].double()).all())
    assert torch.allclose(b1, b2, rtol=0, atol=1e-12 * float(b2.abs().max()))
    assert bool((b1[bc_dofs.long()] == bcs[1]).all())
