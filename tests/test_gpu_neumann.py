"""GPU parity of the Neumann phi-FEM operator (reference demo/neumann/square/main.py:103-161, BASELINE.json configs[1])
through the C ABI against the oracle (closed forms for a P1 level set, brute-force quadrature for P2), and its meaning:
the solved system converges to a manufactured solution of -lap u + u = f, du/dn = u_N."""
import warnings

import numpy as np
import pytest
import torch

from oracle import assembly as OA
from phifem_b200 import assemble, fem, mesh_scripts, quadrature, synthetic

pytestmark = pytest.mark.gpu
RTOL = 1e-12


def _row_scale(indptr, data):
    scale = np.zeros(len(indptr) - 1)
    rows = np.repeat(np.arange(len(indptr) - 1), np.diff(indptr))
    np.maximum.at(scale, rows, np.abs(data))
    return scale, rows


@pytest.mark.parametrize("kphi,robin", [(1, 0.0), (2, 0.0), (1, 0.8), (2, 0.8)])
@pytest.mark.parametrize("kind,n", [("tri", 14), ("tri-unstructured", 10), ("tet", 5), ("tet-unstructured", 4)])
def test_neumann_operator_matches_oracle(kind, n, kphi, robin):
    """robin = 0: demo/neumann (gradient jumps over dS(3)); robin != 0: demo/robin (over dS(2))."""
    gtag = 2 if robin else 3
    if kind.startswith("tri"):
        mesh = synthetic.rectangle_mesh(n, device="cuda")
        center, radius = (0.013, -0.021), 0.61
    else:
        mesh = synthetic.box_mesh(n, device="cuda")
        center, radius = synthetic.SPHERE_CENTER, 0.37
    if kind.endswith("unstructured"):
        mesh = synthetic.unstructured_variant(mesh, jitter=0.2, seed=7)
    V1, Vp = fem.functionspace(mesh, 1), fem.functionspace(mesh, kphi)
    det = fem.Function(V1, synthetic.sphere_levelset(mesh.x, center=center, radius=radius).cpu().numpy())
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", RuntimeWarning)
        ctags, ftags, _, ds, _ = mesh_scripts.compute_tags_measures(mesh, det, 1, box_mode=True)
    phi = synthetic.sphere_levelset(Vp.dof_coordinates_dev(), center=center, radius=radius)
    rng = np.random.default_rng(77)
    f = torch.from_numpy(rng.uniform(-1, 1, mesh.num_vertices)).cuda()
    un = torch.from_numpy(rng.uniform(-1, 1, mesh.num_vertices)).cuda()
    plan = assemble.build_plan_neumann(mesh, ctags, ftags, ds(100), V_phi=Vp, ghost_tag=gtag)
    A, b = assemble.assemble_neumann(plan, phi, f, un, pen_coef=1.3, stab_coef=0.7, robin_coef=robin)
    d = mesh.gdim
    assert A.shape[0] == (d + 1) * mesh.num_vertices + mesh.num_cells
    assert plan.ghost.numel() > 0 and plan.entities.shape[0] > 0
    x = mesh.x.cpu().numpy()
    cells = mesh.cells.cpu().numpy().astype(np.int64)
    ip, ix, data, bo = OA.assemble_neumann(
        x, cells, phi.cpu().numpy(), f.cpu().numpy(), un.cpu().numpy(), ctags.values_dev.cpu().numpy(),
        ftags.values_dev.cpu().numpy(), mesh.c2f.cpu().numpy(), mesh.f2c.cpu().numpy(),
        ds(100).integration_entities, gamma=1.3, sigma=0.7,
        method="closed_form" if kphi == 1 else "quadrature", kphi=kphi, phi_dofmap=Vp.dofmap.astype(np.int64),
        # P2 level set: |grad phi_h| in the load term is not polynomial, the value depends on the rule (dolfinx would
        # use the rule of UFL's estimated degree): compare on the kernel's rule
        rule=quadrature.rules_for_neumann(d, 2)[0] if kphi == 2 else None, robin_coef=robin, ghost_tag=gtag)
    assert np.array_equal(A.indptr.cpu().numpy(), ip) and np.array_equal(A.indices.cpu().numpy(), ix)
    scale, rows = _row_scale(ip, data)
    gs = np.abs(data).max()
    assert np.all(np.abs(A.data.cpu().numpy() - data) <= RTOL * np.maximum(scale[rows], 1e-300 * gs))
    assert np.all(np.abs(b.cpu().numpy() - bo) <= RTOL * np.abs(bo).max())
    u, y, p = plan.split(b)
    assert u.numel() == mesh.num_vertices and y.shape == (mesh.num_vertices, d) and p.numel() == mesh.num_cells


def _neumann_error(n, kphi):
    import scipy.sparse.linalg as spla
    R, C = 0.62, (0.013, -0.021)
    mesh = synthetic.rectangle_mesh(n, device="cuda")
    X = mesh.x
    x0, x1 = X[:, 0] - C[0], X[:, 1] - C[1]
    phi = x0 * x0 + x1 * x1 - R * R
    # u = cos(1.3 x0) exp(0.5 x1): -lap u + u = (1 + 1.69 - 0.25) u;  du/dn with n = grad(phi) / |grad(phi)|
    u = torch.cos(1.3 * x0) * torch.exp(0.5 * x1)
    f = (1.0 + 1.69 - 0.25) * u
    gx, gy = -1.3 * torch.sin(1.3 * x0) * torch.exp(0.5 * x1), 0.5 * u
    r = torch.sqrt(x0 * x0 + x1 * x1).clamp(min=1e-12)
    un = (gx * x0 + gy * x1) / r
    V = fem.functionspace(mesh, 1)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", RuntimeWarning)
        ctags, ftags, _, ds, _ = mesh_scripts.compute_tags_measures(mesh, fem.Function(V, phi), 1, box_mode=True)
    # level set one degree above u, as in the demo (levelset_degree = 2, main.py:43): with a P1 level set the
    # discrete normal grad(phi_h) / |grad(phi_h)| is only first-order accurate and so is the solution
    Vp = fem.functionspace(mesh, kphi)
    phi_a = synthetic.sphere_levelset(Vp.dof_coordinates_dev(), center=C, radius=R)
    plan = assemble.build_plan_neumann(mesh, ctags, ftags, ds(100), V_phi=Vp)
    A, b = assemble.assemble_neumann(plan, phi_a, f, un, pen_coef=1.0, stab_coef=1.0)
    M = A.to_scipy().tocsr()
    keep = np.nonzero(np.asarray(abs(M).sum(axis=1)).ravel() > 0)[0]
    sol = np.zeros(M.shape[0])
    sol[keep] = spla.spsolve(M[keep][:, keep].tocsc(), b.cpu().numpy()[keep])
    uh = plan.split(torch.from_numpy(sol))[0].numpy()
    lam, wq = quadrature.simplex_rule(2, 4)
    cells = mesh.cells[torch.nonzero(ctags.values_dev == 1).reshape(-1)].long().cpu().numpy()
    xc = X.cpu().numpy()[cells]
    xq = np.einsum("qv,mvd->qmd", lam, xc)
    ue = np.cos(1.3 * (xq[..., 0] - C[0])) * np.exp(0.5 * (xq[..., 1] - C[1]))
    e = xc[:, 1:] - xc[:, :1]
    area = 0.5 * np.abs(e[:, 0, 0] * e[:, 1, 1] - e[:, 0, 1] * e[:, 1, 0])
    err2 = float((wq[:, None] * (lam @ uh[cells].T - ue) ** 2 * area[None]).sum())
    nrm2 = float((wq[:, None] * ue ** 2 * area[None]).sum())
    return (err2 / nrm2) ** 0.5


@pytest.mark.parametrize("kphi,min_rate", [(2, 1.7), (1, 1.0)])
def test_neumann_manufactured_solution_converges(kphi, min_rate):
    errs = [_neumann_error(n, kphi) for n in (24, 48, 96)]
    rates = [np.log2(errs[i] / errs[i + 1]) for i in range(2)]
    print("neumann errors", kphi, errs, rates)
    assert errs[-1] < 2e-3 and min(rates) > min_rate, (errs, rates)


def test_neumann_without_cut_cells():
    """Every cell interior (the level set contains the mesh): only the closed-form u-u block and the load of the
    interior kernel remain (plus the one-sided term on the mesh boundary); empty cut / ghost lists must be handled."""
    mesh = synthetic.unstructured_variant(synthetic.rectangle_mesh(7, device="cuda"), jitter=0.2, seed=3)
    phi = synthetic.sphere_levelset(mesh.x, center=(0.0, 0.0), radius=9.0)
    V = fem.functionspace(mesh, 1)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", RuntimeWarning)
        ctags, ftags, _, ds, _ = mesh_scripts.compute_tags_measures(mesh, fem.Function(V, phi), 1, box_mode=True)
    assert bool((ctags.values_dev == 1).all())
    rng = np.random.default_rng(2)
    f = torch.from_numpy(rng.uniform(-1, 1, mesh.num_vertices)).cuda()
    un = torch.from_numpy(rng.uniform(-1, 1, mesh.num_vertices)).cuda()
    plan = assemble.build_plan_neumann(mesh, ctags, ftags, ds(100))
    # no exterior cell: the mesh-boundary facets become Gamma_h (src/phifem/mesh_scripts.py:469-474), so ds(100) is not empty
    assert plan.cut_positions.numel() == 0 and plan.ghost.numel() == 0 and plan.entities.shape[0] == 4 * 7
    A, b = assemble.assemble_neumann(plan, phi, f, un)
    ip, ix, data, bo = OA.assemble_neumann(
        mesh.x.cpu().numpy(), mesh.cells.cpu().numpy().astype(np.int64), phi.cpu().numpy(), f.cpu().numpy(),
        un.cpu().numpy(), ctags.values_dev.cpu().numpy(), ftags.values_dev.cpu().numpy(), mesh.c2f.cpu().numpy(),
        mesh.f2c.cpu().numpy(), ds(100).integration_entities)
    assert np.array_equal(A.indptr.cpu().numpy(), ip) and np.array_equal(A.indices.cpu().numpy(), ix)
    assert np.abs(A.data.cpu().numpy() - data).max() <= RTOL * np.abs(data).max()
    assert np.abs(b.cpu().numpy() - bo).max() <= RTOL * np.abs(bo).max()
