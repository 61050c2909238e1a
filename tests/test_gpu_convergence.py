"""Semantic check of the GPU-assembled operator: solve the strong-Dirichlet phi-FEM system of reference
demo/strong-dirichlet/flower/main.py:104-165 for a manufactured solution and measure the error.

The reference holds no golden matrix (the CSR parity of the oracle is unpinned), so besides the
oracle-vs-CUDA comparisons this test pins what the operator MEANS: with f = -lap(u) and u = 0 on
{phi = 0}, u_h = phi_h w_h must converge to u at the rate of the element (O(h^2) in L2 for P1, O(h^3) for
P2).  The linear solve is test infrastructure (scipy on the host, rows without entries dropped the way MUMPS
ICNTL(24) does at main.py:147-148); assembly and tags run on the GPU through the public API."""
import warnings

import numpy as np
import pytest
import torch

from phifem_b200 import assemble, fem, mesh_scripts, quadrature, synthetic

pytestmark = pytest.mark.gpu
R = 0.62
CENTER = (0.013, -0.021)


def _exact(X):
    """u = phi * w with phi = |x - c|^2 - R^2, w = -exp(x0) cos(x1): u = 0 on the circle."""
    x0, x1 = X[:, 0] - CENTER[0], X[:, 1] - CENTER[1]
    phi = x0 * x0 + x1 * x1 - R * R
    w = -torch.exp(x0) * torch.cos(x1)
    # lap(phi w) = w lap(phi) + 2 grad(phi).grad(w) + phi lap(w); lap(w) = 0 for exp(x0) cos(x1)
    gw0, gw1 = w, torch.exp(x0) * torch.sin(x1)
    lap = 4.0 * w + 2.0 * (2.0 * x0 * gw0 + 2.0 * x1 * gw1)
    return phi * w, -lap


def _solve_and_error(n, k):
    import scipy.sparse.linalg as spla
    mesh = synthetic.rectangle_mesh(n, device="cuda")
    V, V1 = fem.functionspace(mesh, k), fem.functionspace(mesh, 1)
    det = fem.Function(V1, synthetic.sphere_levelset(mesh.x, center=CENTER, radius=R).cpu().numpy())
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", RuntimeWarning)
        ctags, ftags, _, ds, _ = mesh_scripts.compute_tags_measures(mesh, det, 1, box_mode=True)
    Xd = V.dof_coordinates_dev()
    phi = synthetic.sphere_levelset(Xd, center=CENTER, radius=R)
    _, f = _exact(Xd)
    plan = assemble.build_plan(mesh, ctags, ftags, ds(100), V=V, V_phi=V)
    A, b = assemble.assemble_strong_dirichlet(plan, phi, f, stab_coef=1.0)
    M = A.to_scipy().tocsr()
    keep = np.nonzero(np.diff(M.indptr) > 0)[0]
    w = np.zeros(M.shape[0])
    w[keep] = spla.spsolve(M[keep][:, keep].tocsc(), b.cpu().numpy()[keep])
    # L2 error of u_h = phi_h w_h over the cells tagged 1 (interior), by quadrature of degree 8
    lam, wq = quadrature.simplex_rule(2, 8)
    tags = ctags.values_dev
    cells = torch.nonzero(tags == 1).reshape(-1)
    dm = V.dofmap_dev[cells].long().cpu().numpy()
    xc = mesh.x[mesh.cells[cells].long()].cpu().numpy()                     # [m, 3, 2]
    if k == 1:
        basis = lam
    else:
        edges = fem.LOCAL_EDGES["triangle"]
        basis = np.concatenate([lam * (2 * lam - 1)] + [4 * lam[:, [a]] * lam[:, [b]] for a, b in edges], axis=1)
    ph = phi.cpu().numpy()
    uh = (basis @ ph[dm].T) * (basis @ w[dm].T)                               # [nq, m]
    xq = np.einsum("qv,mvd->qmd", lam, xc)
    ue, _ = _exact(torch.from_numpy(xq.reshape(-1, 2)))
    ue = ue.numpy().reshape(uh.shape)
    e = xc[:, 1:] - xc[:, :1]
    area = 0.5 * np.abs(e[:, 0, 0] * e[:, 1, 1] - e[:, 0, 1] * e[:, 1, 0])
    err2 = float((wq[:, None] * (uh - ue) ** 2 * area[None, :]).sum())
    nrm2 = float((wq[:, None] * ue ** 2 * area[None, :]).sum())
    return (err2 / nrm2) ** 0.5


@pytest.mark.parametrize("k,sizes,rate", [(1, (24, 48, 96), 1.7), (2, (12, 24, 48), 2.6)])
def test_manufactured_solution_converges(k, sizes, rate):
    errs = [_solve_and_error(n, k) for n in sizes]
    rates = [np.log2(errs[i] / errs[i + 1]) for i in range(len(errs) - 1)]
    assert errs[-1] < (2e-3 if k == 1 else 5e-5), errs
    assert min(rates) > rate, (errs, rates)
