"""CPU: the two restatements of the weak-Dirichlet operator (closed forms vs brute-force quadrature) agree, and
the symbolic phase of the mixed space reproduces the oracle's pattern."""
import numpy as np
import pytest
import torch

from oracle import assembly as OA
from oracle import tags as OT
from phifem_b200 import fem, synthetic
from phifem_b200.assemble_pk import PkAssemblyPlan


def _case(kind, n):
    mesh = synthetic.rectangle_mesh(n, device="cpu") if kind == "tri" else synthetic.box_mesh(n, device="cpu")
    mesh = synthetic.unstructured_variant(mesh, jitter=0.15, seed=5)
    x = mesh.x_host
    cells = mesh.cells_host.astype(np.int64)
    center = (0.02, -0.03) if kind == "tri" else synthetic.SPHERE_CENTER
    ph = synthetic.sphere_levelset(mesh.x, center=center, radius=0.6 if kind == "tri" else 0.37).numpy()
    ct = mesh.cell_type
    pts = OT.cell_detection_points(ct, 1)
    fpts = OT.facet_points_in_cell(ct, 1)
    ftab = np.asarray([OT.coordinate_basis(ct, p)[0] for p in fpts])
    out = OT.compute_tags_measures(x, cells, ct, ph[cells], OT.point_values_function(ph, cells, ftab),
                                   box_mode=True, detection_points=pts)
    return mesh, x, cells, ph, out


@pytest.mark.parametrize("kind,n", [("tri", 8), ("tet", 3)])
def test_weak_closed_form_equals_quadrature(kind, n):
    mesh, x, cells, ph, out = _case(kind, n)
    rng = np.random.default_rng(0)
    f, ud = rng.uniform(-1, 1, len(x)), rng.uniform(-1, 1, len(x))
    args = (x, cells, cells, len(x), ph, f, ud, out["cell_tags"], out["facet_tags"], out["c2f"], out["f2c"],
            out["ds100"])
    a = OA.assemble_weak_dirichlet(*args, gamma=1.3, sigma=0.7)
    q = OA.assemble_weak_dirichlet(*args, gamma=1.3, sigma=0.7, method="quadrature")
    assert np.array_equal(a[0], q[0]) and np.array_equal(a[1], q[1])
    assert np.abs(a[2] - q[2]).max() <= 1e-13 * np.abs(a[2]).max()
    assert np.abs(a[3] - q[3]).max() <= 1e-13 * np.abs(a[3]).max()
    # the uu block of interior cells is the plain stiffness matrix: constants in its kernel, symmetric cell part
    n2 = 2 * len(x)
    M = OA.to_scipy(a[0], a[1], a[2], n2)
    assert abs(M[1::2][:, 0::2] - M[0::2][:, 1::2].T).max() <= 1e-13 * abs(M).max()   # up = pu^T


@pytest.mark.parametrize("kind,n,k", [("tri", 8, 1), ("tri", 6, 2), ("tet", 3, 1)])
def test_weak_symbolic_phase_matches_oracle_pattern(kind, n, k):
    mesh, x, cells, ph, out = _case(kind, n)
    V = fem.functionspace(mesh, k)
    ct8 = torch.from_numpy(out["cell_tags"].astype(np.int8))
    ft8 = torch.from_numpy(out["facet_tags"].astype(np.int8))
    ents = torch.from_numpy(np.asarray(out["ds100"], dtype=np.int32))
    plan = PkAssemblyPlan(mesh, ct8, ft8, ents, V, V, form="weak")
    mixed = np.concatenate([2 * V.dofmap.astype(np.int64), 2 * V.dofmap.astype(np.int64) + 1], axis=1)
    active = np.nonzero((out["cell_tags"] == 1) | (out["cell_tags"] == 2))[0]
    ghost = np.nonzero(((out["facet_tags"] == 2) | (out["facet_tags"] == 3)) & (out["f2c"][:, 1] >= 0))[0]
    ip, ix = OA.sparsity_pattern(2 * V.num_dofs, mixed, active, ghost, out["f2c"])
    assert plan.n_rows == 2 * V.num_dofs
    assert np.array_equal(plan.indptr.numpy(), ip) and np.array_equal(plan.indices.numpy(), ix)
    nm = mixed.shape[1]
    rows = np.repeat(np.arange(2 * V.num_dofs), np.diff(ip))
    sl = plan.slots_cells.numpy().reshape(nm, nm, len(active))
    for a in (0, nm // 2, nm - 1):
        for b in (1, nm // 2 + 1):
            assert np.array_equal(rows[sl[a, b]], mixed[active][:, a])
            assert np.array_equal(ix[sl[a, b]], mixed[active][:, b])


@pytest.mark.parametrize("kind,n", [("tri", 8), ("tet", 3)])
def test_neumann_closed_form_equals_quadrature_and_symbolic_phase(kind, n):
    """Neumann operator (demo/neumann/square/main.py:103-158): the two restatements agree; the product's symbolic
    phase for the mixed space (u, y, p) reproduces the oracle's pattern and numbering."""
    mesh, x, cells, ph, out = _case(kind, n)
    rng = np.random.default_rng(0)
    f, un = rng.uniform(-1, 1, len(x)), rng.uniform(-1, 1, len(x))
    args = (x, cells, ph, f, un, out["cell_tags"], out["facet_tags"], out["c2f"], out["f2c"], out["ds100"])
    a = OA.assemble_neumann(*args, gamma=1.3, sigma=0.7)
    q = OA.assemble_neumann(*args, gamma=1.3, sigma=0.7, method="quadrature")
    assert np.array_equal(a[0], q[0]) and np.array_equal(a[1], q[1])
    assert np.abs(a[2] - q[2]).max() <= 1e-13 * np.abs(a[2]).max()
    assert np.abs(a[3] - q[3]).max() <= 1e-13 * np.abs(a[3]).max()
    d = x.shape[1]
    nrows = (d + 1) * len(x) + len(cells)
    M = OA.to_scipy(a[0], a[1], a[2], nrows)
    # the cell part of the form is symmetric; only int_ds (y.n) v is not: remove it and compare
    Ab = OA.neumann_boundary_tensors(x, cells, out["ds100"])
    mixed = OA.neumann_mixed_dofmap(cells, len(x))
    import scipy.sparse as sp
    ents = np.asarray(out["ds100"]).reshape(-1, 2)
    dmb = mixed[ents[:, 0]]
    nm = mixed.shape[1]
    B = sp.coo_matrix((Ab.ravel(), (np.repeat(dmb, nm, axis=1).ravel(), np.tile(dmb, (1, nm)).ravel())),
                      shape=(nrows, nrows)).tocsr()
    S = M - B
    assert abs(S - S.T).max() <= 1e-13 * abs(M).max()
    V = fem.functionspace(mesh, 1)
    ct8 = torch.from_numpy(out["cell_tags"].astype(np.int8))
    ft8 = torch.from_numpy(out["facet_tags"].astype(np.int8))
    ents_t = torch.from_numpy(np.asarray(out["ds100"], dtype=np.int32))
    plan = PkAssemblyPlan(mesh, ct8, ft8, ents_t, V, V, form="neumann")
    assert plan.n_rows == nrows and np.array_equal(plan.pattern_dofmap.numpy(), mixed)
    assert np.array_equal(plan.indptr.numpy(), a[0]) and np.array_equal(plan.indices.numpy(), a[1])
    assert np.array_equal(plan.ghost.numpy(),
                          np.nonzero((out["facet_tags"] == 3) & (out["f2c"][:, 1] >= 0))[0])
