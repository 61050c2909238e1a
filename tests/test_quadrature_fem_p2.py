"""CPU: quadrature tables, the torch-built P2 dofmap and the P_k symbolic phase (pattern + slot maps)."""
import math

import numpy as np
import pytest
import torch

from oracle import assembly as OA
from oracle import tags as OT
from phifem_b200 import fem, quadrature, synthetic
from phifem_b200.assemble_pk import PkAssemblyPlan


@pytest.mark.parametrize("d", [1, 2, 3])
@pytest.mark.parametrize("degree", range(1, 9))
def test_rules_integrate_monomials_exactly(d, degree):
    lam, w = quadrature.simplex_rule(d, degree)
    assert lam.shape == (len(w), d + 1) and np.all(w > 0) and np.all(lam > -1e-15)
    assert np.allclose(lam.sum(axis=1), 1.0, atol=1e-15)
    assert quadrature.max_moment_error(lam, w, degree) < 1e-15


def test_rule_sizes_and_oracle_agreement():
    assert len(quadrature.simplex_rule(2, 6)[1]) == 12      # Dunavant
    assert len(quadrature.simplex_rule(3, 6)[1]) == 24      # Keast
    (cl, cw), (fl, fw) = quadrature.rules_for(3, 2, 2)
    assert len(cw) == 24 and len(fw) == 16
    # the same polynomial integrated with the oracle's own Gauss-Jacobi rule (scipy roots_jacobi)
    lo, wo = OA.simplex_rule(3, 6)
    poly = lambda l: l[:, 0] ** 3 * l[:, 1] ** 2 * l[:, 3] + l[:, 2] ** 6
    assert abs((cw * poly(cl)).sum() - (wo * poly(lo)).sum() * math.factorial(3)) < 1e-16


@pytest.mark.parametrize("kind,n", [("tri", 7), ("tet", 4)])
def test_device_dofmap_matches_lexicographic_edge_numbering(kind, n):
    mesh = synthetic.rectangle_mesh(n, device="cpu") if kind == "tri" else synthetic.box_mesh(n, device="cpu")
    mesh = synthetic.unstructured_variant(mesh, jitter=0.1, seed=3)
    V = fem.functionspace(mesh, 2)
    cells = mesh.cells_host.astype(np.int64)
    edges = fem.LOCAL_EDGES[mesh.cell_type]
    pairs = np.sort(np.stack([cells[:, list(e)] for e in edges], axis=1), axis=2)
    uniq, inv = np.unique(pairs.reshape(-1, 2), axis=0, return_inverse=True)
    want = np.concatenate([cells, mesh.num_vertices + inv.reshape(len(cells), len(edges))], axis=1)
    assert np.array_equal(V.dofmap, want) and V.num_dofs == mesh.num_vertices + len(uniq)
    assert V.dofmap.dtype == np.int32
    X = V.dof_coordinates_dev().numpy()
    nodes = V.element.nodes                                   # reference nodes, same local order
    shape = OT.coordinate_basis(mesh.cell_type, nodes)[0]
    xc = mesh.x_host[cells]
    for i in range(V.element.ndofs):
        assert np.allclose(X[V.dofmap[:, i]], np.einsum("v,cvd->cd", shape[i], xc), atol=1e-15)
    # interpolation of a quadratic is exact in P2
    fn = fem.Function(V).interpolate(lambda x: x[0] ** 2 - 3 * x[0] * x[1] + 1)
    assert np.allclose(fn.x.array, X[:, 0] ** 2 - 3 * X[:, 0] * X[:, 1] + 1)


@pytest.mark.parametrize("kind,n,kw", [("tri", 10, 2), ("tet", 4, 2), ("tet", 4, 1)])
def test_pk_symbolic_phase_matches_oracle_pattern(kind, n, kw):
    mesh = synthetic.rectangle_mesh(n, device="cpu") if kind == "tri" else synthetic.box_mesh(n, device="cpu")
    mesh = synthetic.unstructured_variant(mesh, jitter=0.1, seed=5)
    V = fem.functionspace(mesh, kw)
    x = mesh.x_host
    cells = mesh.cells_host.astype(np.int64)
    center = (0.02, -0.03) if kind == "tri" else synthetic.SPHERE_CENTER
    ph = synthetic.sphere_levelset(mesh.x, center=center, radius=0.6 if kind == "tri" else 0.37).numpy()
    ct_name = mesh.cell_type
    pts = OT.cell_detection_points(ct_name, 1)
    fpts = OT.facet_points_in_cell(ct_name, 1)
    ftab = np.asarray([OT.coordinate_basis(ct_name, p)[0] for p in fpts])
    out = OT.compute_tags_measures(x, cells, ct_name, ph[cells], OT.point_values_function(ph, cells, ftab),
                                   box_mode=True, detection_points=pts)
    assert np.array_equal(out["c2f"], mesh.c2f.numpy())
    ct8 = torch.from_numpy(out["cell_tags"].astype(np.int8))
    ft8 = torch.from_numpy(out["facet_tags"].astype(np.int8))
    ents = torch.from_numpy(np.asarray(out["ds100"], dtype=np.int32))
    plan = PkAssemblyPlan(mesh, ct8, ft8, ents, V, V)
    active = np.nonzero((out["cell_tags"] == 1) | (out["cell_tags"] == 2))[0]
    ghost = np.nonzero(((out["facet_tags"] == 2) | (out["facet_tags"] == 3)) & (out["f2c"][:, 1] >= 0))[0]
    ip, ix = OA.sparsity_pattern(V.num_dofs, V.dofmap, active, ghost, out["f2c"])
    assert np.array_equal(plan.indptr.numpy(), ip) and np.array_equal(plan.indices.numpy(), ix)
    # slot maps: entry (i, j) of active cell e sits at the CSR position of (dof_i, dof_j)
    nd = V.dofmap.shape[1]
    rows = np.repeat(np.arange(V.num_dofs), np.diff(ip))
    sl = plan.slots_cells.numpy().reshape(nd, nd, len(active))
    dm = V.dofmap[active]
    for i in range(nd):
        for j in range(nd):
            assert np.array_equal(rows[sl[i, j]], dm[:, i]) and np.array_equal(ix[sl[i, j]], dm[:, j])
    mac = np.concatenate([V.dofmap[out["f2c"][ghost, 0]], V.dofmap[out["f2c"][ghost, 1]]], axis=1)
    sg = np.moveaxis(plan.slots_ghost.numpy().reshape(len(ghost), 2 * nd, 2 * nd), 0, 2)
    for a in range(0, 2 * nd, 3):
        for b in range(2 * nd):
            assert np.array_equal(rows[sg[a, b]], mac[:, a]) and np.array_equal(ix[sg[a, b]], mac[:, b])
    eb = plan.entities.numpy()
    sb = np.moveaxis(plan.slots_boundary.numpy().reshape(len(eb), nd, nd), 0, 2)
    dmb = V.dofmap[eb[:, 0]]
    assert len(eb) > 0 and np.array_equal(rows[sb[1, 2]], dmb[:, 1]) and np.array_equal(ix[sb[1, 2]], dmb[:, 2])
