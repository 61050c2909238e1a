"""CPU: the two restatements of the interface-elasticity operator (reference demo/interface-elasticity/main.py:152-274)
agree -- entry formulas from exact monomial integrals vs the UFL expressions evaluated field by field at brute-force
quadrature points --, the operator has the structure the forms imply, and the product's symbolic phase (vertex graph
with dense NB x NB blocks) reproduces the oracle's generic dof-pair pattern."""
import numpy as np
import pytest
import torch

from oracle import assembly as OA
from oracle import elasticity as OE
from phifem_b200 import elasticity, fem
from test_oracle_weak import _case


def _assemble(kind, n, method, **kw):
    mesh, x, cells, ph, out = _case(kind, n)
    rng = np.random.default_rng(0)
    f = rng.uniform(-1, 1, (len(x), x.shape[1]))
    mat = OE.Material(1.0, 0.3, 0.05, 0.27)
    res = OE.assemble_interface_elasticity(x, cells, ph, f, out["cell_tags"], out["facet_tags"], out["c2f"], out["f2c"],
                                           out["ds100"], out["ds101"], mat=mat, gamma=1.3, sigma_s=0.7, method=method,
                                           **kw)
    return mesh, x, cells, out, res


@pytest.mark.parametrize("kind,n", [("tri", 6), ("tet", 3)])
def test_elasticity_closed_form_equals_quadrature(kind, n):
    mesh, x, cells, out, a = _assemble(kind, n, "closed_form")
    _, _, _, _, q = _assemble(kind, n, "quadrature", nquad=3 if kind == "tet" else 6)
    assert set(np.unique(out["cell_tags"])) == {1, 2, 3}
    assert len(out["ds100"]) and len(out["ds101"])
    assert np.array_equal(a[0], q[0]) and np.array_equal(a[1], q[1])
    assert np.abs(a[2] - q[2]).max() <= 1e-13 * np.abs(a[2]).max()
    assert np.abs(a[3] - q[3]).max() <= 1e-13 * np.abs(a[3]).max()
    d = x.shape[1]
    o = OE.Offsets(d)
    nrows = o.nb * len(x)
    M = OA.to_scipy(a[0], a[1], a[2], nrows)
    # everything but int (y n).v is symmetric: the (u, y) blocks differ from their transposes by the boundary terms only
    blk = np.arange(nrows) % o.nb
    uu = np.nonzero(blk < 2 * d)[0]
    yy = np.nonzero((blk >= o.yi) & (blk < o.p))[0]
    for idx in (uu, yy):
        S = M[idx][:, idx]
        assert abs(S - S.T).max() <= 1e-13 * abs(M).max()
    # rigid translations of u_in are in the kernel of the u_in stiffness on vertices surrounded by interior cells
    interior_v = np.ones(len(x), dtype=bool)
    interior_v[np.unique(cells[out["cell_tags"] != 1])] = False
    if interior_v.any():
        t = np.zeros(nrows)
        t[o.nb * np.arange(len(x)) + o.ui] = 1.0
        r = M @ t
        rows = o.nb * np.nonzero(interior_v)[0][:, None] + np.arange(2 * d)[None, :]
        assert np.abs(r[rows]).max() <= 1e-12 * abs(M).max()
        # ... and so are the infinitesimal rotations (zero strain): u = (-y, x) in 2D, rotations about the axes in 3D
        pairs = [(0, 1)] if d == 2 else [(0, 1), (0, 2), (1, 2)]
        for (a_, b_) in pairs:
            t = np.zeros(nrows)
            t[o.nb * np.arange(len(x)) + o.ui + a_] = -x[:, b_]
            t[o.nb * np.arange(len(x)) + o.ui + b_] = x[:, a_]
            r = M @ t
            assert np.abs(r[rows]).max() <= 1e-12 * abs(M).max() * np.abs(x).max()
    # p only lives on cut cells: rows of p at vertices without a cut cell are empty of values
    cutv = np.zeros(len(x), dtype=bool)
    cutv[np.unique(cells[out["cell_tags"] == 2])] = True
    prow = o.nb * np.nonzero(~cutv)[0][:, None] + o.p + np.arange(d)[None, :]
    assert abs(M[prow.ravel()]).max() == 0.0


@pytest.mark.parametrize("kind,n", [("tri", 6), ("tet", 3)])
def test_elasticity_dirichlet_rows(kind, n):
    mesh, x, cells, out, a = _assemble(kind, n, "closed_form")
    d = x.shape[1]
    o = OE.Offsets(d)
    bv = np.unique(mesh.facet_vertices[mesh.boundary_facets.long()].numpy())
    dofs = (o.nb * bv[:, None] + o.ui + np.arange(d)[None, :]).ravel()
    g = np.random.default_rng(3).uniform(-1, 1, len(dofs))
    _, _, _, _, c = _assemble(kind, n, "closed_form", bc_dofs=dofs, bc_values=g)
    nrows = o.nb * len(x)
    M0, M1 = OA.to_scipy(a[0], a[1], a[2], nrows), OA.to_scipy(c[0], c[1], c[2], nrows)
    gg = np.zeros(nrows)
    gg[dofs] = g
    free = np.setdiff1d(np.arange(nrows), dofs)
    assert abs(M1[dofs][:, free]).max() == 0 and abs(M1[free][:, dofs]).max() == 0
    assert np.array_equal(M1[dofs][:, dofs].toarray(), np.eye(len(dofs)))
    assert abs(M1[free][:, free] - M0[free][:, free]).max() == 0
    assert np.allclose(c[3][free], (a[3] - M0 @ gg)[free], rtol=0, atol=1e-13 * np.abs(a[3]).max())
    assert np.array_equal(c[3][dofs], g)


@pytest.mark.parametrize("kind,n,kphi", [("tri", 6, 1), ("tri", 5, 2), ("tet", 3, 1)])
def test_elasticity_symbolic_phase_matches_oracle_pattern(kind, n, kphi):
    mesh, x, cells, ph, out = _case(kind, n)
    d = x.shape[1]
    o = OE.Offsets(d)
    ct8 = torch.from_numpy(out["cell_tags"].astype(np.int8))
    ft8 = torch.from_numpy(out["facet_tags"].astype(np.int8))
    e100 = torch.from_numpy(np.asarray(out["ds100"], dtype=np.int32))
    e101 = torch.from_numpy(np.asarray(out["ds101"], dtype=np.int32))
    plan = elasticity.ElasticityPlan(mesh, ct8, ft8, e100, e101, fem.functionspace(mesh, kphi))
    mixed = OE.mixed_dofmap(cells, d)
    interior = out["f2c"][:, 1] >= 0
    fac = np.nonzero(((out["facet_tags"] == 3) | (out["facet_tags"] == 4)) & interior)[0]
    ip, ix = OA.sparsity_pattern(o.nb * len(x), mixed, np.arange(len(cells)), fac, out["f2c"])
    assert plan.n_rows == o.nb * len(x) and plan.nb == o.nb
    assert np.array_equal(plan.indptr.numpy(), ip) and np.array_equal(plan.indices.numpy(), ix)
    assert plan.layout == {"u_in": (o.ui, d), "u_out": (o.uo, d), "y_in": (o.yi, d * d), "y_out": (o.yo, d * d),
                           "p": (o.p, d)}
    # the scalar slot map addresses the blocked CSR the way the kernels do
    vptr = plan.vptr.numpy().astype(np.int64)
    pos = plan.pos_cells.numpy().reshape(len(cells), d + 1, d + 1)
    rng = np.random.default_rng(1)
    for _ in range(50):
        c, k, j = rng.integers(len(cells)), rng.integers(d + 1), rng.integers(d + 1)
        a, b = rng.integers(o.nb), rng.integers(o.nb)
        r = cells[c, k]
        addr = o.nb * (o.nb * vptr[r] + a * (vptr[r + 1] - vptr[r]) + pos[c, k, j]) + b
        assert ip[o.nb * r + a] <= addr < ip[o.nb * r + a + 1] and ix[addr] == o.nb * cells[c, j] + b
    assert np.array_equal(np.sort(plan.facets_in.numpy()), np.nonzero((out["facet_tags"] == 3) & interior)[0])
    assert np.array_equal(np.sort(plan.facets_out.numpy()), np.nonzero((out["facet_tags"] == 4) & interior)[0])
    u = plan.split(torch.arange(plan.n_rows, dtype=torch.float64))
    assert u[0].shape == (len(x), d) and u[2].shape == (len(x), d, d)
    assert torch.equal(plan.dofs("y_out", [2])[0], u[3][2].reshape(-1).long())
