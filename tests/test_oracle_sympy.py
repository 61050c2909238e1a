"""Pin of the CSR oracle (the reference holds no golden matrix or vector for this path): the element
tensors of oracle/assembly.py -- BOTH restatements, closed form and brute-force quadrature -- are held to exact sympy
derivations FROM THE INTEGRANDS of the reference's forms (tests/sympy_forms.py):

  demo/strong-dirichlet/flower/main.py:104-128   cells (stiffness + stabilisation), one-sided boundary term, ghost penalty,
                                                 load vector (+ stabilisation)
  demo/weak-dirichlet/flower/main.py:112-151     the mixed (u, p) operator and right-hand side
  demo/neumann/square/main.py:113-158            the mixed (u, y, p) operator, P1 level set; demo/robin/square/main.py:121-172
                                                 with the robin_coef terms (oracle/assembly.py neumann_*)
  demo/interface-elasticity/main.py:183-269      the 5-field operator: cut / uncut cell tensors, load vectors, (y n).v on
                                                 ds(100) / ds(101), stress-jump penalties on dS(3) / dS(4) (oracle/elasticity.py)

on random rational simplices (d = 2, 3), random rational coefficients, random cell-local vertex orders, both sides of a
shared facet.  Tolerance: 1e-13 of the tensor's largest entry (the oracle is fp64, the derivation exact).
What this does NOT pin are the dolfinx conventions the forms are assembled under ('+' side of dS, which cell owns
ds(100), the sparsity pattern): those wait for baseline/dolfinx_reference.py (tests/test_dolfinx_fixture.py)."""
import random

import numpy as np
import pytest
import sympy as sp

import sympy_forms as SF
from oracle import assembly as OA
from oracle import tags as OT

TOL = 1e-13


def _rat(rng, lo=-9, hi=9, den=(1, 2, 3, 4, 5, 7)):
    return sp.Rational(rng.randint(lo, hi), rng.choice(den))


def _simplex(rng, d):
    """Random rational, well-shaped simplex."""
    while True:
        verts = [tuple(_rat(rng) for _ in range(d)) for _ in range(d + 1)]
        M = sp.Matrix([[1] + list(v) for v in verts])
        if abs(M.det()) > sp.Rational(1, 2):
            return verts


def _pair(rng, d):
    """Two simplices sharing local facets (op of '+', om of '-'), vertices in random local order; returns
    (points, cell+, cell-) with cells as index tuples into points."""
    while True:
        verts = _simplex(rng, d)
        X = SF.symbols(d)
        lam = SF.lagrange_basis(verts, X)
        b = tuple(_rat(rng) for _ in range(d))
        val = lam[d].subs(dict(zip(X, b)))
        if val < -sp.Rational(1, 10):                          # b on the other side of the facet opposite vertex d
            break
    pts = list(verts) + [b]
    plus = list(range(d + 1))
    minus = list(range(d)) + [d + 1]
    rng.shuffle(plus)
    rng.shuffle(minus)
    return pts, tuple(plus), tuple(minus)


def _close(got, want, what):
    scale = np.abs(want).max()
    assert scale > 0
    err = np.abs(got - want).max()
    assert err <= TOL * scale, "%s: %.3e of the largest entry" % (what, err / scale)


@pytest.mark.parametrize("d", [2, 3])
@pytest.mark.parametrize("cut", [False, True])
def test_strong_dirichlet_cell_tensors_from_the_integrands(d, cut):
    rng = random.Random(100 * d + cut)
    for _ in range(2):
        verts = _simplex(rng, d)
        phi = [_rat(rng) for _ in range(d + 1)]
        f = [_rat(rng) for _ in range(d + 1)]
        sigma = sp.Rational(rng.randint(1, 9), 4)
        A, b = SF.strong_cell(verts, phi, f, cut, sigma)
        A, b = SF.to_float(A), SF.to_float(b)[:, 0]
        x = np.array(verts, dtype=float)
        cells = np.arange(d + 1)[None, :]
        ph, fv = np.array(phi, dtype=float), np.array(f, dtype=float)
        Ac, bc = OA.cell_tensors_closed_form(x, cells, ph, fv, np.array([cut]), float(sigma))
        Aq, bq = OA.cell_tensors_quadrature(x, cells, ph[cells], fv[cells], np.array([cut]), float(sigma))
        _close(Ac[0], A, "closed-form cell matrix")
        _close(Aq[0], A, "quadrature cell matrix")
        _close(bc[0], b, "closed-form cell vector")
        _close(bq[0], b, "quadrature cell vector")
        assert np.abs(A - A.T).max() <= 1e-15 * np.abs(A).max()      # the cell part of a(., .) is symmetric


@pytest.mark.parametrize("d", [2, 3])
def test_strong_dirichlet_boundary_tensors_from_the_integrands(d):
    rng = random.Random(7 + d)
    verts = _simplex(rng, d)
    phi = [_rat(rng) for _ in range(d + 1)]
    x = np.array(verts, dtype=float)
    cells = np.arange(d + 1)[None, :]
    ph = np.array(phi, dtype=float)
    for o in range(d + 1):                                          # every local facet
        A = SF.to_float(SF.strong_boundary(verts, phi, o))
        ents = np.array([[0, o]])
        _close(OA.boundary_tensors_closed_form(x, cells, ph, ents)[0], A, "closed-form one-sided matrix, facet %d" % o)
        _close(OA.boundary_tensors_quadrature(x, cells, ph[cells], ents)[0], A, "quadrature one-sided matrix")
        assert np.abs(A[o]).max() == 0.0                            # v_o vanishes on the facet opposite vertex o
        assert np.abs(A - A.T).max() > 1e-3 * np.abs(A).max()       # this term is NOT symmetric


def _mesh_of_pair(pts, plus, minus, d):
    x = np.array(pts, dtype=float)
    cells = np.array([plus, minus], dtype=np.int64)
    ct = "triangle" if d == 2 else "tetrahedron"
    c2f, f2c, _ = OT.build_topology(cells, ct)
    shared = np.nonzero(f2c[:, 1] >= 0)[0]
    assert len(shared) == 1 and tuple(f2c[shared[0]]) == (0, 1)     # '+' = first cell of the facet
    fct = int(shared[0])
    op = int(np.nonzero(c2f[0] == fct)[0][0])
    om = int(np.nonzero(c2f[1] == fct)[0][0])
    return x, cells, c2f, f2c, fct, op, om


@pytest.mark.parametrize("d", [2, 3])
def test_strong_dirichlet_ghost_penalty_from_the_integrands(d):
    rng = random.Random(31 + d)
    for _ in range(2):
        pts, plus, minus = _pair(rng, d)
        x, cells, c2f, f2c, fct, op, om = _mesh_of_pair(pts, plus, minus, d)
        phi = [_rat(rng) for _ in range(d + 2)]
        sigma = sp.Rational(rng.randint(1, 9), 4)
        vp, vm = [pts[k] for k in plus], [pts[k] for k in minus]
        assert plus[op] == d and minus[om] == d + 1                 # local facet = opposite the non-shared vertex
        E = SF.to_float(SF.strong_ghost(vp, vm, op, om, [phi[k] for k in plus], [phi[k] for k in minus], sigma))
        ph = np.array(phi, dtype=float)
        Ec, macro = OA.ghost_tensors_closed_form(x, cells, ph, c2f, f2c, [fct], float(sigma))
        assert list(macro[0]) == list(plus) + list(minus)
        Eq = OA.ghost_tensors_quadrature(x, cells, ph[cells[:1]], ph[cells[1:]], c2f, f2c, [fct], float(sigma))
        _close(Ec[0], E, "closed-form ghost-penalty macro matrix")
        _close(Eq[0], E, "quadrature ghost-penalty macro matrix")
        # phi w is continuous across the facet: the jump of its gradient has no tangential part, and the rows of the two
        # copies of a shared vertex, summed as the global assembly does, are what a conforming function sees
        assert np.abs(E - E.T).max() <= 1e-15 * np.abs(E).max()


@pytest.mark.parametrize("d", [2, 3])
@pytest.mark.parametrize("cut", [False, True])
def test_weak_dirichlet_cell_tensors_from_the_integrands(d, cut):
    rng = random.Random(500 + 10 * d + cut)
    verts = _simplex(rng, d)
    phi, f, ud = ([_rat(rng) for _ in range(d + 1)] for _ in range(3))
    gamma, sigma = sp.Rational(rng.randint(1, 9), 2), sp.Rational(rng.randint(1, 9), 4)
    A, b = SF.weak_cell(verts, phi, f, ud, cut, gamma, sigma)
    A, b = SF.to_float(A), SF.to_float(b)[:, 0]
    x = np.array(verts, dtype=float)
    cells = np.arange(d + 1)[None, :]
    ph, fv, uv = (np.array(v, dtype=float) for v in (phi, f, ud))
    Ac, bc = OA.weak_cell_tensors_closed_form(x, cells, ph, fv, uv, np.array([cut]), float(gamma), float(sigma))
    Aq, bq = OA.weak_cell_tensors_quadrature(x, cells, ph[cells], fv[cells], uv[cells], np.array([cut]), float(gamma),
                                             float(sigma))
    _close(Ac[0], A, "closed-form weak cell matrix")
    _close(Aq[0], A, "quadrature weak cell matrix")
    _close(bc[0], b, "closed-form weak cell vector")
    _close(bq[0], b, "quadrature weak cell vector")


@pytest.mark.parametrize("d", [2, 3])
def test_weak_dirichlet_facet_tensors_from_the_integrands(d):
    rng = random.Random(900 + d)
    pts, plus, minus = _pair(rng, d)
    x, cells, c2f, f2c, fct, op, om = _mesh_of_pair(pts, plus, minus, d)
    vp, vm = [pts[k] for k in plus], [pts[k] for k in minus]
    sigma = sp.Rational(rng.randint(1, 9), 4)
    for o in range(d + 1):
        B = SF.to_float(SF.weak_boundary(vp, o))
        ents = np.array([[0, o]])
        _close(OA.weak_boundary_tensors_closed_form(x, cells, ents)[0], B, "closed-form weak one-sided matrix")
        _close(OA.weak_boundary_tensors_quadrature(x, cells, ents)[0], B, "quadrature weak one-sided matrix")
    E = SF.to_float(SF.weak_ghost(vp, vm, op, om, sigma))
    _close(OA.weak_ghost_tensors_closed_form(x, cells, c2f, f2c, [fct], float(sigma))[0], E, "closed-form weak ghost")
    _close(OA.weak_ghost_tensors_quadrature(x, cells, c2f, f2c, [fct], float(sigma))[0], E, "quadrature weak ghost")


def test_strong_dirichlet_p2_tensors_from_the_integrands():
    """P2 trial/test space and P2 level set on a triangle (BASELINE.json configs[2]): the quadrature restatement (the
    only one for P2) against the exact integrals."""
    rng = random.Random(2222)
    d = 2
    pts, plus, minus = _pair(rng, d)
    x, cells, c2f, f2c, fct, op, om = _mesh_of_pair(pts, plus, minus, d)
    vp, vm = [pts[k] for k in plus], [pts[k] for k in minus]
    nd = 6
    # a global quadratic level set / source: the same function seen from both cells
    X = SF.symbols(d)
    glob = [sum(_rat(rng) * m for m in (1, X[0], X[1], X[0] ** 2, X[0] * X[1], X[1] ** 2)) for _ in range(2)]

    def dofs(expr, verts):                       # values at the P2 nodes: vertices, then edge midpoints (1,2),(0,2),(0,1)
        nodes = list(verts) + [tuple((verts[a][k] + verts[b][k]) / 2 for k in range(d)) for a, b in ((1, 2), (0, 2), (0, 1))]
        return [expr.subs(dict(zip(X, p))) for p in nodes]

    phi_p, phi_m, f_p = dofs(glob[0], vp), dofs(glob[0], vm), dofs(glob[1], vp)
    sigma = sp.Rational(3, 2)
    A, b = SF.strong_cell(vp, phi_p, f_p, True, sigma, kw=2, kphi=2)
    ph = np.array([phi_p], dtype=float)
    Aq, bq = OA.cell_tensors_quadrature(x, cells[:1], ph, np.array([f_p], dtype=float), np.array([True]), float(sigma),
                                        kphi=2, kw=2)
    _close(Aq[0], SF.to_float(A), "P2 cell matrix")
    _close(bq[0], SF.to_float(b)[:, 0], "P2 cell vector")
    assert Aq.shape == (1, nd, nd)
    B = SF.strong_boundary(vp, phi_p, op, kw=2, kphi=2)
    _close(OA.boundary_tensors_quadrature(x, cells, ph, np.array([[0, op]]), kphi=2, kw=2)[0], SF.to_float(B),
           "P2 one-sided matrix")
    E = SF.strong_ghost(vp, vm, op, om, phi_p, phi_m, sigma, kw=2, kphi=2)
    Eq = OA.ghost_tensors_quadrature(x, cells, ph, np.array([phi_m], dtype=float), c2f, f2c, [fct], float(sigma),
                                     kphi=2, kw=2)
    _close(Eq[0], SF.to_float(E), "P2 ghost-penalty macro matrix")


@pytest.mark.parametrize("d", [2, 3])
@pytest.mark.parametrize("cut,kappa", [(False, 0), (True, 0), (True, sp.Rational(3, 4))])
def test_neumann_robin_cell_tensors_from_the_integrands(d, cut, kappa):
    """Mixed P1 x P1^d x DG0 operator of demo/neumann (kappa = 0) and demo/robin (kappa = robin_coef), P1 level set:
    both oracle forms against the exact integrals of the literal integrands."""
    rng = random.Random(1300 + 10 * d + cut + int(4 * kappa))
    verts = _simplex(rng, d)
    phi, f, un = ([_rat(rng) for _ in range(d + 1)] for _ in range(3))
    gamma = sp.Rational(rng.randint(1, 9), 2)
    A, b = SF.neumann_cell(verts, phi, f, un, cut, gamma, kappa)
    A, b = SF.to_float(A), SF.to_float(b)[:, 0]
    x = np.array(verts, dtype=float)
    cells = np.arange(d + 1)[None, :]
    ph, fv, uv = (np.array(v, dtype=float) for v in (phi, f, un))
    Ac, bc = OA.neumann_cell_tensors_closed_form(x, cells, ph, fv, uv, np.array([cut]), float(gamma), float(kappa))
    Aq, bq = OA.neumann_cell_tensors_quadrature(x, cells, ph[cells], fv[cells], uv[cells], np.array([cut]),
                                                float(gamma), kappa=float(kappa))
    _close(Ac[0], A, "closed-form Neumann / Robin cell matrix")
    _close(Aq[0], A, "quadrature Neumann / Robin cell matrix")
    _close(bc[0], b, "closed-form Neumann / Robin cell vector")
    _close(bq[0], b, "quadrature Neumann / Robin cell vector")


@pytest.mark.parametrize("d", [2, 3])
def test_neumann_facet_tensors_from_the_integrands(d):
    rng = random.Random(1700 + d)
    pts, plus, minus = _pair(rng, d)
    x, cells, c2f, f2c, fct, op, om = _mesh_of_pair(pts, plus, minus, d)
    vp, vm = [pts[k] for k in plus], [pts[k] for k in minus]
    sigma = sp.Rational(rng.randint(1, 9), 4)
    for o in range(d + 1):
        B = SF.to_float(SF.neumann_boundary(vp, o))
        _close(OA.neumann_boundary_tensors(x, cells, np.array([[0, o]]))[0], B, "Neumann one-sided matrix (y.n) v")
    E = SF.to_float(SF.neumann_ghost(vp, vm, op, om, sigma))
    _close(OA.neumann_ghost_tensors(x, cells, c2f, f2c, [fct], float(sigma))[0], E, "Neumann ghost penalty")


@pytest.mark.parametrize("d,tag", [(2, 2), (2, 1), (2, 3), (3, 2)])
def test_interface_elasticity_cell_tensors_from_the_integrands(d, tag):
    """5-field operator of demo/interface-elasticity (BASELINE configs[3]), P1 level set: entries of both oracle forms
    against the exact integrals of the literal integrands -- every entry in 2D for a cut cell, a random sample of 260
    entries (and the whole load vector) otherwise."""
    from oracle import elasticity as OE
    rng = random.Random(2100 + 10 * d + tag)
    verts = _simplex(rng, d)
    nv = d + 1
    nb = OE.Offsets(d).nb
    nm = nv * nb
    phi = [_rat(rng) for _ in range(nv)]
    fd = [[_rat(rng) for _ in range(d)] for _ in range(nv)]
    Ei, Eo = sp.Rational(rng.randint(2, 9), 2), sp.Rational(rng.randint(1, 5), 7)
    nui, nuo = sp.Rational(3, 10), sp.Rational(1, 4)
    lame = lambda E, nu: (E * nu / (1 + nu) / (1 - 2 * nu), E / 2 / (1 + nu))      # noqa: E731  data.py:5-10
    (li, mi), (lo, mo) = lame(Ei, nui), lame(Eo, nuo)
    ci, co = (Ei / (Ei + Eo)) ** 2, (Eo / (Ei + Eo)) ** 2
    gamma, sigma_s = sp.Rational(rng.randint(1, 9), 2), sp.Rational(rng.randint(1, 9), 4)
    if d == 2 and tag == 2:
        pairs = [(i, j) for i in range(nm) for j in range(i, nm)]
    else:
        pairs = [(rng.randrange(nm), rng.randrange(nm)) for _ in range(260)]
    ent, b = SF.elasticity_cell_entries(verts, phi, fd, tag, li, mi, lo, mo, ci, co, gamma, sigma_s, pairs)
    mat = OE.Material(float(Ei), float(nui), float(Eo), float(nuo))
    x = np.array(verts, dtype=float)
    cells = np.arange(nv)[None, :]
    ph = np.array(phi, dtype=float)
    fv = np.array(fd, dtype=float)
    Ac, bc = OE.cell_tensors_closed_form(x, cells, ph, fv, np.array([tag]), mat, float(gamma), float(sigma_s))
    Aq, bq = OE.cell_tensors_quadrature(x, cells, ph[cells], fv[cells], np.array([tag]), mat, float(gamma),
                                        float(sigma_s))
    want = np.array([float(sp.N(ent[pq], 30)) for pq in pairs])
    scale = max(np.abs(want).max(), 1e-300)
    for A, what in ((Ac[0], "closed form"), (Aq[0], "quadrature")):
        got = np.array([A[i, j] for (i, j) in pairs])
        assert np.abs(got - want).max() <= 1e-11 * scale, what
        assert np.abs(A - A.T).max() <= 1e-12 * scale, what + " (symmetry)"
    bw = np.array([float(sp.N(v, 30)) for v in b])
    for bb, what in ((bc[0], "closed form"), (bq[0], "quadrature")):
        assert np.abs(bb - bw).max() <= 1e-11 * max(np.abs(bw).max(), 1e-300), what + " load vector"


@pytest.mark.parametrize("d", [2, 3])
def test_interface_elasticity_facet_tensors_from_the_integrands(d):
    from oracle import elasticity as OE
    rng = random.Random(2500 + d)
    pts, plus, minus = _pair(rng, d)
    x, cells, c2f, f2c, fct, op, om = _mesh_of_pair(pts, plus, minus, d)
    vp, vm = [pts[k] for k in plus], [pts[k] for k in minus]
    Ei, Eo, nui, nuo = sp.Rational(7, 2), sp.Rational(3, 7), sp.Rational(3, 10), sp.Rational(1, 4)
    lame = lambda E, nu: (E * nu / (1 + nu) / (1 - 2 * nu), E / 2 / (1 + nu))      # noqa: E731
    mat = OE.Material(float(Ei), float(nui), float(Eo), float(nuo))
    sigma_s = sp.Rational(rng.randint(1, 9), 4)
    for side, (E_, nu_) in (("in", (Ei, nui)), ("out", (Eo, nuo))):
        o = rng.randrange(d + 1)
        B = SF.to_float(SF.elasticity_boundary(vp, o, side))
        ents = np.array([[0, o]])
        _close(OE.boundary_tensors_closed_form(x, cells, ents, side)[0], B, "closed-form (y n).v, side " + side)
        _close(OE.boundary_tensors_quadrature(x, cells, ents, side)[0], B, "quadrature (y n).v, side " + side)
        lm, mu = lame(E_, nu_)
        E = SF.to_float(SF.elasticity_facet(vp, vm, op, om, side, lm, mu, sigma_s))
        _close(OE.facet_tensors_closed_form(x, cells, c2f, f2c, [fct], side, mat, float(sigma_s))[0], E,
               "closed-form stress-jump penalty, side " + side)
        _close(OE.facet_tensors_quadrature(x, cells, c2f, f2c, [fct], side, mat, float(sigma_s))[0], E,
               "quadrature stress-jump penalty, side " + side)
