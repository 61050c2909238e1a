"""Exact (sympy, rational arithmetic) element tensors of the reference's UFL forms, derived FROM THE INTEGRANDS
-- a third derivation, independent of both restatements in oracle/assembly.py, that pins the oracle:

  strong Dirichlet   reference demo/strong-dirichlet/flower/main.py:104-128
  weak Dirichlet     reference demo/weak-dirichlet/flower/main.py:112-151

Nothing here knows a closed form: basis functions are built by inverting the vertex matrix over the rationals, the UFL
operators (grad, div, inner, jump, avg, CellDiameter, FacetNormal) are written out as UFL defines them [dep-knowledge:
jump(v, n) = v('+').n('+') + v('-').n('-'), avg(h) = (h('+') + h('-')) / 2, CellDiameter = largest vertex distance,
FacetNormal outward of the integrating cell], and every integral is taken exactly: the integrand is pulled back to
the reference simplex and integrated monomial by monomial (int xi^alpha = prod alpha_i! / (dim + |alpha|)!).  Square roots
(h_T, |F|) stay exact algebraic numbers until the final comparison.
TEST INFRASTRUCTURE ONLY.
"""
import itertools
import math

import sympy as sp


def symbols(d):
    return sp.symbols("x y z")[:d]


def lagrange_basis(verts, X, degree=1):
    """P1 / P2 Lagrange basis on the simplex `verts` (rational points) as polynomials in X; P2 in dolfinx's local
    order: vertices, then edges (1,2),(0,2),(0,1) / (2,3),(1,3),(1,2),(0,3),(0,2),(0,1) [dep-knowledge, SURVEY.md C.7]."""
    d = len(X)
    V = sp.Matrix([[1] + list(v) for v in verts])              # rows (1, x_k)
    C = V.inv()                                                # column i = coefficients of lambda_i
    lam = [sp.expand(C[0, i] + sum(C[1 + k, i] * X[k] for k in range(d))) for i in range(d + 1)]
    if degree == 1:
        return lam
    edges = {2: ((1, 2), (0, 2), (0, 1)), 3: ((2, 3), (1, 3), (1, 2), (0, 3), (0, 2), (0, 1))}[d]
    return [sp.expand(l * (2 * l - 1)) for l in lam] + [sp.expand(4 * lam[a] * lam[b]) for a, b in edges]


def grad(u, X):
    return [sp.diff(u, s) for s in X]


def dot(a, b):
    return sum(p * q for p, q in zip(a, b))


def laplacian(u, X):          # div(grad(u))
    return sum(sp.diff(u, s, 2) for s in X)


def cell_diameter(verts):
    """ufl.CellDiameter: the largest distance between two vertices (exact square root)."""
    return sp.sqrt(max(sum((a - b) ** 2 for a, b in zip(p, q)) for p, q in itertools.combinations(verts, 2)))


def _integrate_reference(poly, xi):
    """Exact integral of a polynomial in xi over the reference simplex {xi >= 0, sum xi <= 1}."""
    dim = len(xi)
    if dim == 0:
        return sp.nsimplify(poly)
    total = 0
    for monom, coeff in sp.Poly(sp.expand(poly), *xi).terms():
        num = 1
        for e in monom:
            num *= math.factorial(e)
        total += coeff * sp.Rational(num, math.factorial(dim + sum(monom)))
    return total


def integrate_simplex(expr, pts, X):
    """int over the (possibly lower-dimensional) simplex with vertices `pts` of the polynomial expr(X), with respect
    to the simplex's own measure.  Returns (reference integral, measure factor): integral = factor * reference
    integral, factor = |simplex| * dim!  (exact; a square root when the simplex is a facet)."""
    dim = len(pts) - 1
    xi = sp.symbols("xi0:%d" % max(dim, 1))[:dim]
    point = [pts[0][k] + sum(xi[m] * (pts[m + 1][k] - pts[0][k]) for m in range(dim)) for k in range(len(X))]
    pulled = expr.subs(dict(zip(X, point)), simultaneous=True)
    E = sp.Matrix([[pts[m + 1][k] - pts[0][k] for k in range(len(X))] for m in range(dim)])
    gram = (E * E.T).det() if dim else sp.Integer(1)
    return _integrate_reference(pulled, xi), sp.sqrt(gram)     # sqrt(det(E E^T)) = dim! |simplex|


def integral(expr, pts, X):
    ref, fac = integrate_simplex(expr, pts, X)
    return ref * fac


def outward_normal(verts, o, X):
    """ufl.FacetNormal of the cell `verts` on its local facet o (opposite vertex o): exact unit vector."""
    lam = lagrange_basis(verts, X)
    g = grad(lam[o], X)                                        # points towards vertex o, i.e. inwards
    nrm = sp.sqrt(dot(g, g))
    return [-c / nrm for c in g]


def facet_points(verts, o):
    return [v for k, v in enumerate(verts) if k != o]


# ---------------------------------------------------------------------------------------------------------------
# strong Dirichlet (reference demo/strong-dirichlet/flower/main.py)
# ---------------------------------------------------------------------------------------------------------------
def strong_cell(verts, phi_dofs, f_dofs, cut, sigma, kw=1, kphi=1):
    """a: inner(grad(phi w), grad(phi v)) dx((1,2))  [:105]  + sigma h_T^2 inner(div grad(phi w), div grad(phi v)) dx(2)
    [:107-112];  L: inner(f, phi v) dx((1,2)) - sigma h_T^2 inner(f, div grad(phi v)) dx(2)  [:126-128].
    Rows = test functions."""
    d = len(verts) - 1
    X = symbols(d)
    W = lagrange_basis(verts, X, kw)
    phi = sum(c * s for c, s in zip(phi_dofs, lagrange_basis(verts, X, kphi)))
    f = sum(c * s for c, s in zip(f_dofs, W))
    h2 = cell_diameter(verts) ** 2
    nd = len(W)
    A = sp.zeros(nd, nd)
    b = sp.zeros(nd, 1)
    for i, v in enumerate(W):
        phiv = phi * v
        for j, w in enumerate(W):
            phiw = phi * w
            A[i, j] = integral(dot(grad(phiw, X), grad(phiv, X)), verts, X)
            if cut:
                A[i, j] += sigma * h2 * integral(laplacian(phiw, X) * laplacian(phiv, X), verts, X)
        b[i] = integral(f * phiv, verts, X)
        if cut:
            b[i] -= sigma * h2 * integral(f * laplacian(phiv, X), verts, X)
    return A, b


def strong_boundary(verts, phi_dofs, o, kw=1, kphi=1):
    """- inner(inner(grad(phi w), n), phi v) ds  [:106] on local facet o of the cell, n outward of THAT cell."""
    d = len(verts) - 1
    X = symbols(d)
    W = lagrange_basis(verts, X, kw)
    phi = sum(c * s for c, s in zip(phi_dofs, lagrange_basis(verts, X, kphi)))
    n = outward_normal(verts, o, X)
    F = facet_points(verts, o)
    nd = len(W)
    A = sp.zeros(nd, nd)
    for i, v in enumerate(W):
        for j, w in enumerate(W):
            A[i, j] = -integral(dot(grad(phi * w, X), n) * phi * v, F, X)
    return A


def _macro_jumps(verts_p, verts_m, op, om, fields_p, fields_m):
    """For the macro element of a facet (cell '+' local facet op, cell '-' local facet om): the list, over the macro
    basis [basis of +, basis of -], of jump(grad(g), n) = grad(g)('+').n('+') + grad(g)('-').n('-'), a basis function
    of one cell being zero on the other."""
    d = len(verts_p) - 1
    X = symbols(d)
    J = []
    for verts, o, fields in ((verts_p, op, fields_p), (verts_m, om, fields_m)):
        n = outward_normal(verts, o, X)
        J += [dot(grad(g, X), n) for g in fields]
    return J, X


def strong_ghost(verts_p, verts_m, op, om, phi_p, phi_m, sigma, kw=1, kphi=1):
    """sigma avg(h_T) inner(jump(grad(phi w), n), jump(grad(phi v), n)) dS((2,3))  [:113-118] over the macro dofs
    [dofs of cell +, dofs of cell -]."""
    d = len(verts_p) - 1
    X = symbols(d)
    fields = []
    for verts, pd in ((verts_p, phi_p), (verts_m, phi_m)):
        phi = sum(c * s for c, s in zip(pd, lagrange_basis(verts, X, kphi)))
        fields.append([phi * w for w in lagrange_basis(verts, X, kw)])
    J, X = _macro_jumps(verts_p, verts_m, op, om, fields[0], fields[1])
    avg_h = (cell_diameter(verts_p) + cell_diameter(verts_m)) / 2
    F = facet_points(verts_p, op)
    n = len(J)
    E = sp.zeros(n, n)
    for a in range(n):
        for c in range(a, n):
            E[a, c] = E[c, a] = sigma * avg_h * integral(J[a] * J[c], F, X)
    return E


# ---------------------------------------------------------------------------------------------------------------
# weak Dirichlet (reference demo/weak-dirichlet/flower/main.py), mixed basis [u-functions, p-functions]
# ---------------------------------------------------------------------------------------------------------------
def weak_cell(verts, phi_dofs, f_dofs, ud_dofs, cut, gamma, sigma, kw=1, kphi=1):
    """a: inner(grad u, grad v) dx((1,2)) [:113] + gamma h^-2 inner(u - h^-1 phi p, v - h^-1 phi q) dx(2) [:115-122]
    + sigma h^2 inner(div grad u, div grad v) dx(2) [:123-128];  L: inner(f, v) dx((1,2)) + gamma h^-2 inner(u_D,
    v - h^-1 phi q) dx(2) - sigma h^2 inner(f, div grad v) dx(2)  [:142-151]."""
    d = len(verts) - 1
    X = symbols(d)
    W = lagrange_basis(verts, X, kw)
    phi = sum(c * s for c, s in zip(phi_dofs, lagrange_basis(verts, X, kphi)))
    f = sum(c * s for c, s in zip(f_dofs, W))
    ud = sum(c * s for c, s in zip(ud_dofs, W))
    h = cell_diameter(verts)
    nd = len(W)
    mixed = [(w, 0) for w in W] + [(0, w) for w in W]          # (u-component, p-component)
    A = sp.zeros(2 * nd, 2 * nd)
    b = sp.zeros(2 * nd, 1)
    for i, (v, q) in enumerate(mixed):
        for j, (u, p) in enumerate(mixed):
            val = integral(dot(grad(sp.sympify(u), X), grad(sp.sympify(v), X)), verts, X)
            if cut:
                val += gamma * h ** -2 * integral((u - phi * p / h) * (v - phi * q / h), verts, X)
                val += sigma * h ** 2 * integral(laplacian(sp.sympify(u), X) * laplacian(sp.sympify(v), X), verts, X)
            A[i, j] = val
        val = integral(f * v, verts, X)
        if cut:
            val += gamma * h ** -2 * integral(ud * (v - phi * q / h), verts, X)
            val -= sigma * h ** 2 * integral(f * laplacian(sp.sympify(v), X), verts, X)
        b[i] = val
    return A, b


def weak_boundary(verts, o, kw=1):
    """- inner(inner(grad u, n), v) ds  [:114]; only the u-u block is non-zero."""
    d = len(verts) - 1
    X = symbols(d)
    W = lagrange_basis(verts, X, kw)
    n = outward_normal(verts, o, X)
    F = facet_points(verts, o)
    nd = len(W)
    A = sp.zeros(2 * nd, 2 * nd)
    for i, v in enumerate(W):
        for j, u in enumerate(W):
            A[i, j] = -integral(dot(grad(u, X), n) * v, F, X)
    return A


def weak_ghost(verts_p, verts_m, op, om, sigma, kw=1):
    """sigma avg(h_T) inner(jump(grad u, n), jump(grad v, n)) dS((2,3))  [:129-134] over the macro dofs
    [mixed dofs of cell + (u then p), mixed dofs of cell -]."""
    d = len(verts_p) - 1
    X = symbols(d)
    Wp, Wm = lagrange_basis(verts_p, X, kw), lagrange_basis(verts_m, X, kw)
    zero = [sp.Integer(0)] * len(Wp)
    J, X = _macro_jumps(verts_p, verts_m, op, om, list(Wp) + zero, list(Wm) + zero)
    avg_h = (cell_diameter(verts_p) + cell_diameter(verts_m)) / 2
    F = facet_points(verts_p, op)
    n = len(J)
    E = sp.zeros(n, n)
    for a in range(n):
        for c in range(a, n):
            E[a, c] = E[c, a] = sigma * avg_h * integral(J[a] * J[c], F, X)
    return E


def to_float(M, digits=30):
    import numpy as np
    return np.array([[float(sp.N(M[i, j], digits)) for j in range(M.shape[1])] for i in range(M.shape[0])])
