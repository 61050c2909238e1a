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


# ---------------------------------------------------------------------------------------------------------------
# Neumann / Robin (reference demo/neumann/square/main.py, demo/robin/square/main.py), mixed space P1 x P1^d x DG0,
# cell-local basis [u_0..u_d, y_(0,0)..y_(d,d-1) (vertex-major), p]; P1 level set (grad phi constant on the cell)
# ---------------------------------------------------------------------------------------------------------------
def _neumann_basis(verts, X):
    d = len(verts) - 1
    W = lagrange_basis(verts, X, 1)
    zero_vec = [sp.Integer(0)] * d
    basis = [(w, list(zero_vec), sp.Integer(0)) for w in W]                     # (u, y, p)
    for j in range(d + 1):
        for c in range(d):
            y = list(zero_vec)
            y[c] = W[j]
            basis.append((sp.Integer(0), y, sp.Integer(0)))
    basis.append((sp.Integer(0), list(zero_vec), sp.Integer(1)))
    return W, basis


def _div(y, X):
    return sum(sp.diff(sp.sympify(y[c]), X[c]) for c in range(len(X)))


def neumann_cell(verts, phi_dofs, f_dofs, un_dofs, cut, gamma, kappa=0):
    """a: (inner(grad u, grad v) + u v) dx((1,2)) + gamma [ inner(y + grad u, z + grad v) + (div y + u)(div z + v)
    + h^-2 (y.grad phi - |grad phi| kappa u + h^-1 p phi)(z.grad phi - |grad phi| kappa v + h^-1 q phi) ] dx(2)
    (neumann main.py:113-131, kappa = 0; robin main.py:121-140, kappa = robin_coef);
    L: f v dx((1,2)) + gamma [ -h^-2 u_N |grad phi| (z.grad phi - |grad phi| kappa v + h^-1 q phi) + f (div z + v) ] dx(2)
    (neumann :148-158, robin :157-172)."""
    d = len(verts) - 1
    X = symbols(d)
    W, basis = _neumann_basis(verts, X)
    phi = sum(c * s for c, s in zip(phi_dofs, W))
    f = sum(c * s for c, s in zip(f_dofs, W))
    un = sum(c * s for c, s in zip(un_dofs, W))
    gphi = grad(phi, X)
    ng = sp.sqrt(dot(gphi, gphi))
    h = cell_diameter(verts)
    nm = len(basis)

    def s1(u, y):
        gu = grad(sp.sympify(u), X)
        return [y[c] + gu[c] for c in range(d)]

    def s3(u, y, p_):
        return dot(y, gphi) - ng * kappa * u + p_ * phi / h

    A = sp.zeros(nm, nm)
    b = sp.zeros(nm, 1)
    for i, (v, z, q) in enumerate(basis):
        for j, (u, y, p_) in enumerate(basis):
            val = integral(dot(grad(sp.sympify(u), X), grad(sp.sympify(v), X)) + u * v, verts, X)
            if cut:
                val += gamma * integral(dot(s1(u, y), s1(v, z)) + (_div(y, X) + u) * (_div(z, X) + v)
                                        + h ** -2 * s3(u, y, p_) * s3(v, z, q), verts, X)
            A[i, j] = val
        val = integral(f * v, verts, X)
        if cut:
            val += gamma * integral(-h ** -2 * un * ng * s3(v, z, q) + f * (_div(z, X) + v), verts, X)
        b[i] = val
    return A, b


def neumann_boundary(verts, o):
    """inner(inner(y, n), v) ds  (neumann main.py:114, robin :123): rows v, columns y."""
    d = len(verts) - 1
    X = symbols(d)
    W, basis = _neumann_basis(verts, X)
    n = outward_normal(verts, o, X)
    F = facet_points(verts, o)
    nm = len(basis)
    A = sp.zeros(nm, nm)
    for i, (v, z, q) in enumerate(basis):
        for j, (u, y, p_) in enumerate(basis):
            A[i, j] = integral(dot(y, n) * v, F, X)
    return A


def neumann_ghost(verts_p, verts_m, op, om, sigma):
    """sigma avg(h_T) inner(jump(grad u, n), jump(grad v, n)) dS  (neumann main.py:132-135 over dS(3), robin :141-148 over
    dS(2)) on the macro dofs [mixed dofs of cell +, mixed dofs of cell -]: only the u functions jump."""
    d = len(verts_p) - 1
    X = symbols(d)
    Wp, bp = _neumann_basis(verts_p, X)
    Wm, bm = _neumann_basis(verts_m, X)
    J, X = _macro_jumps(verts_p, verts_m, op, om, [u for u, _, _ in bp], [u for u, _, _ in bm])
    avg_h = (cell_diameter(verts_p) + cell_diameter(verts_m)) / 2
    F = facet_points(verts_p, op)
    n = len(J)
    E = sp.zeros(n, n)
    for a in range(n):
        for c in range(a, n):
            E[a, c] = E[c, a] = sigma * avg_h * integral(J[a] * J[c], F, X)
    return E


# ---------------------------------------------------------------------------------------------------------------
# interface elasticity (reference demo/interface-elasticity/main.py:152-269, data.py:24-36), mixed space
# (u_in, u_out, y_in, y_out, p) in P1^d x P1^d x P1^(d x d) x P1^(d x d) x P1^d; cell-local dof (vertex k, offset o) -> k NB + o
# with o: u_in c -> c, u_out c -> d + c, y_in (r, s) -> 2 d + r d + s, y_out (r, s) -> 2 d + d^2 + r d + s, p c -> 2 d + 2 d^2 + c
# ---------------------------------------------------------------------------------------------------------------
def _elasticity_basis(verts, X):
    d = len(verts) - 1
    W = lagrange_basis(verts, X, 1)
    zv = lambda: [sp.Integer(0)] * d                             # noqa: E731
    zm = lambda: [[sp.Integer(0)] * d for _ in range(d)]         # noqa: E731
    basis = []
    for k in range(d + 1):
        for c in range(d):
            f = dict(ui=zv(), uo=zv(), yi=zm(), yo=zm(), p=zv())
            f["ui"][c] = W[k]
            basis.append(f)
        for c in range(d):
            f = dict(ui=zv(), uo=zv(), yi=zm(), yo=zm(), p=zv())
            f["uo"][c] = W[k]
            basis.append(f)
        for name in ("yi", "yo"):
            for r in range(d):
                for s_ in range(d):
                    f = dict(ui=zv(), uo=zv(), yi=zm(), yo=zm(), p=zv())
                    f[name][r][s_] = W[k]
                    basis.append(f)
        for c in range(d):
            f = dict(ui=zv(), uo=zv(), yi=zm(), yo=zm(), p=zv())
            f["p"][c] = W[k]
            basis.append(f)
    return W, basis


def _grad_vec(u, X):          # grad(u)_ij = d_j u_i
    return [[sp.diff(sp.sympify(u[i]), X[j]) for j in range(len(X))] for i in range(len(X))]


def _sigma_eps(u, X, lmbda, mu):
    d = len(X)
    g = _grad_vec(u, X)
    eps = [[(g[i][j] + g[j][i]) / 2 for j in range(d)] for i in range(d)]
    tr = sum(g[i][i] for i in range(d))
    sig = [[lmbda * tr * (1 if i == j else 0) + 2 * mu * eps[i][j] for j in range(d)] for i in range(d)]
    return sig, eps


def _ddot(a, b):
    return sum(a[i][j] * b[i][j] for i in range(len(a)) for j in range(len(a)))


def _div_mat(y, X):           # (div y)_i = d_j y_ij
    return [sum(sp.diff(sp.sympify(y[i][j]), X[j]) for j in range(len(X))) for i in range(len(X))]


def elasticity_cell_entries(verts, phi_dofs, f_dofs, tag, lam_in, mu_in, lam_out, mu_out, c_in, c_out, gamma, sigma_s,
                            pairs):
    """Entries A[i, j] for (i, j) in `pairs` and the whole vector b of the cell tensor of main.py:227-236 / :254-269 for a
    cell tagged `tag` (1: dx(1,2) only, 3: dx(2,3) only, 2: every cell term).  f_dofs: [d+1][d] nodal values of f."""
    d = len(verts) - 1
    X = symbols(d)
    W, basis = _elasticity_basis(verts, X)
    phi = sum(c * s for c, s in zip(phi_dofs, W))
    gphi = grad(phi, X)
    f = [sum(f_dofs[k][c] * W[k] for k in range(d + 1)) for c in range(d)]
    h = cell_diameter(verts)

    def terms(B):
        si, ei = _sigma_eps(B["ui"], X, lam_in, mu_in)
        so, eo = _sigma_eps(B["uo"], X, lam_out, mu_out)
        Ti = [[B["yi"][i][j] + si[i][j] for j in range(d)] for i in range(d)]
        To = [[B["yo"][i][j] + so[i][j] for j in range(d)] for i in range(d)]
        R = [sum((B["yi"][i][j] - B["yo"][i][j]) * gphi[j] for j in range(d)) for i in range(d)]
        S = [B["ui"][i] - B["uo"][i] + B["p"][i] * phi / h for i in range(d)]
        return dict(si=si, ei=ei, so=so, eo=eo, Ti=Ti, To=To, R=R, S=S, dyi=_div_mat(B["yi"], X), dyo=_div_mat(B["yo"], X))

    T = [terms(B) for B in basis]
    out = {}
    for (i, j) in pairs:          # i: test function, j: trial function
        tv, tu = T[i], T[j]
        val = 0
        if tag in (1, 2):
            val += integral(_ddot(tu["si"], tv["ei"]), verts, X)
        if tag in (2, 3):
            val += integral(_ddot(tu["so"], tv["eo"]), verts, X)
        if tag == 2:
            val += gamma * integral(c_out * _ddot(tu["Ti"], tv["Ti"]) + c_in * _ddot(tu["To"], tv["To"])
                                    + h ** -2 * dot(tu["R"], tv["R"]) + h ** -2 * dot(tu["S"], tv["S"]), verts, X)
            val += sigma_s * h ** 2 * integral(dot(tu["dyi"], tv["dyi"]) + dot(tu["dyo"], tv["dyo"]), verts, X)
        out[(i, j)] = val
    b = []
    for i, B in enumerate(basis):
        val = 0
        if tag in (1, 2):
            val += integral(dot(f, B["ui"]), verts, X)
        if tag in (2, 3):
            val += integral(dot(f, B["uo"]), verts, X)
        if tag == 2:
            val += sigma_s * h ** 2 * integral(dot(f, [a_ + c_ for a_, c_ in zip(T[i]["dyi"], T[i]["dyo"])]), verts, X)
        b.append(val)
    return out, b


def elasticity_boundary(verts, o, side):
    """inner(dot(y, n), v) on a (cell, local facet) entity: ds(100) with (y_in, v_in) [main.py:183, 235] or ds(101) with
    (y_out, v_out) [:184, 236]; (y n)_i = y_ij n_j."""
    d = len(verts) - 1
    X = symbols(d)
    W, basis = _elasticity_basis(verts, X)
    n = outward_normal(verts, o, X)
    F = facet_points(verts, o)
    yk, uk = ("yi", "ui") if side == "in" else ("yo", "uo")
    nm = len(basis)
    A = sp.zeros(nm, nm)
    for i, V in enumerate(basis):
        if all(c == 0 for c in V[uk]):
            continue
        for j, U in enumerate(basis):
            yn = [sum(U[yk][r][c] * n[c] for c in range(d)) for r in range(d)]
            if all(c == 0 for c in yn):
                continue
            A[i, j] = integral(dot(yn, V[uk]), F, X)
    return A


def elasticity_facet(verts_p, verts_m, op, om, side, lmbda, mu, sigma_s):
    """sigma_s avg(h_T) inner(jump(sigma(u), n), jump(sigma(v), n)) over dS(3) (side "in") / dS(4) ("out")
    [main.py:207-211, 221-225, 231-232] on the macro dofs [mixed dofs of cell +, of cell -]; jump(s, n) = s+ n+ + s- n-."""
    d = len(verts_p) - 1
    X = symbols(d)
    uk = "ui" if side == "in" else "uo"
    J = []
    for verts, o in ((verts_p, op), (verts_m, om)):
        W, basis = _elasticity_basis(verts, X)
        n = outward_normal(verts, o, X)
        for B in basis:
            sg, _ = _sigma_eps(B[uk], X, lmbda, mu)
            J.append([sum(sg[i][j] * n[j] for j in range(d)) for i in range(d)])
    avg_h = (cell_diameter(verts_p) + cell_diameter(verts_m)) / 2
    F = facet_points(verts_p, op)
    nm = len(J)
    E = sp.zeros(nm, nm)
    live = [a for a in range(nm) if any(c != 0 for c in J[a])]
    for a in live:
        for c in live:
            if c >= a:
                E[a, c] = E[c, a] = sigma_s * avg_h * integral(dot(J[a], J[c]), F, X)
    return E


def to_float(M, digits=30):
    import numpy as np
    return np.array([[float(sp.N(M[i, j], digits)) for j in range(M.shape[1])] for i in range(M.shape[0])])
