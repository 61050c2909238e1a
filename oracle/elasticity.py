"""CPU oracle, part 3: the interface-elasticity phi-FEM operator in CSR (TEST INFRASTRUCTURE ONLY).

Restates, on arrays, what dolfinx assembles for the forms of the reference demo demo/interface-elasticity/main.py
(BASELINE.json configs[3]) on the mixed space (u_in, u_out, y_in, y_out, p) in P1^d x P1^d x P1^(d x d) x P1^(d x d) x P1^d
(`mixd_element`, main.py:127-129) on triangles / tetrahedra:

  a =   int_{dx(1,2)} sigma_in(u_in):eps(v_in) + int_{dx(2,3)} sigma_out(u_out):eps(v_out)              (:186-187, 228-229)
      + gamma int_{dx(2)} [ c_out (y_in + sigma_in(u_in)):(z_in + sigma_in(v_in))
                          + c_in (y_out + sigma_out(u_out)):(z_out + sigma_out(v_out))
                          + h^-2 ((y_in - y_out) grad phi).((z_in - z_out) grad phi)
                          + h^-2 (u_in - u_out + h^-1 p phi).(v_in - v_out + h^-1 q phi) ]               (:189-205, 230)
      + sigma_s int_{dS(3)} avg(h) [sigma_in(u_in) n].[sigma_in(v_in) n]                                (:207-211, 231)
      + sigma_s int_{dS(4)} avg(h) [sigma_out(u_out) n].[sigma_out(v_out) n]                            (:221-225, 232)
      + sigma_s int_{dx(2)} h^2 (div y_in.div z_in + div y_out.div z_out)                               (:213-219, 233-234)
      + int_{ds(100)} (y_in n).v_in + int_{ds(101)} (y_out n).v_out                                    (:183-184, 235-236)
  L =   int_{dx(1,2)} f.v_in + int_{dx(2,3)} f.v_out + sigma_s int_{dx(2)} h^2 f.(div z_in + div z_out)  (:254-269)

with c_in = (E_in / (E_in + E_out))^2, c_out = (E_out / (E_in + E_out))^2 (:189-190), sigma_x(u) = lambda_x div(u) I +
2 mu_x eps(u) (data.py:24-36), [s n] = s+ n+ + s- n- (ufl.jump(tensor, n)), (y n)_i = y_ij n_j, (div y)_i = d_j y_ij.
Dirichlet conditions (main.py:173-179, 238, 271-274): rows and columns of the constrained dofs are zeroed, their diagonal
is 1, b <- b - A g on the free rows (apply_lifting) and b = g on the constrained ones (bc.set).

The demo's `f` is a UFL expression of x (main.py:150); here f is a P1 vector field given by its nodal values.

Mixed numbering (ours; an input convention of the kernels): NB = 3 d + 2 d^2 dofs per vertex, global dof = NB vertex + o,
o: u_in c -> c, u_out c -> d + c, y_in (r, s) -> 2 d + r d + s, y_out (r, s) -> 2 d + d^2 + r d + s, p c -> 2 d + 2 d^2 + c.
Cell-local order: node-major (local vertex k, offset o) -> k NB + o.

PARITY PARTIALLY PINNED (no golden matrix in the reference, dolfinx not installable here): two independent
restatements, `*_quadrature` (the UFL expressions evaluated field by field at brute-force quadrature points) and
`*_closed_form` (entry formulas from the exact integrals of barycentric monomials, P1 level set), must agree to ~1e-14,
and BOTH are held to exact sympy integrals of the literal integrands (tests/test_oracle_sympy.py: cell tensors, load
vectors, one-sided and stress-jump facet tensors).  The dolfinx conventions of the global assembly (pattern, '+' side,
lifting) stay unpinned until baseline/dolfinx_reference.py has been run.
"""
import math

import numpy as np

from .assembly import (_bary_moment, _scatter, facet_geometry, lagrange_eval, simplex_geometry, simplex_rule,
                       sparsity_pattern)


class Offsets:
    def __init__(self, d):
        self.d = d
        self.ui, self.uo = 0, d
        self.yi, self.yo = 2 * d, 2 * d + d * d
        self.p = 2 * d + 2 * d * d
        self.nb = 3 * d + 2 * d * d


def lame(E, nu):
    """data.py:5-10."""
    return E * nu / (1.0 + nu) / (1.0 - 2.0 * nu), E / 2.0 / (1.0 + nu)


class Material:
    """E_in / E_out / nu of data.py:13-22 and the weights of main.py:189-190."""

    def __init__(self, E_in=1.0, nu_in=0.3, E_out=0.001, nu_out=0.3):
        self.lmbda_in, self.mu_in = lame(E_in, nu_in)
        self.lmbda_out, self.mu_out = lame(E_out, nu_out)
        self.coef_in = (E_in / (E_in + E_out)) ** 2
        self.coef_out = (E_out / (E_in + E_out)) ** 2


def mixed_dofmap(cells, d):
    nb = Offsets(d).nb
    return (nb * cells.astype(np.int64)[:, :, None] + np.arange(nb)[None, None, :]).reshape(len(cells), -1)


# --------------------------------------------------------------------------------------
# restatement 1: the UFL expressions, field by field, at quadrature points
# --------------------------------------------------------------------------------------
def _fields(lam, G):
    """For every mixed basis function m of one cell at the points lam: u_in [q, m, d], grad u_in [q, m, d, d]
    (grad(u)_ij = d_j u_i), same for u_out, y_in [q, m, d, d], div y_in [q, m, d], same for y_out, p [q, m, d]."""
    nq, nv = lam.shape
    d = nv - 1
    o = Offsets(d)
    nm = nv * o.nb
    z = lambda *s: np.zeros((nq, nm) + s)        # noqa: E731
    ui, gui, uo, guo = z(d), z(d, d), z(d), z(d, d)
    yi, dyi, yo, dyo, p = z(d, d), z(d), z(d, d), z(d), z(d)
    for k in range(nv):
        for c in range(d):
            for (u, gu, off) in ((ui, gui, o.ui), (uo, guo, o.uo)):
                m = k * o.nb + off + c
                u[:, m, c] = lam[:, k]
                gu[:, m, c, :] = G[k][None, :]
            p[:, k * o.nb + o.p + c, c] = lam[:, k]
            for s in range(d):
                for (y, dy, off) in ((yi, dyi, o.yi), (yo, dyo, o.yo)):
                    m = k * o.nb + off + c * d + s
                    y[:, m, c, s] = lam[:, k]
                    dy[:, m, c] = G[k, s]
    return ui, gui, uo, guo, yi, dyi, yo, dyo, p


def _sigma(gu, lmbda, mu):
    d = gu.shape[-1]
    eps = 0.5 * (gu + np.swapaxes(gu, -1, -2))
    tr = np.trace(gu, axis1=-2, axis2=-1)
    return lmbda * tr[..., None, None] * np.eye(d) + 2.0 * mu * eps, eps


def cell_tensors_quadrature(x, cells, phi_dofs, f_dofs, tags, mat, gamma, sigma_s, kphi=1, n=6):
    """phi_dofs [n, nd_phi], f_dofs [n, nv, d] cell-local; tags [n] in {1, 2, 3}.  -> A [n, NM, NM], b [n, NM]."""
    d = x.shape[1]
    G, vol, h = simplex_geometry(x, cells)
    lam, W = simplex_rule(d, n)
    out_A, out_b = [], []
    for c in range(len(cells)):
        ui, gui, uo, guo, yi, dyi, yo, dyo, p = _fields(lam, G[c])
        pv, pg, _ = lagrange_eval(lam, G[c], kphi)
        ph = pv @ phi_dofs[c]
        gph = np.einsum("qkd,k->qd", pg, phi_dofs[c])
        wq = W * math.factorial(d) * vol[c]
        fq = lam @ f_dofs[c]                                    # [q, d]
        si, ei = _sigma(gui, mat.lmbda_in, mat.mu_in)
        so, eo = _sigma(guo, mat.lmbda_out, mat.mu_out)
        nm = ui.shape[1]
        A = np.zeros((nm, nm))
        b = np.zeros(nm)
        t = tags[c]
        if t in (1, 2):
            A += np.einsum("q,qbij,qaij->ab", wq, si, ei)
            b += np.einsum("q,qi,qai->a", wq, fq, ui)
        if t in (2, 3):
            A += np.einsum("q,qbij,qaij->ab", wq, so, eo)
            b += np.einsum("q,qi,qai->a", wq, fq, uo)
        if t == 2:
            hh = h[c]
            Ti, To = yi + si, yo + so
            R = np.einsum("qmij,qj->qmi", yi - yo, gph)
            S = ui - uo + p * (ph / hh)[:, None, None]
            A += gamma * (mat.coef_out * np.einsum("q,qbij,qaij->ab", wq, Ti, Ti)
                          + mat.coef_in * np.einsum("q,qbij,qaij->ab", wq, To, To)
                          + np.einsum("q,qbi,qai->ab", wq, R, R) / hh ** 2
                          + np.einsum("q,qbi,qai->ab", wq, S, S) / hh ** 2)
            A += sigma_s * hh ** 2 * (np.einsum("q,qbi,qai->ab", wq, dyi, dyi)
                                      + np.einsum("q,qbi,qai->ab", wq, dyo, dyo))
            b += sigma_s * hh ** 2 * np.einsum("q,qi,qai->a", wq, fq, dyi + dyo)
        out_A.append(A)
        out_b.append(b)
    return np.array(out_A), np.array(out_b)


def boundary_tensors_quadrature(x, cells, ents, side, n=4):
    """int_F (y n).v on (cell, local facet) pairs; side "in" (ds(100)) or "out" (ds(101))."""
    ents = np.asarray(ents).reshape(-1, 2)
    d = x.shape[1]
    nv = d + 1
    o = Offsets(d)
    nrm, area = facet_geometry(x, cells, ents)
    G, _, _ = simplex_geometry(x, cells[ents[:, 0]])
    flam, fw = simplex_rule(d - 1, n)
    fw = fw * math.factorial(d - 1)
    out = []
    for e, (_, of) in enumerate(ents):
        on = [k for k in range(nv) if k != of]
        lam = np.zeros((len(fw), nv))
        lam[:, on] = flam
        ui, _, uo, _, yi, _, yo, _, _ = _fields(lam, G[e])
        u, y = (ui, yi) if side == "in" else (uo, yo)
        yn = np.einsum("qmij,j->qmi", y, nrm[e])
        out.append(np.einsum("q,qbi,qai->ab", fw * area[e], yn, u))
    return np.array(out).reshape(len(ents), nv * o.nb, nv * o.nb)


def facet_tensors_quadrature(x, cells, c2f, f2c, facets, side, mat, sigma_s):
    """sigma_s avg(h) int_F [sigma(u) n].[sigma(v) n] over the macro dofs [cell +, cell -]; side "in" / "out"."""
    d = x.shape[1]
    nv = d + 1
    o = Offsets(d)
    nm = nv * o.nb
    facets = np.asarray(facets)
    lmbda, mu = (mat.lmbda_in, mat.mu_in) if side == "in" else (mat.lmbda_out, mat.mu_out)
    E = np.zeros((len(facets), 2 * nm, 2 * nm))
    if len(facets) == 0:
        return E
    J = np.zeros((len(facets), 2 * nm, d))
    hsum = np.zeros(len(facets))
    area = None
    lam0 = np.full((1, nv), 1.0 / nv)             # sigma(u) is constant on the cell
    for s in (0, 1):
        cc = f2c[facets, s]
        lf = np.argmax(c2f[cc] == facets[:, None], axis=1)
        G, _, h = simplex_geometry(x, cells[cc])
        nrm, ar = facet_geometry(x, cells, np.stack([cc, lf], axis=1))
        for e in range(len(facets)):
            _, gui, _, guo, *_ = _fields(lam0, G[e])
            sg, _ = _sigma(gui if side == "in" else guo, lmbda, mu)
            J[e, s * nm:(s + 1) * nm] = np.einsum("mij,j->mi", sg[0], nrm[e])
        hsum += h
        if s == 0:
            area = ar
    return (sigma_s * 0.5 * hsum * area)[:, None, None] * np.einsum("eai,ebi->eab", J, J)


# --------------------------------------------------------------------------------------
# restatement 2: entry formulas from exact monomial integrals (P1 level set)
# --------------------------------------------------------------------------------------
def cell_tensors_closed_form(x, cells, phi, f, tags, mat, gamma, sigma_s):
    """phi [Nv], f [Nv, d] nodal."""
    d = x.shape[1]
    nv = d + 1
    o = Offsets(d)
    nm = nv * o.nb
    G, vol, h = simplex_geometry(x, cells)
    n = len(cells)
    pc, fc = phi[cells], f[cells]
    M = np.array([[_bary_moment(d, (i, j)) for j in range(nv)] for i in range(nv)])
    m1 = 1.0 / nv
    T3 = np.array([[[_bary_moment(d, (i, j, l)) for l in range(nv)] for j in range(nv)] for i in range(nv)])
    T4 = np.array([[[[_bary_moment(d, (i, j, l, q)) for q in range(nv)] for l in range(nv)] for j in range(nv)]
                   for i in range(nv)])
    Mphi = np.einsum("kjl,nl->nkj", T3, pc)              # int lam_k lam_j phi / |K|
    Mphi2 = np.einsum("kjlq,nl,nq->nkj", T4, pc, pc)     # int lam_k lam_j phi^2 / |K|
    g = np.einsum("nk,nkd->nd", pc, G)
    GG = np.einsum("nkd,njd->nkj", G, G)
    fM = np.einsum("nlc,lk->nkc", fc, M)                 # int f_c lam_k / |K|
    fbar = fc.mean(axis=1)
    inn = np.isin(tags, (1, 2)).astype(float)
    out = np.isin(tags, (2, 3)).astype(float)
    cut = (tags == 2).astype(float)
    A = np.zeros((n, nm, nm))
    b = np.zeros((n, nm))
    dl = lambda a_, b_: 1.0 if a_ == b_ else 0.0        # noqa: E731
    sides = ((o.ui, o.yi, mat.lmbda_in, mat.mu_in, inn, mat.coef_out, +1.0),
             (o.uo, o.yo, mat.lmbda_out, mat.mu_out, out, mat.coef_in, -1.0))
    for k in range(nv):
        for j in range(nv):
            for (ou, oy, lm, mu, act, cpen, sg) in sides:
                for c in range(d):
                    a = k * o.nb + ou + c
                    for c2 in range(d):
                        bb = j * o.nb + ou + c2
                        dd = G[:, k, c] * G[:, j, c2]                       # div v_a div u_b
                        ee = dl(c, c2) * GG[:, k, j] + G[:, j, c] * G[:, k, c2]
                        A[:, a, bb] += vol * act * (lm * dd + mu * ee)
                        A[:, a, bb] += vol * cut * gamma * cpen * ((d * lm * lm + 4 * lm * mu) * dd + 2 * mu * mu * ee)
                    # y_b = lam_j E_rs against sigma(v_a): |K| m1 sigma(v_a)_rs
                    for r in range(d):
                        for s in range(d):
                            bb = j * o.nb + oy + r * d + s
                            sv = lm * G[:, k, c] * dl(r, s) + mu * (dl(r, c) * G[:, k, s] + G[:, k, r] * dl(s, c))
                            A[:, a, bb] += vol * cut * gamma * cpen * m1 * sv
                            A[:, bb, a] += vol * cut * gamma * cpen * m1 * sv      # symmetric partner (k <-> j swapped)
                for r in range(d):
                    for s in range(d):
                        a = k * o.nb + oy + r * d + s
                        A[:, a, j * o.nb + oy + r * d + s] += vol * cut * gamma * cpen * M[k, j]
                        for s2 in range(d):
                            bb = j * o.nb + oy + r * d + s2
                            A[:, a, bb] += vol * cut * sigma_s * h ** 2 * G[:, k, s] * G[:, j, s2]
                            # R term: both sides couple, sign sg * sg'
                            for (_, oy2, _, _, _, _, sg2) in sides:
                                bb2 = j * o.nb + oy2 + r * d + s2
                                A[:, a, bb2] += vol * cut * gamma * sg * sg2 * g[:, s] * g[:, s2] * M[k, j] / h ** 2
                # S term: u_in - u_out + p phi / h
                for c in range(d):
                    a = k * o.nb + ou + c
                    for (ou2, _, _, _, _, _, sg2) in sides:
                        A[:, a, j * o.nb + ou2 + c] += vol * cut * gamma * sg * sg2 * M[k, j] / h ** 2
                    bp = j * o.nb + o.p + c
                    A[:, a, bp] += vol * cut * gamma * sg * Mphi[:, k, j] / h ** 3
                    A[:, bp, a] += vol * cut * gamma * sg * Mphi[:, k, j] / h ** 3
            for c in range(d):
                A[:, k * o.nb + o.p + c, j * o.nb + o.p + c] += vol * cut * gamma * Mphi2[:, k, j] / h ** 4
        for (ou, oy, _, _, act, _, _) in sides:
            for c in range(d):
                b[:, k * o.nb + ou + c] += vol * act * fM[:, k, c]
                for s in range(d):
                    b[:, k * o.nb + oy + c * d + s] += vol * cut * sigma_s * h ** 2 * G[:, k, s] * fbar[:, c]
    return A, b


def boundary_tensors_closed_form(x, cells, ents, side):
    ents = np.asarray(ents).reshape(-1, 2)
    d = x.shape[1]
    nv = d + 1
    o = Offsets(d)
    ou, oy = (o.ui, o.yi) if side == "in" else (o.uo, o.yo)
    nrm, area = facet_geometry(x, cells, ents)
    A = np.zeros((len(ents), nv * o.nb, nv * o.nb))
    for e, (_, of) in enumerate(ents):
        on = [k for k in range(nv) if k != of]
        for k in on:
            for j in on:
                mkj = area[e] * (2.0 if k == j else 1.0) / (d * (d + 1))
                for c in range(d):
                    for s in range(d):
                        A[e, k * o.nb + ou + c, j * o.nb + oy + c * d + s] = nrm[e, s] * mkj
    return A


def facet_tensors_closed_form(x, cells, c2f, f2c, facets, side, mat, sigma_s):
    d = x.shape[1]
    nv = d + 1
    o = Offsets(d)
    nm = nv * o.nb
    facets = np.asarray(facets)
    lm, mu = (mat.lmbda_in, mat.mu_in) if side == "in" else (mat.lmbda_out, mat.mu_out)
    ou = o.ui if side == "in" else o.uo
    J = np.zeros((len(facets), 2 * nm, d))
    hsum = np.zeros(len(facets))
    area = np.zeros(len(facets))
    for s in (0, 1):
        if len(facets) == 0:
            break
        cc = f2c[facets, s]
        lf = np.argmax(c2f[cc] == facets[:, None], axis=1)
        G, _, h = simplex_geometry(x, cells[cc])
        nrm, ar = facet_geometry(x, cells, np.stack([cc, lf], axis=1))
        for k in range(nv):
            gn = (G[:, k] * nrm).sum(axis=1)
            for c in range(d):
                v = lm * G[:, k, c][:, None] * nrm + mu * G[:, k] * nrm[:, c][:, None]
                v[:, c] += mu * gn
                J[:, s * nm + k * o.nb + ou + c] = v
        hsum += h
        if s == 0:
            area = ar
    return (sigma_s * 0.5 * hsum * area)[:, None, None] * np.einsum("eai,ebi->eab", J, J)


# --------------------------------------------------------------------------------------
# global assembly
# --------------------------------------------------------------------------------------
def apply_dirichlet(indptr, indices, data, b, bc_dofs, bc_values):
    """dolfinx assemble_matrix(bcs) + apply_lifting + bc.set [dep-knowledge]: in place."""
    n = len(indptr) - 1
    is_bc = np.zeros(n, dtype=bool)
    is_bc[bc_dofs] = True
    g = np.zeros(n)
    g[bc_dofs] = bc_values
    rows = np.repeat(np.arange(n), np.diff(indptr))
    cols = indices.astype(np.int64)
    lift = np.where(is_bc[cols] & ~is_bc[rows], data * g[cols], 0.0)
    np.subtract.at(b, rows, lift)
    data[is_bc[rows] | is_bc[cols]] = 0.0
    data[(rows == cols) & is_bc[rows]] = 1.0
    b[bc_dofs] = bc_values


def assemble_interface_elasticity(x, cells, phi, f, cell_tags, facet_tags, c2f, f2c, ds100, ds101, mat=None, gamma=1.0,
                                  sigma_s=1.0, method="closed_form", kphi=1, phi_dofmap=None, bc_dofs=None,
                                  bc_values=None, nquad=6):
    """(indptr, indices, data, b) of the operator above in box mode on the mixed numbering of the header."""
    mat = mat or Material()
    d = x.shape[1]
    nvtx = len(x)
    o = Offsets(d)
    mixed = mixed_dofmap(cells, d)
    n_rows = o.nb * nvtx
    nm = mixed.shape[1]
    allc = np.nonzero(np.isin(cell_tags, (1, 2, 3)))[0]
    interior = f2c[:, 1] >= 0
    f3 = np.nonzero((facet_tags == 3) & interior)[0]
    f4 = np.nonzero((facet_tags == 4) & interior)[0]
    indptr, indices = sparsity_pattern(n_rows, mixed, allc, np.concatenate([f3, f4]), f2c)
    data = np.zeros(len(indices))
    b = np.zeros(n_rows)
    tg = cell_tags[allc]
    if method == "closed_form":
        assert kphi == 1
        A, be = cell_tensors_closed_form(x, cells[allc], phi, f, tg, mat, gamma, sigma_s)
    else:
        pdm = cells if phi_dofmap is None else phi_dofmap
        A, be = cell_tensors_quadrature(x, cells[allc], phi[pdm[allc]], f[cells[allc]], tg, mat, gamma, sigma_s, kphi,
                                        n=nquad)
    dm = mixed[allc]
    _scatter(indptr, indices, data, np.repeat(dm, nm, axis=1).ravel(), np.tile(dm, (1, nm)).ravel(), A.ravel())
    np.add.at(b, dm.ravel(), be.ravel())
    bt = boundary_tensors_closed_form if method == "closed_form" else boundary_tensors_quadrature
    ft = facet_tensors_closed_form if method == "closed_form" else facet_tensors_quadrature
    for ents, side in ((ds100, "in"), (ds101, "out")):
        ents = np.asarray(ents).reshape(-1, 2)
        if len(ents):
            dmb = mixed[ents[:, 0]]
            _scatter(indptr, indices, data, np.repeat(dmb, nm, axis=1).ravel(), np.tile(dmb, (1, nm)).ravel(),
                     bt(x, cells, ents, side).ravel())
    for fac, side in ((f3, "in"), (f4, "out")):
        if len(fac):
            mac = np.concatenate([mixed[f2c[fac, 0]], mixed[f2c[fac, 1]]], axis=1)
            _scatter(indptr, indices, data, np.repeat(mac, 2 * nm, axis=1).ravel(),
                     np.tile(mac, (1, 2 * nm)).ravel(), ft(x, cells, c2f, f2c, fac, side, mat, sigma_s).ravel())
    if bc_dofs is not None and len(bc_dofs):
        apply_dirichlet(indptr, indices, data, b, np.asarray(bc_dofs), np.asarray(bc_values))
    return indptr, indices, data, b
