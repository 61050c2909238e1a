"""CPU oracle, part 2: the strong-Dirichlet phi-FEM operator in CSR (TEST INFRASTRUCTURE ONLY).

Restates, on arrays, what dolfinx assembles for the forms of the reference demo
demo/strong-dirichlet/flower/main.py:

  a(w,v) =  int_{dx(1,2)} grad(phi w).grad(phi v)                       (:105)
          - int_{ds(100)} (grad(phi w).n) phi v                         (:106)
          + sigma h_T^2 int_{dx(2)} div grad(phi w) div grad(phi v)     (:107-112)
          + sigma int_{dS(2,3)} avg(h_T) jump(grad(phi w),n) jump(grad(phi v),n)   (:113-118)
  L(v)   =  int_{dx(1,2)} f phi v - sigma h_T^2 int_{dx(2)} f div grad(phi v)      (:126-128)

PARITY PARTIALLY PINNED: the reference holds no golden matrix/vector and dolfinx/FFCx/PETSc are not
installable here, so no dolfinx-produced number backs this file.  What does: the element tensors of every operator
below (strong / weak Dirichlet, Neumann / Robin) -- both restatements -- are held to exact sympy integrals of the
literal integrands (tests/test_oracle_sympy.py); the dolfinx conventions of the global assembly ('+' side of dS, the
cell owning ds(100), the pattern) stay unpinned until baseline/dolfinx_reference.py has been run.
Two independent restatements live in this file and must agree to ~1e-14:
  * `*_closed_form`  -- exact element tensors for P1 phi, P1 w/v, P1 f;
  * `*_quadrature`   -- brute-force Gauss-Jacobi quadrature (exact to degree 11) for P1 or P2.
Matrix convention: row = test dof (v), column = trial dof (w), like PETSc's assembled A.
Sparsity pattern: union over the integral domains, structural zeros kept, rows without
contributions empty (dolfinx create_sparsity_pattern [dep-knowledge, SURVEY.md C.3]).
"""
import math

import numpy as np
import scipy.sparse as sp
from scipy.special import roots_jacobi


# P2 local edge order of dolfinx/basix [dep-knowledge, SURVEY.md C.7]
P2_EDGES = {2: ((1, 2), (0, 2), (0, 1)),
            3: ((2, 3), (1, 3), (1, 2), (0, 3), (0, 2), (0, 1))}


# --------------------------------------------------------------------------------------
# geometry
# --------------------------------------------------------------------------------------
def simplex_geometry(x, cells):
    """G [Nc, d+1, d] = grad(lambda_i), vol [Nc] = |K|, h [Nc] = CellDiameter (max vertex distance)."""
    xc = x[cells]
    d = x.shape[1]
    J = np.stack([xc[:, k + 1] - xc[:, 0] for k in range(d)], axis=2)  # columns = edge vectors
    det = np.linalg.det(J)
    Jinv = np.linalg.inv(J)            # rows of J^-1 are grad(lambda_{k+1})
    G = np.empty((len(cells), d + 1, d))
    G[:, 1:, :] = Jinv
    G[:, 0, :] = -Jinv.sum(axis=1)
    vol = np.abs(det) / math.factorial(d)
    h = np.zeros(len(cells))
    for a in range(d + 1):
        for b in range(a + 1, d + 1):
            h = np.maximum(h, np.sqrt(((xc[:, a] - xc[:, b]) ** 2).sum(axis=1)))
    return G, vol, h


def facet_geometry(x, cells, ents):
    """Outward unit normal [m, d] and measure |F| [m] of (cell, local facet) pairs
    (local facet i opposite vertex i => n = -G_i/|G_i|, |F| = d |K| |G_i|)."""
    ents = np.asarray(ents).reshape(-1, 2)
    G, vol, _ = simplex_geometry(x, cells[ents[:, 0]])
    d = x.shape[1]
    Go = G[np.arange(len(ents)), ents[:, 1]]
    norm = np.sqrt((Go ** 2).sum(axis=1))
    return -Go / norm[:, None], d * vol * norm


# --------------------------------------------------------------------------------------
# closed forms, P1 everywhere
# --------------------------------------------------------------------------------------
def cell_tensors_closed_form(x, cells, phi, f, cut, sigma):
    """Element matrix [n, d+1, d+1] and vector [n, d+1] of cells (P1 dofs = vertices).
    `cut` [n] bool adds the stabilisation terms of dx(2)."""
    d = x.shape[1]
    nv = d + 1
    G, vol, h = simplex_geometry(x, cells)
    p = phi[cells]                                   # [n, nv]
    fv = f[cells]
    g = np.einsum("nk,nkd->nd", p, G)                # grad(phi_h)
    gg = (g * g).sum(axis=1)
    a = np.einsum("nd,nid->ni", g, G)                # g . G_i
    D = np.einsum("nid,njd->nij", G, G)              # G_i . G_j
    cM = vol / ((d + 1) * (d + 2))
    M = cM[:, None, None] * (1.0 + np.eye(nv))[None]
    m = np.einsum("nk,nki->ni", p, M)                # int lambda_i phi
    mu = (p * m).sum(axis=1)                         # int phi^2
    A = (gg[:, None, None] * M + a[:, :, None] * m[:, None, :] + m[:, :, None] * a[:, None, :]
         + D * mu[:, None, None])
    stab = np.where(cut, sigma * h * h * vol, 0.0)
    A = A + (4.0 * stab)[:, None, None] * a[:, :, None] * a[:, None, :]
    c3 = vol * math.factorial(d) / math.factorial(d + 3)
    F, P, FP = fv.sum(axis=1), p.sum(axis=1), (fv * p).sum(axis=1)
    b = c3[:, None] * ((F * P + FP)[:, None] + fv * P[:, None] + F[:, None] * p + 2.0 * fv * p)
    b = b - (2.0 * stab * fv.mean(axis=1))[:, None] * a
    return A, b


def boundary_tensors_closed_form(x, cells, phi, ents):
    """-int_F (grad(phi w).n) phi v over (cell, local facet) pairs -> [m, d+1, d+1] (row = test)."""
    ents = np.asarray(ents).reshape(-1, 2)
    d = x.shape[1]
    nv = d + 1
    cs = cells[ents[:, 0]]
    G, vol, _ = simplex_geometry(x, cs)
    n, area = facet_geometry(x, cells, ents)
    p = phi[cs]
    g = np.einsum("nk,nkd->nd", p, G)
    gn = (g * n).sum(axis=1)
    Gn = np.einsum("njd,nd->nj", G, n)
    cF = area * math.factorial(d - 1) / math.factorial(d + 2)
    A = np.zeros((len(ents), nv, nv))
    for e in range(len(ents)):
        o = ents[e, 1]
        on = [k for k in range(nv) if k != o]
        for i in on:
            for j in range(nv):
                acc = 0.0
                for k in on:
                    t_jki = _alpha((j, k, i)) if j != o else 0.0
                    s = 0.0
                    for l in on:
                        s += p[e, l] * _alpha((l, k, i))
                    acc += p[e, k] * (gn[e] * t_jki + Gn[e, j] * s)
                A[e, i, j] = -cF[e] * acc
    return A


def _alpha(idx):
    out = 1
    for v in set(idx):
        out *= math.factorial(idx.count(v))
    return float(out)


def ghost_tensors_closed_form(x, cells, phi, c2f, f2c, facets, sigma):
    """sigma avg(h) int_F jump(grad(phi w),n) jump(grad(phi v),n) -> macro matrices
    [ng, 2(d+1), 2(d+1)] over dofs [cell+ vertices, cell- vertices], + = first cell of the facet."""
    d = x.shape[1]
    nv = d + 1
    facets = np.asarray(facets)
    E = np.zeros((len(facets), 2 * nv, 2 * nv))
    macro = np.zeros((len(facets), 2 * nv), dtype=np.int64)
    if len(facets) == 0:
        return E, macro
    cp, cm = f2c[facets, 0], f2c[facets, 1]
    hs = []
    Jv = np.zeros((len(facets), 2 * nv, d))  # value of J_a at the d facet vertices (ordered as in cell +)
    area = None
    fverts = None
    for side, cc in enumerate((cp, cm)):
        lf = np.argmax(c2f[cc] == facets[:, None], axis=1)
        ents = np.stack([cc, lf], axis=1)
        G, vol, h = simplex_geometry(x, cells[cc])
        n, ar = facet_geometry(x, cells, ents)
        hs.append(h)
        p = phi[cells[cc]]
        g = np.einsum("nk,nkd->nd", p, G)
        gn = (g * n).sum(axis=1)
        Gn = np.einsum("njd,nd->nj", G, n)
        macro[:, side * nv:(side + 1) * nv] = cells[cc]
        if side == 0:
            area = ar
            fverts = np.stack([np.delete(cells[c], l) for c, l in zip(cc, lf)])  # [ng, d]
        for e in range(len(facets)):
            for a in range(nv):
                va = cells[cc[e], a]
                for mi in range(d):
                    vm = fverts[e, mi]
                    Jv[e, side * nv + a, mi] = (gn[e] if va == vm else 0.0) + Gn[e, a] * phi[vm]
    MF = (area / (d * (d + 1)))[:, None, None] * (1.0 + np.eye(d))[None]
    coef = sigma * 0.5 * (hs[0] + hs[1])
    E = coef[:, None, None] * np.einsum("eam,emn,ebn->eab", Jv, MF, Jv)
    return E, macro


# --------------------------------------------------------------------------------------
# brute-force quadrature, P1 or P2 (independent restatement)
# --------------------------------------------------------------------------------------
def simplex_rule(d, n=6):
    """Collapsed Gauss-Jacobi rule on the reference d-simplex, exact to degree 2n-1.
    Returns barycentric points [nq, d+1] and weights summing to 1/d!."""
    if d == 1:
        t, w = roots_jacobi(n, 0, 0)
        t = 0.5 * (t + 1)
        return np.stack([1 - t, t], axis=1), 0.5 * w
    if d == 2:
        a, wa = roots_jacobi(n, 0, 0)
        b, wb = roots_jacobi(n, 1, 0)
        a, b = 0.5 * (a + 1), 0.5 * (b + 1)
        X = (a[:, None] * (1 - b[None, :])).ravel()
        Y = np.repeat(b[None, :], n, axis=0).ravel()
        W = (wa[:, None] * wb[None, :]).ravel() / 8.0
        return np.stack([1 - X - Y, X, Y], axis=1), W
    a, wa = roots_jacobi(n, 0, 0)
    b, wb = roots_jacobi(n, 1, 0)
    c, wc = roots_jacobi(n, 2, 0)
    a, b, c = 0.5 * (a + 1), 0.5 * (b + 1), 0.5 * (c + 1)
    A, B, C = np.meshgrid(a, b, c, indexing="ij")
    X = (A * (1 - B) * (1 - C)).ravel()
    Y = (B * (1 - C)).ravel()
    Z = C.ravel()
    W = (wa[:, None, None] * wb[None, :, None] * wc[None, None, :]).ravel() / 64.0
    return np.stack([1 - X - Y - Z, X, Y, Z], axis=1), W


def lagrange_eval(lam, G, degree):
    """Values [nq, nd], gradients [nq, nd, d] and Laplacians [nd] (constant for k<=2) of the P_k
    Lagrange basis of ONE simplex at barycentric points lam [nq, d+1]; G [d+1, d]."""
    nq, nv = lam.shape
    d = nv - 1
    if degree == 1:
        return lam.copy(), np.repeat(G[None], nq, axis=0), np.zeros(nv)
    edges = P2_EDGES[d]
    nd = nv + len(edges)
    val = np.empty((nq, nd))
    grad = np.empty((nq, nd, d))
    lap = np.empty(nd)
    for i in range(nv):
        val[:, i] = lam[:, i] * (2 * lam[:, i] - 1)
        grad[:, i, :] = (4 * lam[:, i] - 1)[:, None] * G[i][None, :]
        lap[i] = 4.0 * G[i] @ G[i]
    for e, (a, b) in enumerate(edges):
        val[:, nv + e] = 4 * lam[:, a] * lam[:, b]
        grad[:, nv + e, :] = 4 * (lam[:, a][:, None] * G[b][None] + lam[:, b][:, None] * G[a][None])
        lap[nv + e] = 8.0 * G[a] @ G[b]
    return val, grad, lap


def _phi_w_derivs(lam, G, pc, kphi, kw):
    """For u_j = phi * psi_j on one simplex: values, gradients, Laplacians at the points."""
    pv, pg, pl = lagrange_eval(lam, G, kphi)
    wv, wg, wl = lagrange_eval(lam, G, kw)
    ph = pv @ pc                                   # [nq]
    gph = np.einsum("qkd,k->qd", pg, pc)            # [nq, d]
    lph = pl @ pc                                  # scalar
    u = ph[:, None] * wv                           # [nq, nd]
    gu = wv[:, :, None] * gph[:, None, :] + ph[:, None, None] * wg
    lu = wv * lph + 2 * np.einsum("qd,qjd->qj", gph, wg) + ph[:, None] * wl[None, :]
    return u, gu, lu, wv


def cell_tensors_quadrature(x, cells, phi_dofs, f_dofs, cut, sigma, kphi=1, kw=1, n=6):
    """phi_dofs [n, nd_phi], f_dofs [n, nd_w]: cell-local coefficient arrays."""
    d = x.shape[1]
    G, vol, h = simplex_geometry(x, cells)
    lam, W = simplex_rule(d, n)
    out_A, out_b = [], []
    for c in range(len(cells)):
        u, gu, lu, wv = _phi_w_derivs(lam, G[c], phi_dofs[c], kphi, kw)
        wq = W * math.factorial(d) * vol[c]
        A = np.einsum("q,qid,qjd->ij", wq, gu, gu)
        fq = wv @ f_dofs[c]
        b = np.einsum("q,q,qi->i", wq, fq, u)
        if cut[c]:
            s = sigma * h[c] ** 2
            A = A + s * np.einsum("q,qi,qj->ij", wq, lu, lu)
            b = b - s * np.einsum("q,q,qi->i", wq, fq, lu)
        out_A.append(A)
        out_b.append(b)
    return np.array(out_A), np.array(out_b)


def _facet_bary(d, o, n):
    """Quadrature on local facet o (opposite vertex o) as barycentric points of the cell."""
    flam, W = simplex_rule(d - 1, n)
    lam = np.zeros((len(W), d + 1))
    lam[:, [k for k in range(d + 1) if k != o]] = flam
    return lam, W * math.factorial(d - 1)


def boundary_tensors_quadrature(x, cells, phi_dofs, ents, kphi=1, kw=1, n=6):
    ents = np.asarray(ents).reshape(-1, 2)
    d = x.shape[1]
    G, _, _ = simplex_geometry(x, cells[ents[:, 0]])
    nrm, area = facet_geometry(x, cells, ents)
    out = []
    for e, (c, o) in enumerate(ents):
        lam, W = _facet_bary(d, o, n)
        u, gu, _, _ = _phi_w_derivs(lam, G[e], phi_dofs[e], kphi, kw)
        out.append(-np.einsum("q,qj,qi->ij", W * area[e], gu @ nrm[e], u))
    return np.array(out)


def ghost_tensors_quadrature(x, cells, phi_dofs_plus, phi_dofs_minus, c2f, f2c, facets, sigma,
                             kphi=1, kw=1, n=6):
    """Macro matrices over [dofs of cell +, dofs of cell -] (cell-local P_k dof order)."""
    d = x.shape[1]
    facets = np.asarray(facets)
    out = []
    for e, fct in enumerate(facets):
        cc = f2c[fct]
        J = []
        hsum = 0.0
        lam_plus_phys = None
        for side, c in enumerate(cc):
            o = int(np.nonzero(c2f[c] == fct)[0][0])
            G, vol, h = simplex_geometry(x, cells[c:c + 1])
            nrm, area = facet_geometry(x, cells, [(c, o)])
            lam, W = _facet_bary(d, o, n)
            if side == 0:
                lam_plus_phys = lam @ x[cells[c]]
                Wp, ar = W, area[0]
            else:
                # same physical points: barycentric coordinates of cell - at the points of cell +
                T = np.concatenate([x[cells[c]].T, np.ones((1, d + 1))], axis=0)
                rhs = np.concatenate([lam_plus_phys.T, np.ones((1, len(Wp)))], axis=0)
                lam = np.linalg.solve(T, rhs).T
            pd = (phi_dofs_plus if side == 0 else phi_dofs_minus)[e]
            _, gu, _, _ = _phi_w_derivs(lam, G[0], pd, kphi, kw)
            J.append(gu @ nrm[0])          # [nq, nd]
            hsum += h[0]
        Jm = np.concatenate(J, axis=1)     # [nq, 2 nd]
        out.append(sigma * 0.5 * hsum * np.einsum("q,qa,qb->ab", Wp * ar, Jm, Jm))
    return np.array(out)


# --------------------------------------------------------------------------------------
# global assembly
# --------------------------------------------------------------------------------------
def sparsity_pattern(n_rows, dofmap, active_cells, ghost_facets, f2c):
    """CSR pattern (indptr int32 [n_rows+1], indices int32 sorted per row): all dof pairs of every
    active cell, all pairs among the union of both cells' dofs for every ghost facet."""
    nd = dofmap.shape[1]
    dm = dofmap[active_cells].astype(np.int64)
    rows = [np.repeat(dm, nd, axis=1).ravel()]
    cols = [np.tile(dm, (1, nd)).ravel()]
    if len(ghost_facets):
        mac = np.concatenate([dofmap[f2c[ghost_facets, 0]], dofmap[f2c[ghost_facets, 1]]],
                             axis=1).astype(np.int64)
        rows.append(np.repeat(mac, 2 * nd, axis=1).ravel())
        cols.append(np.tile(mac, (1, 2 * nd)).ravel())
    key = np.unique(np.concatenate(rows) * n_rows + np.concatenate(cols))
    r, c = key // n_rows, key % n_rows
    indptr = np.zeros(n_rows + 1, dtype=np.int64)
    np.add.at(indptr, r + 1, 1)
    return np.cumsum(indptr).astype(np.int32), c.astype(np.int32)


def _scatter(indptr, indices, data, rows, cols, vals):
    n = len(indptr) - 1
    key = indptr[:-1].astype(np.int64)  # row starts
    # position of (row, col) inside the row by binary search on the sorted column list
    allkeys = np.repeat(np.arange(n, dtype=np.int64), np.diff(indptr)) * n + indices
    pos = np.searchsorted(allkeys, rows.astype(np.int64) * n + cols)
    assert np.array_equal(allkeys[pos], rows.astype(np.int64) * n + cols), "entry outside the pattern"
    np.add.at(data, pos, vals)
    del key


def assemble_strong_dirichlet(x, cells, dofmap, n_rows, phi, f, cell_tags, facet_tags, c2f, f2c,
                              ds100, sigma=1.0, method="closed_form", kphi=1, kw=1,
                              phi_dofmap=None):
    """Assemble (indptr, indices, data, b) of the strong-Dirichlet operator in box mode.
    P1 closed forms need dofmap == cells (vertex dofs); the quadrature method takes P1 or P2
    dofmaps (`dofmap` for w/v/f, `phi_dofmap` for phi)."""
    phi_dofmap = dofmap if phi_dofmap is None else phi_dofmap
    active = np.nonzero((cell_tags == 1) | (cell_tags == 2))[0]
    ghost = np.nonzero(((facet_tags == 2) | (facet_tags == 3)) & (f2c[:, 1] >= 0))[0]
    ents = np.asarray(ds100).reshape(-1, 2)
    indptr, indices = sparsity_pattern(n_rows, dofmap, active, ghost, f2c)
    data = np.zeros(len(indices))
    b = np.zeros(n_rows)
    nd = dofmap.shape[1]
    cut = cell_tags[active] == 2
    if method == "closed_form":
        assert kphi == 1 and kw == 1
        A, be = cell_tensors_closed_form(x, cells[active], phi, f, cut, sigma)
        Ab = boundary_tensors_closed_form(x, cells, phi, ents)
        Eg, _ = ghost_tensors_closed_form(x, cells, phi, c2f, f2c, ghost, sigma)
    else:
        A, be = cell_tensors_quadrature(x, cells[active], phi[phi_dofmap[active]], f[dofmap[active]],
                                        cut, sigma, kphi, kw)
        Ab = boundary_tensors_quadrature(x, cells, phi[phi_dofmap[ents[:, 0]]], ents, kphi, kw)
        Eg = ghost_tensors_quadrature(x, cells, phi[phi_dofmap[f2c[ghost, 0]]],
                                      phi[phi_dofmap[f2c[ghost, 1]]], c2f, f2c, ghost, sigma, kphi, kw)
    dm = dofmap[active].astype(np.int64)
    _scatter(indptr, indices, data, np.repeat(dm, nd, axis=1).ravel(), np.tile(dm, (1, nd)).ravel(),
             A.ravel())
    np.add.at(b, dm.ravel(), be.ravel())
    if len(ents):
        dmb = dofmap[ents[:, 0]].astype(np.int64)
        _scatter(indptr, indices, data, np.repeat(dmb, nd, axis=1).ravel(),
                 np.tile(dmb, (1, nd)).ravel(), Ab.ravel())
    if len(ghost):
        mac = np.concatenate([dofmap[f2c[ghost, 0]], dofmap[f2c[ghost, 1]]], axis=1).astype(np.int64)
        _scatter(indptr, indices, data, np.repeat(mac, 2 * nd, axis=1).ravel(),
                 np.tile(mac, (1, 2 * nd)).ravel(), Eg.ravel())
    return indptr, indices, data, b


def to_scipy(indptr, indices, data, n):
    return sp.csr_matrix((data, indices, indptr), shape=(n, n))


# --------------------------------------------------------------------------------------
# weak-Dirichlet (dual) phi-FEM operator, mixed P_k x P_k space (u, p)
# reference demo/weak-dirichlet/flower/main.py:102-154
#   a = int_{dx(1,2)} grad u.grad v - int_ds (grad u.n) v                                   (:113-114)
#       + gamma h^-2 int_{dx(2)} (u - h^-1 phi p)(v - h^-1 phi q)                           (:115-122)
#       + sigma h^2 int_{dx(2)} lap u lap v + sigma int_{dS(2,3)} avg(h) [grad u.n][grad v.n]   (:123-134)
#   L = int_{dx(1,2)} f v + gamma h^-2 int_{dx(2)} u_D (v - h^-1 phi q) - sigma h^2 int_{dx(2)} f lap v   (:142-151)
# Cell-local mixed dof order: [u dofs of the cell, p dofs of the cell] (sub-element 0 first, as a
# basix mixed element lays them out [dep-knowledge]); the global numbering is an input (`mixed_dofmap`).
# Two independent restatements: brute-force quadrature (P1 / P2) and closed forms for P1 from the exact
# integrals of barycentric monomials, int prod lambda^e = |K| d! prod e! / (d + sum e)!.
# --------------------------------------------------------------------------------------
def weak_cell_tensors_quadrature(x, cells, phi_dofs, f_dofs, ud_dofs, cut, gamma, sigma, kphi=1, kw=1, n=6):
    """[n, 2nd, 2nd] matrices and [n, 2nd] vectors; phi_dofs [n, nd_phi], f_dofs / ud_dofs [n, nd]."""
    d = x.shape[1]
    G, vol, h = simplex_geometry(x, cells)
    lam, W = simplex_rule(d, n)
    out_A, out_b = [], []
    for c in range(len(cells)):
        wv, wg, wl = lagrange_eval(lam, G[c], kw)
        pv, _, _ = lagrange_eval(lam, G[c], kphi)
        nd = wv.shape[1]
        ph = pv @ phi_dofs[c]
        wq = W * math.factorial(d) * vol[c]
        fq = wv @ f_dofs[c]
        A = np.zeros((2 * nd, 2 * nd))
        b = np.zeros(2 * nd)
        A[:nd, :nd] = np.einsum("q,qid,qjd->ij", wq, wg, wg)
        b[:nd] = np.einsum("q,q,qi->i", wq, fq, wv)
        if cut[c]:
            hh = h[c]
            # test / trial combination  v - h^-1 phi q  over the mixed basis [psi_i, 0] and [0, psi_i]
            T = np.concatenate([wv, -(ph / hh)[:, None] * wv], axis=1)          # [nq, 2nd]
            A += gamma / hh ** 2 * np.einsum("q,qa,qb->ab", wq, T, T)
            A[:nd, :nd] += sigma * hh ** 2 * np.einsum("q,i,j->ij", wq, wl, wl)
            uq = wv @ ud_dofs[c]
            b += gamma / hh ** 2 * np.einsum("q,q,qa->a", wq, uq, T)
            b[:nd] -= sigma * hh ** 2 * np.einsum("q,q,i->i", wq, fq, wl)
        out_A.append(A)
        out_b.append(b)
    return np.array(out_A), np.array(out_b)


def weak_boundary_tensors_quadrature(x, cells, ents, kw=1, n=6):
    """-int_F (grad u_j.n) v_i on (cell, local facet) pairs -> [m, 2nd, 2nd] (only the uu block is non-zero)."""
    ents = np.asarray(ents).reshape(-1, 2)
    d = x.shape[1]
    G, _, _ = simplex_geometry(x, cells[ents[:, 0]])
    nrm, area = facet_geometry(x, cells, ents)
    out = []
    for e, (c, o) in enumerate(ents):
        lam, W = _facet_bary(d, o, n)
        wv, wg, _ = lagrange_eval(lam, G[e], kw)
        nd = wv.shape[1]
        A = np.zeros((2 * nd, 2 * nd))
        A[:nd, :nd] = -np.einsum("q,qj,qi->ij", W * area[e], wg @ nrm[e], wv)
        out.append(A)
    return np.array(out)


def weak_ghost_tensors_quadrature(x, cells, c2f, f2c, facets, sigma, kw=1, n=6):
    """sigma avg(h) int_F [grad u.n][grad v.n] over macro dofs [mixed dofs of cell +, mixed dofs of cell -]
    -> [ng, 4nd, 4nd]; only u-u entries are non-zero."""
    d = x.shape[1]
    facets = np.asarray(facets)
    out = []
    for fct in facets:
        cc = f2c[fct]
        J, hsum = [], 0.0
        for side, c in enumerate(cc):
            o = int(np.nonzero(c2f[c] == fct)[0][0])
            G, vol, h = simplex_geometry(x, cells[c:c + 1])
            nrm, area = facet_geometry(x, cells, [(c, o)])
            lam, W = _facet_bary(d, o, n)
            if side == 0:
                xq = lam @ x[cells[c]]
                Wp, ar = W, area[0]
            else:
                T = np.concatenate([x[cells[c]].T, np.ones((1, d + 1))], axis=0)
                rhs = np.concatenate([xq.T, np.ones((1, len(Wp)))], axis=0)
                lam = np.linalg.solve(T, rhs).T
            _, wg, _ = lagrange_eval(lam, G[0], kw)
            Ju = wg @ nrm[0]                                   # [nq, nd]
            J.append(np.concatenate([Ju, np.zeros_like(Ju)], axis=1))
            hsum += h[0]
        Jm = np.concatenate(J, axis=1)                         # [nq, 4nd]
        out.append(sigma * 0.5 * hsum * np.einsum("q,qa,qb->ab", Wp * ar, Jm, Jm))
    return np.array(out)


def _bary_moment(d, idx):
    """int lambda_{idx[0]} lambda_{idx[1]} ... / |K|."""
    e = [idx.count(v) for v in set(idx)]
    return math.factorial(d) * math.prod(math.factorial(v) for v in e) / math.factorial(d + len(idx))


def weak_cell_tensors_closed_form(x, cells, phi, f, ud, cut, gamma, sigma):
    """P1 x P1 element tensors, vertex dofs; [n, 2nv, 2nv], [n, 2nv]."""
    d = x.shape[1]
    nv = d + 1
    G, vol, h = simplex_geometry(x, cells)
    p, fv, uv = phi[cells], f[cells], ud[cells]
    M = np.array([[_bary_moment(d, (i, j)) for j in range(nv)] for i in range(nv)])
    T3 = np.array([[[_bary_moment(d, (k, i, j)) for j in range(nv)] for i in range(nv)] for k in range(nv)])
    Q4 = np.array([[[[_bary_moment(d, (k, l, i, j)) for j in range(nv)] for i in range(nv)]
                    for l in range(nv)] for k in range(nv)])
    n = len(cells)
    A = np.zeros((n, 2 * nv, 2 * nv))
    b = np.zeros((n, 2 * nv))
    A[:, :nv, :nv] = vol[:, None, None] * np.einsum("nid,njd->nij", G, G)
    b[:, :nv] = vol[:, None] * (fv @ M)
    c = np.where(cut, 1.0, 0.0)
    g2 = c * gamma / h ** 2 * vol
    A[:, :nv, :nv] += g2[:, None, None] * M[None]
    up = -(g2 / h)[:, None, None] * np.einsum("nk,kij->nij", p, T3)     # (test v_i, trial p_j)
    A[:, :nv, nv:] += up
    A[:, nv:, :nv] += np.transpose(up, (0, 2, 1))
    A[:, nv:, nv:] += (g2 / h ** 2)[:, None, None] * np.einsum("nk,nl,klij->nij", p, p, Q4)
    b[:, :nv] += g2[:, None] * (uv @ M)
    b[:, nv:] += -(g2 / h)[:, None] * np.einsum("nk,nl,kli->ni", uv, p, T3)
    del sigma  # lap(P1) = 0: the h^2 lap.lap term vanishes (main.py:123-128)
    return A, b


def weak_boundary_tensors_closed_form(x, cells, ents):
    ents = np.asarray(ents).reshape(-1, 2)
    d = x.shape[1]
    nv = d + 1
    G, _, _ = simplex_geometry(x, cells[ents[:, 0]])
    nrm, area = facet_geometry(x, cells, ents)
    Gn = np.einsum("njd,nd->nj", G, nrm)
    A = np.zeros((len(ents), 2 * nv, 2 * nv))
    for e, (_, o) in enumerate(ents):
        for i in range(nv):
            if i != o:
                A[e, i, :nv] = -Gn[e] * area[e] / d
    return A


def weak_ghost_tensors_closed_form(x, cells, c2f, f2c, facets, sigma):
    d = x.shape[1]
    nv = d + 1
    facets = np.asarray(facets)
    E = np.zeros((len(facets), 4 * nv, 4 * nv))
    if len(facets) == 0:
        return E
    J = np.zeros((len(facets), 4 * nv))
    hsum = np.zeros(len(facets))
    area = None
    for side in (0, 1):
        cc = f2c[facets, side]
        lf = np.argmax(c2f[cc] == facets[:, None], axis=1)
        G, _, h = simplex_geometry(x, cells[cc])
        nrm, ar = facet_geometry(x, cells, np.stack([cc, lf], axis=1))
        J[:, side * 2 * nv:side * 2 * nv + nv] = np.einsum("njd,nd->nj", G, nrm)
        hsum += h
        if side == 0:
            area = ar
    return (sigma * 0.5 * hsum * area)[:, None, None] * J[:, :, None] * J[:, None, :]


def assemble_weak_dirichlet(x, cells, dofmap, n_scalar_dofs, phi, f, ud, cell_tags, facet_tags, c2f, f2c,
                            ds100, gamma=1.0, sigma=1.0, method="closed_form", kphi=1, kw=1,
                            phi_dofmap=None):
    """(indptr, indices, data, b) of the weak-Dirichlet operator in box mode.  `dofmap` [Nc, nd]: scalar P_kw
    dofmap; the mixed space numbers u at scalar dof s as 2 s and p as 2 s + 1 (cell-local order
    [u dofs, p dofs]).  Pattern: all mixed-dof pairs of every active cell and of every ghost-facet macro
    element (structural zeros kept: dolfinx builds the pattern from the mixed dofmap, not from the blocks that
    happen to be non-zero [dep-knowledge])."""
    phi_dofmap = dofmap if phi_dofmap is None else phi_dofmap
    mixed = np.concatenate([2 * dofmap.astype(np.int64), 2 * dofmap.astype(np.int64) + 1], axis=1)
    n_rows = 2 * n_scalar_dofs
    active = np.nonzero((cell_tags == 1) | (cell_tags == 2))[0]
    ghost = np.nonzero(((facet_tags == 2) | (facet_tags == 3)) & (f2c[:, 1] >= 0))[0]
    ents = np.asarray(ds100).reshape(-1, 2)
    indptr, indices = sparsity_pattern(n_rows, mixed, active, ghost, f2c)
    data = np.zeros(len(indices))
    b = np.zeros(n_rows)
    cut = cell_tags[active] == 2
    if method == "closed_form":
        assert kphi == 1 and kw == 1
        A, be = weak_cell_tensors_closed_form(x, cells[active], phi, f, ud, cut, gamma, sigma)
        Ab = weak_boundary_tensors_closed_form(x, cells, ents)
        Eg = weak_ghost_tensors_closed_form(x, cells, c2f, f2c, ghost, sigma)
    else:
        A, be = weak_cell_tensors_quadrature(x, cells[active], phi[phi_dofmap[active]], f[dofmap[active]],
                                             ud[dofmap[active]], cut, gamma, sigma, kphi, kw)
        Ab = weak_boundary_tensors_quadrature(x, cells, ents, kw)
        Eg = weak_ghost_tensors_quadrature(x, cells, c2f, f2c, ghost, sigma, kw)
    nm = mixed.shape[1]
    dm = mixed[active]
    _scatter(indptr, indices, data, np.repeat(dm, nm, axis=1).ravel(), np.tile(dm, (1, nm)).ravel(), A.ravel())
    np.add.at(b, dm.ravel(), be.ravel())
    if len(ents):
        dmb = mixed[ents[:, 0]]
        _scatter(indptr, indices, data, np.repeat(dmb, nm, axis=1).ravel(), np.tile(dmb, (1, nm)).ravel(),
                 Ab.ravel())
    if len(ghost):
        mac = np.concatenate([mixed[f2c[ghost, 0]], mixed[f2c[ghost, 1]]], axis=1)
        _scatter(indptr, indices, data, np.repeat(mac, 2 * nm, axis=1).ravel(),
                 np.tile(mac, (1, 2 * nm)).ravel(), Eg.ravel())
    return indptr, indices, data, b


# --------------------------------------------------------------------------------------
# Neumann phi-FEM operator on the mixed space (u, y, p) in P1 x P1^d x DG0
# reference demo/neumann/square/main.py:103-158 (BASELINE.json configs[1])
#   a = int_{dx(1,2)} (grad u.grad v + u v) + int_ds (y.n) v                                     (:119-120)
#       + gamma int_{dx(2)} [ (y + grad u).(z + grad v) + (div y + u)(div z + v)
#                             + h^-2 (y.grad phi + h^-1 p phi)(z.grad phi + h^-1 q phi) ]          (:121-135)
#       + sigma int_{dS(3)} avg(h) [grad u.n][grad v.n]                                          (:136-139)
#   L = int_{dx(1,2)} f v + gamma int_{dx(2)} [ -h^-2 u_N |grad phi| (z.grad phi + h^-1 q phi) + f (div z + v) ]
#                                                                                                (:146-158)
# Cell-local mixed dof order: [u at the nv vertices, y node-major (vertex i, component c -> nv + i d + c), p].
# Global numbering (an input of the kernels; this is the one phifem_b200 uses): u at vertex s -> (d+1) s,
# y_c at s -> (d+1) s + 1 + c, p of cell k -> (d+1) Nv + k.  With every basis function X carrying
# s1 = y + grad u, s2 = div y + u, s3 = y.grad phi + h^-1 p phi, the cut-cell integrand is
# s1_b.s1_a + s2_b s2_a + h^-2 s3_b s3_a.  |grad phi_h| is not polynomial for a P2 level set: the quadrature
# restatement is exact only for P1 phi (dolfinx would use the rule of UFL's estimated degree there).
# --------------------------------------------------------------------------------------
def neumann_mixed_dofmap(cells, n_vertices):
    d = cells.shape[1] - 1
    c = cells.astype(np.int64)
    u = (d + 1) * c
    y = ((d + 1) * c[:, :, None] + 1 + np.arange(d)[None, None, :]).reshape(len(c), -1)
    p = (d + 1) * n_vertices + np.arange(len(c), dtype=np.int64)[:, None]
    return np.concatenate([u, y, p], axis=1)


def _neumann_fields(lam, G, pc, kphi, h, kappa=0.0):
    """Per quadrature point and mixed basis function: u, grad u, s1, s2, s3; plus phi, grad phi.  kappa = the Robin
    coefficient (demo/robin/square/main.py:124-133: s3 = y.grad phi - |grad phi| kappa u + h^-1 p phi; 0 = Neumann)."""
    nq, nv = lam.shape
    d = nv - 1
    nm = nv * (1 + d) + 1
    pv, pg, _ = lagrange_eval(lam, G, kphi)
    ph = pv @ pc
    gph = np.einsum("qkd,k->qd", pg, pc)
    U = np.zeros((nq, nm))
    GU = np.zeros((nq, nm, d))
    S1 = np.zeros((nq, nm, d))
    S2 = np.zeros((nq, nm))
    S3 = np.zeros((nq, nm))
    ngp = np.sqrt((gph ** 2).sum(axis=1))
    for j in range(nv):
        U[:, j] = lam[:, j]
        GU[:, j, :] = G[j][None, :]
        S1[:, j, :] = G[j][None, :]
        S2[:, j] = lam[:, j]
        S3[:, j] = -kappa * ngp * lam[:, j]
        for c in range(d):
            k = nv + j * d + c
            S1[:, k, c] = lam[:, j]
            S2[:, k] = G[j, c]
            S3[:, k] = lam[:, j] * gph[:, c]
    S3[:, nm - 1] = ph / h
    return U, GU, S1, S2, S3, ph, gph


def neumann_cell_tensors_quadrature(x, cells, phi_dofs, f_dofs, un_dofs, cut, gamma, kphi=1, n=6, rule=None,
                                    kappa=0.0):
    """rule = (barycentric points, weights summing to 1): use this rule instead of the degree-11 one -- for a P2
    level set the load term holds |grad phi_h| (not polynomial), so the result depends on the rule."""
    d = x.shape[1]
    G, vol, h = simplex_geometry(x, cells)
    lam, W = simplex_rule(d, n)
    if rule is not None:
        lam, W = np.asarray(rule[0]), np.asarray(rule[1]) / math.factorial(d)
    out_A, out_b = [], []
    for c in range(len(cells)):
        U, GU, S1, S2, S3, ph, gph = _neumann_fields(lam, G[c], phi_dofs[c], kphi, h[c], kappa)
        wq = W * math.factorial(d) * vol[c]
        fq = lam @ f_dofs[c]
        A = np.einsum("q,qad,qbd->ab", wq, GU, GU) + np.einsum("q,qa,qb->ab", wq, U, U)
        b = np.einsum("q,q,qa->a", wq, fq, U)
        if cut[c]:
            hh = h[c]
            A += gamma * (np.einsum("q,qad,qbd->ab", wq, S1, S1) + np.einsum("q,qa,qb->ab", wq, S2, S2)
                          + np.einsum("q,qa,qb->ab", wq, S3, S3) / hh ** 2)
            unq = lam @ un_dofs[c]
            ngp = np.sqrt((gph ** 2).sum(axis=1))
            b += gamma * (-np.einsum("q,q,q,qa->a", wq, unq, ngp, S3) / hh ** 2 + np.einsum("q,q,qa->a", wq, fq, S2))
        out_A.append(A)
        out_b.append(b)
    return np.array(out_A), np.array(out_b)


def neumann_cell_tensors_closed_form(x, cells, phi, f, un, cut, gamma, kappa=0.0):
    """P1 level set: every integrand is polynomial (grad phi constant per cell); exact monomial integrals."""
    d = x.shape[1]
    nv = d + 1
    nm = nv * (1 + d) + 1
    G, vol, h = simplex_geometry(x, cells)
    p, fv, uv = phi[cells], f[cells], un[cells]
    M = np.array([[_bary_moment(d, (i, j)) for j in range(nv)] for i in range(nv)])
    m1 = 1.0 / nv
    n = len(cells)
    A = np.zeros((n, nm, nm))
    b = np.zeros((n, nm))
    GG = np.einsum("nid,njd->nij", G, G)
    g = np.einsum("nk,nkd->nd", p, G)
    ng = np.sqrt((g * g).sum(axis=1))
    cg = np.where(cut, gamma, 0.0)
    yi = lambda i, c: nv + i * d + c          # noqa: E731
    A[:, :nv, :nv] = (vol * (1.0 + cg))[:, None, None] * (GG + M[None])
    b[:, :nv] = (vol * (1.0 + cg))[:, None] * (fv @ M)
    fbar = fv.mean(axis=1)                     # int f / |K|
    for i in range(nv):
        for c in range(d):
            a = yi(i, c)
            # rows of z = psi_i e_c against u_j: s1 -> psi_i dG_j/dc, s2 -> dpsi_i/dc psi_j
            A[:, a, :nv] = (cg * vol)[:, None] * (m1 * G[:, :, c] + G[:, i, c][:, None] * m1)
            A[:, :nv, a] = A[:, a, :nv]
            for j in range(nv):
                for c2 in range(d):
                    a2 = yi(j, c2)
                    A[:, a, a2] = cg * vol * ((M[i, j] if c == c2 else 0.0) + G[:, i, c] * G[:, j, c2]
                                              + g[:, c] * g[:, c2] * M[i, j] / h ** 2)
            A[:, a, nm - 1] = cg * vol * g[:, c] * (p @ M[:, i]) / h ** 3
            A[:, nm - 1, a] = A[:, a, nm - 1]
            b[:, a] = cg * vol * (-ng * g[:, c] * (uv @ M[:, i]) / h ** 2 + G[:, i, c] * fbar)
    A[:, nm - 1, nm - 1] = cg * vol * np.einsum("nk,nl,kl->n", p, p, M) / h ** 4
    b[:, nm - 1] = -cg * vol * ng * np.einsum("nk,nl,kl->n", uv, p, M) / h ** 3
    if kappa != 0.0:   # Robin: s3 of u_j gains -kappa |g| lambda_j
        kg = kappa * ng
        A[:, :nv, :nv] += (cg * vol * kg * kg / h ** 2)[:, None, None] * M[None]
        for i in range(nv):
            for c in range(d):
                a = yi(i, c)
                t = -(cg * vol * kg * g[:, c] / h ** 2)[:, None] * M[i][None, :]
                A[:, a, :nv] += t
                A[:, :nv, a] += t
        t = -(cg * vol * kg / h ** 3)[:, None] * (p @ M)
        A[:, nm - 1, :nv] += t
        A[:, :nv, nm - 1] += t
        b[:, :nv] += (cg * vol * ng * kg / h ** 2)[:, None] * (uv @ M)
    return A, b


def neumann_boundary_tensors(x, cells, ents):
    """int_F (y.n) v on (cell, local facet) pairs: rows u_i (i on the facet), columns y_(j,c) (j on the facet):
    n_c |F| (1 + delta_ij) / (d (d+1))."""
    ents = np.asarray(ents).reshape(-1, 2)
    d = x.shape[1]
    nv = d + 1
    nm = nv * (1 + d) + 1
    nrm, area = facet_geometry(x, cells, ents)
    A = np.zeros((len(ents), nm, nm))
    for e, (_, o) in enumerate(ents):
        on = [k for k in range(nv) if k != o]
        for i in on:
            for j in on:
                for c in range(d):
                    A[e, i, nv + j * d + c] = nrm[e, c] * area[e] * (2.0 if i == j else 1.0) / (d * (d + 1))
    return A


def neumann_ghost_tensors(x, cells, c2f, f2c, facets, sigma):
    """sigma avg(h) int_F [grad u.n][grad v.n] over macro dofs [mixed dofs of cell +, of cell -]; u-u entries only."""
    d = x.shape[1]
    nv = d + 1
    nm = nv * (1 + d) + 1
    facets = np.asarray(facets)
    E = np.zeros((len(facets), 2 * nm, 2 * nm))
    if len(facets) == 0:
        return E
    J = np.zeros((len(facets), 2 * nm))
    hsum = np.zeros(len(facets))
    area = None
    for side in (0, 1):
        cc = f2c[facets, side]
        lf = np.argmax(c2f[cc] == facets[:, None], axis=1)
        G, _, h = simplex_geometry(x, cells[cc])
        nrm, ar = facet_geometry(x, cells, np.stack([cc, lf], axis=1))
        J[:, side * nm:side * nm + nv] = np.einsum("njd,nd->nj", G, nrm)
        hsum += h
        if side == 0:
            area = ar
    return (sigma * 0.5 * hsum * area)[:, None, None] * J[:, :, None] * J[:, None, :]


def assemble_neumann(x, cells, phi, f, un, cell_tags, facet_tags, c2f, f2c, ds100, gamma=1.0, sigma=1.0,
                     method="closed_form", kphi=1, phi_dofmap=None, rule=None, robin_coef=0.0, ghost_tag=3):
    """(indptr, indices, data, b) of the Neumann operator in box mode on the mixed numbering above.  robin_coef != 0
    with ghost_tag = 2 gives the Robin operator of demo/robin/square/main.py:118-174 (`un` = the Robin data u_R)."""
    nvtx = len(x)
    d = x.shape[1]
    mixed = neumann_mixed_dofmap(cells, nvtx)
    n_rows = (d + 1) * nvtx + len(cells)
    phi_dofmap = cells if phi_dofmap is None else phi_dofmap
    active = np.nonzero((cell_tags == 1) | (cell_tags == 2))[0]
    ghost = np.nonzero((facet_tags == ghost_tag) & (f2c[:, 1] >= 0))[0]
    ents = np.asarray(ds100).reshape(-1, 2)
    indptr, indices = sparsity_pattern(n_rows, mixed, active, ghost, f2c)
    data = np.zeros(len(indices))
    b = np.zeros(n_rows)
    cut = cell_tags[active] == 2
    if method == "closed_form":
        assert kphi == 1
        A, be = neumann_cell_tensors_closed_form(x, cells[active], phi, f, un, cut, gamma, robin_coef)
    else:
        A, be = neumann_cell_tensors_quadrature(x, cells[active], phi[phi_dofmap[active]], f[cells[active]],
                                                un[cells[active]], cut, gamma, kphi, rule=rule, kappa=robin_coef)
    Ab = neumann_boundary_tensors(x, cells, ents)
    Eg = neumann_ghost_tensors(x, cells, c2f, f2c, ghost, sigma)
    nm = mixed.shape[1]
    dm = mixed[active]
    _scatter(indptr, indices, data, np.repeat(dm, nm, axis=1).ravel(), np.tile(dm, (1, nm)).ravel(), A.ravel())
    np.add.at(b, dm.ravel(), be.ravel())
    if len(ents):
        dmb = mixed[ents[:, 0]]
        _scatter(indptr, indices, data, np.repeat(dmb, nm, axis=1).ravel(), np.tile(dmb, (1, nm)).ravel(),
                 Ab.ravel())
    if len(ghost):
        mac = np.concatenate([mixed[f2c[ghost, 0]], mixed[f2c[ghost, 1]]], axis=1)
        _scatter(indptr, indices, data, np.repeat(mac, 2 * nm, axis=1).ravel(),
                 np.tile(mac, (1, 2 * nm)).ravel(), Eg.ravel())
    return indptr, indices, data, b
