"""CPU oracle, part 1: level-set cut-cell classification (TEST INFRASTRUCTURE ONLY).

numpy restatement of reference src/phifem/mesh_scripts.py.  Works on plain arrays:

  x      [Nv, gdim] float64      vertex coordinates
  cells  [Nc, nvpc] int          cell -> vertex, dolfinx local vertex order
                                 (quadrilaterals in tensor order v0=(0,0) v1=(1,0) v2=(0,1) v3=(1,1))

The level set reaches the oracle as point values:
  phi_cell   [Nc, npts]          phi at the detection points of each cell  (mesh_scripts.py:320-340)
  phi_facet  [Nc, nfpc, nq]      phi at the detection points of each local facet (mesh_scripts.py:434-447)
`point_values_*` below produce them from a P_k coefficient vector + tabulated basis
(FFCx-style sequential sums) or from a callable of the physical coordinates
(the UFL-expression mode of reference tests/test_compute_meshtags.py:160-161).

Conventions taken from dolfinx 0.9 [dep-knowledge, SURVEY.md Appendix C]: simplex local
facet i is opposite local vertex i; quadrilateral facets (v0v1),(v0v2),(v1v3),(v2v3);
serial facet numbering = lexicographic rank of the sorted vertex tuple; the cells of a
facet are listed in ascending order.  Tetrahedra are OUR extension (the reference raises
NotImplementedError, mesh_scripts.py:326-329).

All floating-point sums are sequential, left to right, without FMA contraction
(numpy elementwise ops never contract), mirroring the FFCx-generated C.
"""
import warnings

import numpy as np

LOCAL_FACETS = {
    "triangle": ((1, 2), (0, 2), (0, 1)),
    "quadrilateral": ((0, 1), (0, 2), (1, 3), (2, 3)),
    "tetrahedron": ((1, 2, 3), (0, 2, 3), (0, 1, 3), (0, 1, 2)),
}
TDIM = {"triangle": 2, "quadrilateral": 2, "tetrahedron": 3}


# --------------------------------------------------------------------------------------
# detection points  (mesh_scripts.py:28-92)
# --------------------------------------------------------------------------------------
def segment_points(N):
    """mesh_scripts.py:28-40: N+1 equispaced points on [0,1]; N=0 -> midpoint."""
    if N <= 0:
        return np.array([[0.5]])
    return np.linspace(0, 1, N + 1).astype(np.float64).reshape(-1, 1)


def triangle_boundary_points(N):
    """mesh_scripts.py:43-65: walk the boundary v0->v1->v2->v0; 3N points (N=1: the 3 vertices)."""
    if N <= 0:
        return np.array([[1.0 / 3.0, 1.0 / 3.0]])
    t = np.linspace(0, 1, N + 1)
    pts = [(ti, 0.0) for ti in t]              # edge v0->v1, both ends
    pts += [(1 - ti, ti) for ti in t[1:]]      # edge v1->v2, without its start
    if N > 1:
        pts += [(0.0, 1 - ti) for ti in t[1:-1]]  # edge v2->v0, without both ends
    return np.array(pts, dtype=np.float64)


def square_boundary_points(N):
    """mesh_scripts.py:68-92: walk (0,0)->(1,0)->(1,1)->(0,1)->(0,0); 4N points (N=1: 4 corners)."""
    if N <= 0:
        return np.array([[0.5, 0.5]])
    t = np.linspace(0, 1, N + 1)
    pts = [(ti, 0.0) for ti in t]
    pts += [(1.0, ti) for ti in t[1:]]
    pts += [(1.0 - ti, 1.0) for ti in t[1:]]
    if N > 1:
        pts += [(0.0, 1.0 - ti) for ti in t[1:-1]]
    return np.array(pts, dtype=np.float64)


TET_EDGES = ((0, 1), (0, 2), (0, 3), (1, 2), (1, 3), (2, 3))
_TET_REF = np.array([[0.0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1]])


def tetrahedron_boundary_points(N):
    """OUR 3D extension: lattice points of order N on the surface of the
    reference tetrahedron.  Order: the 4 vertices, then the interior points of the 6 edges
    (edge (a,b), a<b, walked a->b), then face-interior lattice points of faces 0..3.
    N=1 -> the 4 vertices in order; N=2 -> vertices + 6 edge midpoints."""
    if N <= 0:
        return np.array([[0.25, 0.25, 0.25]])
    t = np.linspace(0, 1, N + 1)
    pts = [tuple(v) for v in _TET_REF]
    for a, b in TET_EDGES:
        for ti in t[1:-1]:
            pts.append(tuple((1 - ti) * _TET_REF[a] + ti * _TET_REF[b]))
    for face in LOCAL_FACETS["tetrahedron"]:
        a, b, c = (_TET_REF[i] for i in face)
        for i in range(1, N):
            for j in range(1, N - i):
                pts.append(tuple(a + t[i] * (b - a) + t[j] * (c - a)))
    return np.array(pts, dtype=np.float64)


def cell_detection_points(cell_type, N):
    if cell_type == "triangle":
        return triangle_boundary_points(N)
    if cell_type == "quadrilateral":
        return square_boundary_points(N)
    if cell_type == "tetrahedron":
        return tetrahedron_boundary_points(N)
    raise NotImplementedError(cell_type)  # mesh_scripts.py:326-329


def facet_detection_points(cell_type, N):
    """Points on the reference facet (mesh_scripts.py:434); triangle facets for tetrahedra."""
    return triangle_boundary_points(N) if cell_type == "tetrahedron" else segment_points(N)


_REF_VERTS = {
    "triangle": np.array([[0.0, 0], [1, 0], [0, 1]]),
    "quadrilateral": np.array([[0.0, 0], [1, 0], [0, 1], [1, 1]]),
    "tetrahedron": _TET_REF,
}


def facet_points_in_cell(cell_type, N):
    """[nfpc, nq, tdim]: facet detection points pushed to each local facet of the reference
    cell; reference-facet vertex k sits on the k-th vertex of LOCAL_FACETS[...] (ascending)."""
    fp = facet_detection_points(cell_type, N)
    rv = _REF_VERTS[cell_type]
    out = []
    for lf in LOCAL_FACETS[cell_type]:
        v0 = rv[lf[0]]
        p = np.tile(v0, (len(fp), 1))
        for k in range(1, len(lf)):
            p = p + fp[:, k - 1:k] * (rv[lf[k]] - v0)
        out.append(p)
    return np.array(out)


# --------------------------------------------------------------------------------------
# geometry: P1/Q1 coordinate element
# --------------------------------------------------------------------------------------
def coordinate_basis(cell_type, pts):
    """Values [npts, nvpc] and reference gradients [npts, nvpc, tdim] of the affine /
    bilinear coordinate element."""
    pts = np.asarray(pts, dtype=np.float64)
    if cell_type == "quadrilateral":
        X, Y = pts[:, 0], pts[:, 1]
        val = np.stack([(1 - X) * (1 - Y), X * (1 - Y), (1 - X) * Y, X * Y], axis=1)
        dX = np.stack([-(1 - Y), (1 - Y), -Y, Y], axis=1)
        dY = np.stack([-(1 - X), -X, (1 - X), X], axis=1)
        return val, np.stack([dX, dY], axis=2)
    tdim = TDIM[cell_type]
    val = np.concatenate([(1 - pts.sum(1))[:, None], pts], axis=1)
    grad = np.zeros((len(pts), tdim + 1, tdim))
    grad[:, 0, :] = -1.0
    for d in range(tdim):
        grad[:, d + 1, d] = 1.0
    return val, grad


def _seq_dot(weights, values):
    """sum_k weights[k] * values[..., k] accumulated left to right, skipping exact-zero
    weights (FFCx drops zero table columns)."""
    acc = None
    for k, w in enumerate(weights):
        if w == 0.0:
            continue
        term = values[..., k] if w == 1.0 else w * values[..., k]
        acc = term if acc is None else acc + term
    if acc is None:
        acc = np.zeros(values.shape[:-1])
    return acc


def physical_points(x, cells, cell_type, ref_pts):
    """x_q = sum_v N_v(xi_q) x_v -> [Nc, npts, gdim]."""
    val, _ = coordinate_basis(cell_type, ref_pts)
    xc = x[cells]  # [Nc, nvpc, gdim]
    out = np.empty((len(cells), len(ref_pts), x.shape[1]))
    for q in range(len(ref_pts)):
        for d in range(x.shape[1]):
            out[:, q, d] = _seq_dot(val[q], xc[:, :, d])
    return out


def cell_scale(x, cells, cell_type, ref_pts):
    """|det J| at each detection point -> [Nc, npts] (constant per cell on simplices)."""
    xc = x[cells]
    npts = len(ref_pts)
    if cell_type == "triangle":
        j00 = xc[:, 1, 0] - xc[:, 0, 0]
        j01 = xc[:, 2, 0] - xc[:, 0, 0]
        j10 = xc[:, 1, 1] - xc[:, 0, 1]
        j11 = xc[:, 2, 1] - xc[:, 0, 1]
        det = j00 * j11 - j01 * j10
        return np.repeat(np.abs(det)[:, None], npts, axis=1)
    if cell_type == "tetrahedron":
        a = xc[:, 1] - xc[:, 0]
        b = xc[:, 2] - xc[:, 0]
        c = xc[:, 3] - xc[:, 0]
        # columns of J are a, b, c; cofactor expansion along the first row of J
        det = (a[:, 0] * (b[:, 1] * c[:, 2] - c[:, 1] * b[:, 2])
               - b[:, 0] * (a[:, 1] * c[:, 2] - c[:, 1] * a[:, 2])
               + c[:, 0] * (a[:, 1] * b[:, 2] - b[:, 1] * a[:, 2]))
        return np.repeat(np.abs(det)[:, None], npts, axis=1)
    _, grad = coordinate_basis(cell_type, ref_pts)
    out = np.empty((len(cells), npts))
    for q in range(npts):
        j00 = _seq_dot(grad[q, :, 0], xc[:, :, 0])
        j01 = _seq_dot(grad[q, :, 1], xc[:, :, 0])
        j10 = _seq_dot(grad[q, :, 0], xc[:, :, 1])
        j11 = _seq_dot(grad[q, :, 1], xc[:, :, 1])
        out[:, q] = np.abs(j00 * j11 - j01 * j10)
    return out


def facet_scale(x, cells, cell_type):
    """Facet integral scale of every (cell, local facet): edge length in 2D, |e1 x e2| (twice the
    area) in 3D -> [Nc, nfpc]."""
    xc = x[cells]
    out = np.empty((len(cells), len(LOCAL_FACETS[cell_type])))
    for i, lf in enumerate(LOCAL_FACETS[cell_type]):
        if len(lf) == 2:
            d = xc[:, lf[1]] - xc[:, lf[0]]
            out[:, i] = np.sqrt(d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1])
        else:
            e1 = xc[:, lf[1]] - xc[:, lf[0]]
            e2 = xc[:, lf[2]] - xc[:, lf[0]]
            cx = e1[:, 1] * e2[:, 2] - e1[:, 2] * e2[:, 1]
            cy = e1[:, 2] * e2[:, 0] - e1[:, 0] * e2[:, 2]
            cz = e1[:, 0] * e2[:, 1] - e1[:, 1] * e2[:, 0]
            out[:, i] = np.sqrt(cx * cx + cy * cy + cz * cz)
    return out


# --------------------------------------------------------------------------------------
# level-set point values
# --------------------------------------------------------------------------------------
def point_values_function(coeffs, dofmap, table):
    """phi_q = sum_i c_i psi_i(xi_q), sequential over i.  table [..., npts, nd] -> [Nc, ..., npts]."""
    table = np.asarray(table)
    cd = coeffs[dofmap]  # [Nc, nd]
    lead = table.shape[:-1]
    out = np.empty((len(dofmap),) + lead)
    for idx in np.ndindex(*lead):
        out[(slice(None),) + idx] = _seq_dot(table[idx], cd)
    return out


def point_values_expression(func, x, cells, cell_type, ref_pts):
    """phi at physical detection points, func called like a UFL expression of x: func(x)[...]
    with x of shape (gdim, npoints).  ref_pts [..., npts, tdim] -> [Nc, ..., npts]."""
    ref_pts = np.asarray(ref_pts)
    lead = ref_pts.shape[:-1]
    flat = ref_pts.reshape(-1, ref_pts.shape[-1])
    xq = physical_points(x, cells, cell_type, flat)  # [Nc, P, gdim]
    with np.errstate(all="ignore"):
        vals = func(xq.reshape(-1, x.shape[1]).T)
    return np.asarray(vals, dtype=np.float64).reshape((len(cells),) + lead)


# --------------------------------------------------------------------------------------
# topology
# --------------------------------------------------------------------------------------
def build_topology(cells, cell_type):
    """c2f [Nc, nfpc], f2c [Nf, 2] (ascending cells, -1 pad), facet_vertices [Nf, nvpf] (sorted
    tuples, rows in lexicographic order = facet index)."""
    cells = np.asarray(cells, dtype=np.int64)
    lfs = LOCAL_FACETS[cell_type]
    nc, nf_loc = len(cells), len(lfs)
    keys = np.sort(np.stack([cells[:, list(lf)] for lf in lfs], axis=1), axis=2)  # [Nc, nfpc, nvpf]
    uniq, inv = np.unique(keys.reshape(nc * nf_loc, -1), axis=0, return_inverse=True)
    c2f = inv.reshape(nc, nf_loc).astype(np.int32)
    f2c = -np.ones((len(uniq), 2), dtype=np.int32)
    owner = np.repeat(np.arange(nc, dtype=np.int32), nf_loc)
    order = np.argsort(c2f.ravel(), kind="stable")  # stable => ascending cell per facet
    fs, cs = c2f.ravel()[order], owner[order]
    first = np.r_[True, fs[1:] != fs[:-1]]
    f2c[fs[first], 0] = cs[first]
    f2c[fs[~first], 1] = cs[~first]
    return c2f, f2c, uniq.astype(np.int32)


# --------------------------------------------------------------------------------------
# detection + tagging  (mesh_scripts.py:95-134, 284-390, 393-558)
# --------------------------------------------------------------------------------------
def detection_ratio(num, den, warn=True):
    """mesh_scripts.py:124-133."""
    with np.errstate(all="ignore"):
        ok = den > 0.0
        d = np.full_like(num, 0.5)
        d[ok] = num[ok] / den[ok]
    if warn and np.any(np.isclose(den, 0.0)):
        warnings.warn("The detection function is zero everywhere on a cell. We mark it as 'cut' "
                      "but this can be incorrect and should be carefully checked.", RuntimeWarning)
    return d


def detection_sums_cells(phi_cell, scale):
    """num_c = sum_q phi_q*s_q, den_c = sum_q |phi_q|*s_q, sequential in q (mesh_scripts.py:113-119)."""
    num = np.zeros(len(phi_cell))
    den = np.zeros(len(phi_cell))
    with np.errstate(all="ignore"):
        for q in range(phi_cell.shape[1]):
            t = phi_cell[:, q] * scale[:, q]
            num = num + t
            den = den + np.abs(t)
    return num, den


def tag_cells(phi_cell, scale, cells=None, single_layer_cut=False, warn=True):
    """mesh_scripts.py:284-390 -> int32 [Nc] with 1 interior, 2 cut, 3 exterior, 0 = untagged
    (only possible when the ratio is NaN with a positive denominator)."""
    num, den = detection_sums_cells(phi_cell, scale)
    d = detection_ratio(num, den, warn)
    tags = np.zeros(len(d), dtype=np.int32)
    tags[(d > -1.0) & (d < 1.0)] = 2
    tags[d == 1.0] = 3
    tags[d == -1.0] = 1
    if single_layer_cut:  # mesh_scripts.py:349-358
        nv = int(cells.max()) + 1
        touches_interior = np.zeros(nv, dtype=bool)
        touches_interior[np.unique(cells[tags == 1])] = True
        isolated = (tags == 2) & ~touches_interior[cells].any(axis=1)
        tags[isolated] = 3
    return tags


def detection_sums_facets(phi_facet, fscale, c2f, f2c):
    """The `ds` detection (mesh_scripts.py:443-447): per exterior facet sum_q phi_q * s_f (sequential),
    added into the owner cell's entry in ascending facet index."""
    nc = len(c2f)
    num = np.zeros(nc)
    den = np.zeros(nc)
    bnd = f2c[:, 1] < 0
    with np.errstate(all="ignore"):
        for f in np.nonzero(bnd)[0]:  # ascending facet index
            c = f2c[f, 0]
            lf = int(np.nonzero(c2f[c] == f)[0][0])
            fn = 0.0
            fd = 0.0
            for q in range(phi_facet.shape[2]):
                t = phi_facet[c, lf, q] * fscale[c, lf]
                fn = fn + t
                fd = fd + abs(t)
            num[c] = num[c] + fn
            den[c] = den[c] + fd
    return num, den


def _mark(nf, facets):
    m = np.zeros(nf, dtype=bool)
    m[np.asarray(facets).ravel()] = True
    return m


def tag_facets(cell_tags, c2f, f2c, phi_facet, fscale, warn=True, strict=False):
    """mesh_scripts.py:393-558 as per-facet membership masks.  Returns int32 [Nf] (0 untagged) and
    the list of facets the reference would emit twice (inconsistent combinations);
    for those, the tag written here follows the precedence 6 > 4 > 2 > 3 > 1 > 5 (last writer
    below).  strict=True raises on them."""
    nf = len(f2c)
    I, C, E = (cell_tags == 1), (cell_tags == 2), (cell_tags == 3)
    inI, inC, inE = _mark(nf, c2f[I]), _mark(nf, c2f[C]), _mark(nf, c2f[E])
    bnd = f2c[:, 1] < 0                                            # :430-432
    num, den = detection_sums_facets(phi_facet, fscale, c2f, f2c)  # :445-447
    d = detection_ratio(num, den, warn)
    kcell = (d > -1.0) & (d < 1.0)                                 # :448-452
    cut_bnd = _mark(nf, c2f[kcell]) & bnd                          # :454-456
    uncut_bnd = _mark(nf, c2f[~kcell]) & bnd & ~inE & ~inI         # :457-461
    int_bnd = inI & inC                                            # :464-466  -> 3
    if not E.any():                                                # :469-470
        boundary = bnd.copy()
    else:                                                          # :471-474
        boundary = (inE & inC) | uncut_bnd
    direct = inE & inI                                             # :476-478  -> 6
    cutf = (inC & ~(boundary | int_bnd | direct | uncut_bnd)) | cut_bnd   # :480-485 -> 2
    rem = int_bnd | boundary | direct
    interior = inI & ~rem                                          # :488-490  -> 1
    exterior = inE & ~rem                                          # :493-495  -> 5
    boundary = boundary & ~cutf                                    # :497      -> 4
    sets = [(5, exterior), (1, interior), (3, int_bnd), (2, cutf), (4, boundary), (6, direct)]
    mult = sum(m.astype(np.int32) for _, m in sets)
    dup = np.nonzero(mult > 1)[0]
    if strict and len(dup):
        raise ValueError("facets tagged twice by the reference algebra: %s" % dup[:10])
    tags = np.zeros(nf, dtype=np.int32)
    for v, m in sets:
        tags[m] = v
    return tags, dup, kcell


# --------------------------------------------------------------------------------------
# one-sided measures, submesh  (mesh_scripts.py:137-192, 217-281, 617-645)
# --------------------------------------------------------------------------------------
def integration_entities(c2f, f2c, cell_mask, facet_mask):
    """mesh_scripts.py:137-192: flat [cell, local_facet, ...] int32.  Cells appear in the order of
    first appearance when walking the selected facets in ascending index and, per facet, its cells
    in REVERSE link order (that is what `_reshape_map` :195-214 produces); per cell, local facets
    ascending."""
    facets = np.nonzero(facet_mask)[0]
    conn = f2c[facets][:, ::-1]            # reverse link order; boundary facets: [-1, c]
    one = conn[:, 0] < 0
    conn = conn.copy()
    conn[one, 0], conn[one, 1] = conn[one, 1], -1   # _reshape_map puts the single link first
    flat = conn.ravel()
    flat = flat[(flat >= 0)]
    flat = flat[cell_mask[flat]]
    _, first = np.unique(flat, return_index=True)
    ordered = flat[np.sort(first)]
    out = []
    for c in ordered:
        for lf in range(c2f.shape[1]):
            if facet_mask[c2f[c, lf]]:
                out.extend((c, lf))
    return np.asarray(out, dtype=np.int32)


def submesh(x, cells, keep):
    """Order-preserving submesh of the cells `keep` (sorted): vertices renumbered in ascending
    parent index (what dolfinx create_submesh does in serial [probed, SURVEY.md A.5])."""
    sub_parent = cells[keep]
    v_map = np.unique(sub_parent)
    renum = -np.ones(int(cells.max()) + 1, dtype=np.int64)
    renum[v_map] = np.arange(len(v_map))
    return x[v_map], renum[sub_parent].astype(np.int32), v_map.astype(np.int32)


def transfer_facet_tags(facet_tags, parent_c2f, sub_c2f, c_map):
    """mesh_scripts.py:244-260: first occurrence of each submesh facet in the flattened submesh
    c->f map picks the parent facet at the same position of parent_c2f[c_map]."""
    src = parent_c2f[c_map].ravel()
    _, first = np.unique(sub_c2f.ravel(), return_index=True)
    return facet_tags[src[first]].astype(np.int32)


def compute_tags_measures(x, cells, cell_type, phi_cell, phi_facet, box_mode=False,
                          single_layer_cut=False, detection_points=None, warn=False):
    """mesh_scripts.py:571-653 on arrays.  Returns a dict with cell_tags, facet_tags (parent or
    submesh numbering), ds100 / ds101 entity lists (box mode) or the submesh arrays."""
    c2f, f2c, fverts = build_topology(cells, cell_type)
    pts = detection_points
    scale = cell_scale(x, cells, cell_type, pts)
    fscale = facet_scale(x, cells, cell_type)
    ct = tag_cells(phi_cell, scale, cells, single_layer_cut, warn)
    ft, dup, kcell = tag_facets(ct, c2f, f2c, phi_facet, fscale, warn)
    out = {"c2f": c2f, "f2c": f2c, "facet_vertices": fverts, "duplicates": dup,
           "parent_cell_tags": ct, "parent_facet_tags": ft}
    if box_mode:  # :617-634
        out["cell_tags"], out["facet_tags"] = ct, ft
        out["ds100"] = integration_entities(c2f, f2c, (ct == 1) | (ct == 2), ft == 4)
        out["ds101"] = integration_entities(c2f, f2c, (ct == 2) | (ct == 3), ft == 3)
    else:         # :635-645
        keep = np.nonzero((ct == 1) | (ct == 2))[0]
        sx, scells, v_map = submesh(x, cells, keep)
        sc2f, sf2c, _ = build_topology(scells, cell_type)
        out.update(cell_tags=ct[keep], facet_tags=transfer_facet_tags(ft, c2f, sc2f, keep),
                   sub_x=sx, sub_cells=scells, c_map=keep.astype(np.int32), v_map=v_map)
    return out


def outward_normals(x, cells, cell_type, entities):
    """Outward unit normal and measure of each (cell, local_facet) entity -> ([m, gdim], [m]).
    Used to restate reference tests/test_one_sided_integral.py:137-141."""
    ents = np.asarray(entities).reshape(-1, 2)
    xc = x[cells[ents[:, 0]]]                      # [m, nvpc, gdim]
    centroid = xc.mean(axis=1)
    lfs = LOCAL_FACETS[cell_type]
    n = np.zeros((len(ents), x.shape[1]))
    meas = np.zeros(len(ents))
    for i, (c, lf) in enumerate(ents):
        p = xc[i][list(lfs[lf])]
        if len(p) == 2:
            t = p[1] - p[0]
            nn = np.array([t[1], -t[0]])
            meas[i] = np.linalg.norm(t)
        else:
            nn = np.cross(p[1] - p[0], p[2] - p[0])
            meas[i] = 0.5 * np.linalg.norm(nn)
        nn = nn / np.linalg.norm(nn)
        if np.dot(nn, p.mean(axis=0) - centroid[i]) < 0:
            nn = -nn
        n[i] = nn
    return n, meas
