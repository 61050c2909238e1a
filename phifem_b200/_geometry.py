"""Reference-cell data of the hot path: detection points, local facets, coordinate element.

Host-side constants only (a few dozen doubles); they parameterise the CUDA kernels.
Follows reference src/phifem/mesh_scripts.py:28-92 (point generators) and the dolfinx 0.9
reference-cell conventions [dep-knowledge, SURVEY.md Appendix C].
"""
import numpy as np

CELL_TYPES = ("triangle", "quadrilateral", "tetrahedron")
CELL_TYPE_ID = {name: i for i, name in enumerate(CELL_TYPES)}
TDIM = {"triangle": 2, "quadrilateral": 2, "tetrahedron": 3}
NVPC = {"triangle": 3, "quadrilateral": 4, "tetrahedron": 4}
# simplex local facet i is opposite local vertex i; quadrilateral facets in tensor order
LOCAL_FACETS = {
    "triangle": ((1, 2), (0, 2), (0, 1)),
    "quadrilateral": ((0, 1), (0, 2), (1, 3), (2, 3)),
    "tetrahedron": ((1, 2, 3), (0, 2, 3), (0, 1, 3), (0, 1, 2)),
}
REF_VERTICES = {
    "triangle": np.array([[0.0, 0.0], [1.0, 0.0], [0.0, 1.0]]),
    "quadrilateral": np.array([[0.0, 0.0], [1.0, 0.0], [0.0, 1.0], [1.0, 1.0]]),
    "tetrahedron": np.array([[0.0, 0, 0], [1.0, 0, 0], [0, 1.0, 0], [0, 0, 1.0]]),
}
TET_EDGES = ((0, 1), (0, 2), (0, 3), (1, 2), (1, 3), (2, 3))


def reference_segment_points(N):
    """Reference mesh_scripts.py:28-40."""
    if N > 0:
        return np.linspace(0, 1, N + 1).astype(np.float64)[:, None]
    return np.array([[0.5]])


def reference_triangle_boundary_points(N):
    """Reference mesh_scripts.py:43-65 (order: v0 -> v1 -> v2 -> back towards v0)."""
    if N <= 0:
        return np.array([[1.0 / 3.0, 1.0 / 3.0]])
    t = np.linspace(0, 1, N + 1)
    rows = [np.stack([t, np.zeros_like(t)], axis=1),
            np.stack([1 - t[1:], t[1:]], axis=1)]
    if N > 1:
        rows.append(np.stack([np.zeros_like(t[1:-1]), 1 - t[1:-1]], axis=1))
    return np.concatenate(rows, axis=0).astype(np.float64)


def reference_square_boundary_points(N):
    """Reference mesh_scripts.py:68-92."""
    if N <= 0:
        return np.array([[0.5, 0.5]])
    t = np.linspace(0, 1, N + 1)
    rows = [np.stack([t, np.zeros_like(t)], axis=1),
            np.stack([np.ones_like(t[1:]), t[1:]], axis=1),
            np.stack([1.0 - t[1:], np.ones_like(t[1:])], axis=1)]
    if N > 1:
        rows.append(np.stack([np.zeros_like(t[1:-1]), 1.0 - t[1:-1]], axis=1))
    return np.concatenate(rows, axis=0).astype(np.float64)


def reference_tetrahedron_boundary_points(N):
    """3D extension (the reference stops at 2D, mesh_scripts.py:326-329): order-N lattice points on
    the surface of the reference tetrahedron -- 4 vertices, edge-interior points (edges
    (0,1),(0,2),(0,3),(1,2),(1,3),(2,3) walked low->high), face-interior points of faces 0..3."""
    rv = REF_VERTICES["tetrahedron"]
    if N <= 0:
        return np.array([[0.25, 0.25, 0.25]])
    t = np.linspace(0, 1, N + 1)
    pts = [rv[i] for i in range(4)]
    for a, b in TET_EDGES:
        for s in t[1:-1]:
            pts.append((1 - s) * rv[a] + s * rv[b])
    for face in LOCAL_FACETS["tetrahedron"]:
        a, b, c = (rv[i] for i in face)
        for i in range(1, N):
            for j in range(1, N - i):
                pts.append(a + t[i] * (b - a) + t[j] * (c - a))
    return np.array(pts, dtype=np.float64)


def cell_detection_points(cell_type, N):
    if cell_type == "triangle":
        return reference_triangle_boundary_points(N)
    if cell_type == "quadrilateral":
        return reference_square_boundary_points(N)
    if cell_type == "tetrahedron":
        return reference_tetrahedron_boundary_points(N)
    raise NotImplementedError(
        "Mesh tags computation does not support other cell types than 'triangle', "
        "'quadrilateral' or 'tetrahedron'")


def facet_detection_points(cell_type, N):
    """Points on the reference facet, mapped onto every local facet -> [nfpc, nq, tdim]."""
    if cell_type == "tetrahedron":
        fp = reference_triangle_boundary_points(N)
    else:
        fp = reference_segment_points(N)
    rv = REF_VERTICES[cell_type]
    out = []
    for lf in LOCAL_FACETS[cell_type]:
        p = np.tile(rv[lf[0]], (len(fp), 1))
        for k in range(1, len(lf)):
            p = p + fp[:, k - 1:k] * (rv[lf[k]] - rv[lf[0]])
        out.append(p)
    return np.array(out)


def coordinate_basis(cell_type, pts):
    """P1 / Q1 coordinate-element values at reference points -> [npts, nvpc]."""
    pts = np.asarray(pts, dtype=np.float64)
    if cell_type == "quadrilateral":
        X, Y = pts[:, 0], pts[:, 1]
        return np.stack([(1 - X) * (1 - Y), X * (1 - Y), (1 - X) * Y, X * Y], axis=1)
    return np.concatenate([(1 - pts.sum(axis=1))[:, None], pts], axis=1)


def coordinate_basis_grad(cell_type, pts):
    """Reference gradients [npts, nvpc, tdim] of the coordinate element."""
    pts = np.asarray(pts, dtype=np.float64)
    if cell_type == "quadrilateral":
        X, Y = pts[:, 0], pts[:, 1]
        dX = np.stack([-(1 - Y), (1 - Y), -Y, Y], axis=1)
        dY = np.stack([-(1 - X), -X, (1 - X), X], axis=1)
        return np.stack([dX, dY], axis=2)
    tdim = TDIM[cell_type]
    g = np.zeros((len(pts), tdim + 1, tdim))
    g[:, 0, :] = -1.0
    for d in range(tdim):
        g[:, d + 1, d] = 1.0
    return g
