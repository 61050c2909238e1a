"""Shared test data: the reference's level sets and mesh fixtures.

Level sets restate reference tests/test_compute_meshtags.py:21-104 and
tests/test_one_sided_integral.py:15-90 (numpy mode; x has shape (gdim|3, npoints))."""
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_MESHES = np.load(os.path.join(HERE, "golden", "meshes.npz"))
_GOLD = None

MESH_CELL_TYPE = {"coarse_square": "triangle", "disk": "triangle", "square_tri": "triangle",
                  "square_quad": "quadrilateral"}


def load_mesh_arrays(name):
    """(x, cells in dolfinx vertex order, cell_type)."""
    x = _MESHES[name + "_x"]
    cells = _MESHES[name + "_cells"]
    ct = MESH_CELL_TYPE[name]
    if ct == "quadrilateral":
        cells = cells[:, [0, 1, 3, 2]]
    return x, np.ascontiguousarray(cells), ct


def golden(stem):
    global _GOLD
    if _GOLD is None:
        _GOLD = np.load(os.path.join(HERE, "golden", "golden_tags.npz"))
    return _GOLD[stem] if stem in _GOLD.files else None


def quadratic(x0, a, x1, b, c):
    return lambda x: (a * x[0] - x0) ** 2 + (b * x[1] - x1) ** 2 + c


def square_levelset(x):
    return np.maximum(abs(x[0]), abs(x[1])) - 1.0


def nasty_levelset(x):
    th = np.arctan2(x[1], x[0])
    return np.sqrt(x[0] ** 2 + x[1] ** 2) * (abs(th) * np.sin(1.0 / abs(th))) - 0.25


TAG_DATA = [
    ("circle_in_circle", "disk", quadratic(0.0, 1.0, 0.0, 1.0, -0.125)),
    ("boundary_crossing_circle", "disk", quadratic(0.0, 1.0, -0.5, 1.0, -0.125)),
    ("circle_in_square", "square_quad", quadratic(0.0, 1.0, 0.0, 1.0, -0.125)),
    ("square_in_square", "square_tri", square_levelset),
    ("ellipse_in_square", "square_quad", quadratic(0.0, 1.0, 0.1, 0.3, -0.65)),
    ("circle_near_boundary", "coarse_square", quadratic(0.5, 1.0, 0.5, 1.0, -0.2)),
    ("nasty_levelset", "square_tri", nasty_levelset),
]


def golden_names(data_name, degree, discretize, box_mode, single_layer):
    """File-name scheme of reference tests/test_compute_meshtags.py:139-151."""
    mid = "_"
    if discretize:
        mid += "discretize_"
    if not box_mode:
        mid += "submesh_"
    if single_layer:
        mid += "single_layer_"
    stem = "%s_%d%s" % (data_name, degree, mid)
    return stem + "cells_tags", stem + "facets_tags"


def case_class(data_name, mesh_name, degree, discretize):
    """'exact' | 'hist' | 'degenerate' -- SURVEY.md Appendix D.4."""
    if mesh_name == "disk":
        return "hist"
    if data_name == "square_in_square" and (discretize or degree == 3):
        return "degenerate"
    if data_name == "nasty_levelset" and discretize:
        return "degenerate"
    if data_name == "ellipse_in_square" and degree == 3 and not discretize:
        return "degenerate"
    return "exact"


# one-sided integrals: reference tests/test_one_sided_integral.py
ONE_SIDED = [
    ("line_in_square_quad", "square_quad", lambda x: x[0] + 0.35, (3.0, -3.0), "signed"),
    ("square_in_square_quad", "square_quad",
     lambda x: np.maximum(abs(x[0]), abs(x[1])) - 0.35, (3.2, 2.4), "abs"),
    ("square_in_square_tri", "square_tri",
     lambda x: np.maximum(abs(x[0]), abs(x[1])) - 0.325, (3.2, 2.4), "abs"),
]


# flower domain of the reference demos (demo/weak-dirichlet/flower/data.py:27-99, demo/strong-dirichlet/flower/
# data.py:16-62): a disc of radius 2 with eight petals; `flower_detection` is the non-smooth min used for the
# tags (data.py:57-82), `flower_levelset` the graded smooth-min used in the forms (:27-54), `flower_source`
# 10 on a small disc inside the first petal (:85-99).
def _flower_parts(x):
    s = np.cos(np.pi / 8.0) + np.sin(np.pi / 8.0)
    rp = np.sqrt(2.0) * 2.0 * s * np.sin(np.pi / 8.0)
    yield x[0] ** 2 + x[1] ** 2 - 4.0
    for i in range(1, 9):
        cx, cy = 2.0 * s * np.cos(i * np.pi / 4.0), 2.0 * s * np.sin(i * np.pi / 4.0)
        yield (x[0] - cx) ** 2 + (x[1] - cy) ** 2 - rp ** 2


def flower_detection(x):
    val = None
    for part in _flower_parts(x):
        val = part if val is None else np.minimum(val, part)
    return val


def flower_levelset(x):
    r = np.sqrt(x[0] ** 2 + x[1] ** 2)
    k = (np.pi / 2.0 - np.arctan(50.0 * (r - 2.0))) / np.pi / 2.0          # graded smoothing width in (0, 1/2)
    val = None
    for part in _flower_parts(x):
        if val is None:
            val = part
            continue
        lo = np.minimum(val, part)
        val = np.maximum(k, lo) - np.sqrt(np.maximum(k - val, 0.0) ** 2 + np.maximum(k - part, 0.0) ** 2)
    return val


def flower_source(x):
    s = np.cos(np.pi / 8.0) + np.sin(np.pi / 8.0)
    r1 = np.sqrt(2.0) * 2.0 * s * np.sin(np.pi / 8.0)
    return np.where((x[0] - 2.0 * s) ** 2 + x[1] ** 2 <= r1 ** 2 / 2.0, 10.0, 0.0)
