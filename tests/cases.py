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


# flower domain of the reference demos (restated in phifem_b200/synthetic.py; checked against the reference's own
# data.py values in tests/golden/flower_demo.npz)
from phifem_b200.synthetic import flower_detection, flower_levelset, flower_source  # noqa: E402,F401
