"""ctypes wrapper of oracle/csrc/phifem_oracle.c (TEST INFRASTRUCTURE / CPU baseline only)."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "_build", "libphifem_oracle.so")
_lib = None

_d = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_i = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_i64, _int, _dbl = ctypes.c_int64, ctypes.c_int, ctypes.c_double


def load():
    global _lib
    if _lib is None:
        src = os.path.join(_HERE, "csrc", "phifem_oracle.c")
        if not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(src):
            subprocess.run(["make", "-s", "-C", _HERE], check=True)
        lib = ctypes.CDLL(_LIB)
        lib.oracle_num_threads.restype = _int
        lib.oracle_set_num_threads.argtypes = [_int]
        lib.oracle_set_num_threads.restype = None
        lib.oracle_tag_cells_p1.argtypes = [_d, _int, _i, _i64, _d, _i]
        lib.oracle_tag_facets_p1.argtypes = [_d, _int, _i, _i, _i, _i64, _i64, _d, _i, _i]
        lib.oracle_assemble_cells_p1.argtypes = [_d, _int, _i, _d, _d, _i, _i, _i64, _i, _dbl, _d, _d]
        lib.oracle_assemble_boundary_p1.argtypes = [_d, _int, _i, _d, _i, _i64, _i, _d]
        lib.oracle_assemble_ghost_p1.argtypes = [_d, _int, _i, _i, _i, _d, _i, _i64, _i, _dbl, _d]
        for name in ("oracle_tag_cells_p1", "oracle_tag_facets_p1", "oracle_assemble_cells_p1",
                     "oracle_assemble_boundary_p1", "oracle_assemble_ghost_p1"):
            getattr(lib, name).restype = None
        _lib = lib
    return _lib


def num_threads():
    return load().oracle_num_threads()


def set_num_threads(n):
    load().oracle_set_num_threads(int(n))


def tag_cells_p1(x, cells, phi):
    tags = np.empty(len(cells), dtype=np.int32)
    load().oracle_tag_cells_p1(x, x.shape[1], cells, len(cells), phi, tags)
    return tags


def tag_facets_p1(x, cells, c2f, f2c, phi, ctags):
    ftags = np.empty(len(f2c), dtype=np.int32)
    load().oracle_tag_facets_p1(x, x.shape[1], cells, c2f, f2c, len(cells), len(f2c), phi, ctags, ftags)
    return ftags


def assemble_p1(x, cells, c2f, f2c, phi, f, ctags, active, slots_cells, entities, slots_boundary,
                ghost, slots_ghost, sigma, nnz):
    lib = load()
    data = np.zeros(nnz)
    b = np.zeros(len(x))
    D = x.shape[1]
    lib.oracle_assemble_cells_p1(x, D, cells, phi, f, ctags, active, len(active), slots_cells, sigma, data, b)
    if len(entities):
        lib.oracle_assemble_boundary_p1(x, D, cells, phi, entities, len(entities), slots_boundary, data)
    if len(ghost):
        lib.oracle_assemble_ghost_p1(x, D, cells, c2f, f2c, phi, ghost, len(ghost), slots_ghost, sigma, data)
    return data, b
