"""CPU-only: the C-ABI library loads and exports every symbol include/phifem_b200.h declares."""
import ctypes
import os
import re

from phifem_b200 import _lib, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "phifem_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(phifem_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    build.build()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    declared = _declared_symbols()
    assert len(declared) >= 9
    for name in declared:
        assert hasattr(lib, name), name
    assert sorted(_lib.EXPORTED_SYMBOLS) == declared
    assert _lib.load().phifem_abi_version() == 1


def test_argument_errors_do_not_need_a_gpu():
    lib = _lib.load()
    rc = lib.phifem_tag_cells(None, None, 0, None, None, None, None, None)
    assert rc == -1 and b"mesh" in lib.phifem_last_error()
    m = _lib.CMesh(7, 2, 0, 0, 0, 1, 1, 1, 1, 0.0, 0.0, None, 0)
    ls = _lib.CLevelset()
    rc = lib.phifem_tag_cells(ctypes.byref(m), ctypes.byref(ls), 0, 1, 1, None, 1, None)
    assert rc == -3 and b"unsupported cell type" in lib.phifem_last_error()


def test_post_to_host_rejects_bad_arguments_without_a_gpu():
    lib = _lib.load()
    assert lib.phifem_post_to_host(None, None, 16, None) == -1 and b"null pointer" in lib.phifem_last_error()
    buf = (ctypes.c_int64 * 4)()
    assert lib.phifem_post_to_host(ctypes.addressof(buf), ctypes.addressof(buf), 0, None) == -1
    assert b"between 1 and 4096" in lib.phifem_last_error()


def test_hot_path_refuses_cpu_meshes():
    import numpy as np
    import pytest
    from phifem_b200 import mesh_scripts
    from phifem_b200.mesh import Mesh
    m = Mesh(np.array([[0.0, 0], [1, 0], [0, 1]]), np.array([[0, 1, 2]]), "triangle", device="cpu")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        mesh_scripts.compute_tags_measures(m, lambda x: x[0] - 0.5, 1, box_mode=True)
