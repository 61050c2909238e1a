"""The symbolic phase of the P1 strong-Dirichlet operator through the C ABI (csrc/symbolic.cu): CSR pattern and
entity -> slot maps built on the device with cub sort / unique, for hosts that do not go through the torch-based plan of
phifem_b200/assemble.py.  `DevicePattern` wraps the library-owned arrays as torch tensors (zero copy) and assembles with
the per-entity kernels.

    pat = symbolic.DevicePattern(mesh, cells_tags, facets_tags, ds_bdy(100))
    A, b = pat.assemble(phi, f, stab_coef=1.0)
"""
import ctypes

import torch

from . import _lib
from .assemble import CSRMatrix, _device_vector, _plan_inputs


def _as_tensor(ptr, n, device):
    """int32 device array owned by the library -> torch view (kept alive by the DevicePattern)."""
    if n == 0:
        return torch.zeros(0, dtype=torch.int32, device=device)

    class _Holder:
        __cuda_array_interface__ = {"shape": (int(n),), "typestr": "<i4", "data": (int(ptr), False), "version": 3}
    return torch.as_tensor(_Holder(), device=device)


class DevicePattern:
    def __init__(self, mesh, cells_tags, facets_tags, ds=None):
        _lib.require_cuda(mesh)
        c8, f8, ents = _plan_inputs(mesh, cells_tags, facets_tags, ds)
        self.mesh, self.cell_tags8 = mesh, c8
        self.entities = ents.reshape(-1, 2).to(torch.int32).contiguous()
        lib = _lib.load()
        handle = ctypes.c_void_p()
        _lib.check(lib.phifem_pattern_create_p1(_lib.c_mesh(mesh), _lib.ptr(c8), _lib.ptr(f8),
                                                _lib.ptr(self.entities) if self.entities.numel() else None,
                                                self.entities.shape[0], ctypes.byref(handle), _lib.stream()))
        self._handle = handle
        v = _lib.CPatternView()
        _lib.check(lib.phifem_pattern_view_of(handle, ctypes.byref(v)))
        nv = mesh.cells.shape[1]
        dev = mesh.device
        self.n_rows, self.nnz = int(v.n_rows), int(v.nnz)
        self.indptr = _as_tensor(v.indptr, v.n_rows + 1, dev)
        self.indices = _as_tensor(v.indices, v.nnz, dev)
        self.active = _as_tensor(v.active, v.n_active, dev)
        self.ghost = _as_tensor(v.ghost, v.n_ghost, dev)
        self.slots_cells = _as_tensor(v.slots_cells, v.n_active * nv * nv, dev).reshape(-1, nv * nv)
        self.slots_ghost = _as_tensor(v.slots_ghost, v.n_ghost * (nv + 1) ** 2, dev).reshape(-1, (nv + 1) ** 2)
        self.slots_boundary = _as_tensor(v.slots_boundary, v.n_entities * nv * nv, dev).reshape(-1, nv * nv)

    def __del__(self):
        h = getattr(self, "_handle", None)
        if h and _lib is not None:       # (module globals are gone at interpreter shutdown)
            try:
                _lib.load().phifem_pattern_destroy(h)
            except Exception:
                pass
            self._handle = None

    def assemble(self, phi_h, f_h, stab_coef=1.0):
        """A, b of reference demo/strong-dirichlet/flower/main.py:121-131 through the per-entity kernels."""
        mesh = self.mesh
        lib = _lib.load()
        cm = _lib.c_mesh(mesh)
        phi, f = _device_vector(mesh, phi_h), _device_vector(mesh, f_h)
        data = torch.zeros(self.nnz, dtype=torch.float64, device=mesh.device)
        b = torch.zeros(self.n_rows, dtype=torch.float64, device=mesh.device)
        st = _lib.stream()
        p = _lib.ptr
        if self.active.numel():
            _lib.check(lib.phifem_assemble_cells_p1(cm, p(phi), p(f), p(self.cell_tags8), p(self.active),
                                                    self.active.numel(), p(self.slots_cells), float(stab_coef),
                                                    p(data), p(b), st))
        if self.entities.shape[0]:
            _lib.check(lib.phifem_assemble_boundary_p1(cm, p(phi), p(self.entities), self.entities.shape[0],
                                                       p(self.slots_boundary), p(data), st))
        if self.ghost.numel():
            _lib.check(lib.phifem_assemble_ghost_p1(cm, p(phi), p(self.ghost), self.ghost.numel(),
                                                    p(self.slots_ghost), float(stab_coef), p(data), st))
        A = CSRMatrix(self.indptr, self.indices, data, (self.n_rows, self.n_rows))
        A._pattern = self          # indptr / indices are views of arrays this object owns
        return A, b
