"""The exchange step of a sharded classification over NVLink peer memory (csrc/peer.cu): every rank stores its count
of exterior cells into a slot of every peer's HBM (CUDA IPC mapping) and reads the sum from its own slots -- the global
"any exterior cell" flag of reference src/phifem/mesh_scripts.py:469-474 without a collective kernel.  One process per
GPU on one node; `torch.distributed` only carries the 64-byte IPC handles once, at set-up."""
import ctypes

import torch

from . import _lib


class PeerFlags:
    def __init__(self, rank, world, group=None):
        import torch.distributed as dist
        lib = _lib.load()
        self._p = ctypes.c_void_p()
        handle = ctypes.create_string_buffer(64)
        _lib.check(lib.phifem_peer_flags_create(world, rank, ctypes.byref(self._p), handle))
        handles = [None] * world
        dist.all_gather_object(handles, handle.raw, group=group)
        blob = ctypes.create_string_buffer(b"".join(handles), 64 * world)
        _lib.check(lib.phifem_peer_flags_connect(self._p, blob))
        dist.barrier(group=group)          # everybody has mapped everybody before the first store
        self.rank, self.world = rank, world

    def publish(self, value):
        """value: int64 device tensor (one element, e.g. a slice of the counters): this rank's count."""
        _lib.check(_lib.load().phifem_peer_flags_publish(self._p, _lib.ptr(value), _lib.stream()))

    def collect(self, out):
        """out (int64 device tensor, one element) = sum over the ranks of what they published this epoch."""
        _lib.check(_lib.load().phifem_peer_flags_collect(self._p, _lib.ptr(out), _lib.stream()))

    def timed_out(self):
        return bool(_lib.load().phifem_peer_flags_error(self._p))

    def close(self):
        if self._p:
            _lib.load().phifem_peer_flags_destroy(self._p)
            self._p = ctypes.c_void_p()


def try_create(rank, world, group=None):
    """PeerFlags, or None where peer mapping is not available (CPU / gloo runs, GPUs without peer access)."""
    if world <= 1 or not torch.cuda.is_available():
        return None
    import torch.distributed as dist
    try:
        flags = PeerFlags(rank, world, group)
        ok = 1
    except Exception:   # noqa: BLE001 -- every rank must agree, see below
        flags, ok = None, 0
    t = torch.tensor([ok], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MIN, group=group)
    if int(t.item()) == 0:
        if flags is not None:
            flags.close()
        return None
    return flags
