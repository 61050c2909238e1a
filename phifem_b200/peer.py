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


class HaloExchange:
    """Contributions to rows owned by another rank, pushed into that rank's HBM over NVLink and added there by the same
    kernel (csrc/peer.cu, phifem_halo_exchange; SURVEY.md 8(b) export (5)).  `send_counts[q]` / `recv_counts[p]`: values
    this rank sends to q / receives from p per exchange; `send_src[q]`: first element of q's contiguous segment of the
    source array (for calls without a send index).  `torch.distributed` carries the IPC handles and the counts once."""

    def __init__(self, rank, world, send_counts, recv_counts, send_src=None, group=None):
        import torch.distributed as dist
        lib = _lib.load()
        self.rank, self.world = rank, world
        recv_counts = [int(c) for c in recv_counts]
        send_counts = [int(c) for c in send_counts]
        recv_off = [0] * (world + 1)
        for p in range(world):
            recv_off[p + 1] = recv_off[p] + recv_counts[p]
        self.n_recv, self.recv_offsets = recv_off[world], recv_off
        # everybody learns everybody's receive layout: capacity = the largest buffer, my offset inside each peer's
        tables = [None] * world
        dist.all_gather_object(tables, recv_off, group=group)
        capacity = max(t[world] for t in tables)
        self._h = ctypes.c_void_p()
        handle = ctypes.create_string_buffer(64)
        _lib.check(lib.phifem_halo_create(world, rank, capacity, ctypes.byref(self._h), handle))
        handles = [None] * world
        dist.all_gather_object(handles, handle.raw, group=group)
        blob = ctypes.create_string_buffer(b"".join(handles), 64 * world)
        I64 = ctypes.c_int64 * (world + 1)
        remote = I64(*([tables[q][rank] for q in range(world)] + [0]))
        ptr = [0] * (world + 1)
        for q in range(world):
            ptr[q + 1] = ptr[q] + send_counts[q]
        self.n_send = ptr[world]
        src0 = I64(*([int(v) for v in (send_src or [0] * world)] + [0]))
        _lib.check(lib.phifem_halo_connect(self._h, blob, remote, I64(*ptr), src0, self.n_recv))
        dist.barrier(group=group)          # everybody has mapped everybody before the first push

    def exchange(self, src, recv_index, dst, send_index=None):
        """dst[recv_index[i]] += value i received; values sent: src[send_index] (int64 device list ordered by peer) or,
        without it, the contiguous segments named at construction.  On the current stream, no host synchronisation."""
        _lib.check(_lib.load().phifem_halo_exchange(
            self._h, _lib.ptr(src), _lib.ptr(send_index) if send_index is not None and send_index.numel() else None,
            _lib.ptr(recv_index) if recv_index.numel() else None, _lib.ptr(dst), _lib.stream()))

    def timed_out(self):
        return bool(_lib.load().phifem_halo_error(self._h))

    def close(self):
        if self._h:
            _lib.load().phifem_halo_destroy(self._h)
            self._h = ctypes.c_void_p()


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
