"""Symbolic phase of the owner-computes assembly (csrc/assemble.cu, k_assemble_blocked_p1).

Turns an `AssemblyPlan` (entity lists + entity -> CSR-slot maps) into the arrays of
`phifem_blocked_plan` (include/phifem_b200.h):

  1. rows with contributions are binned into spatially compact boxes of vertices (so that the cells
     touching a block are mostly interior to it) and bins are cut into blocks holding at most
     `capacity` contributions (= doubles of shared memory);
  2. every contribution (entity, local entry) is keyed by (block of its row, destination) and the keys
     are sorted once (stable): contributions to the same CSR entry become adjacent, the rank inside
     the block is the shared-memory position the kernel writes to, run boundaries are the segments the
     reduction threads sum;
  3. an entity is listed once per block it touches ("instance") with the positions of its entries in
     that block (-1 where the row belongs to another block, which recomputes the entity itself).

Sort/unique plumbing written with torch ops; runs on the device of the mesh (CPU tensors work too,
which is how the CPU tests check it against the oracle).
"""
import ctypes

import numpy as np
import torch

from . import _lib

DESC_INTS = 12
DEFAULT_CAPACITY = 25000     # doubles: 200 KB buffer + ~20 KB segment tables of the 227 KB a CTA may use on sm_100a
_KEY_SHIFT = 33              # key = block << 33 | is_b << 32 | index


CBlockedPlan = _lib.CBlockedPlan


def _bin_shape(rows_target, d):
    """Vertices per bin side (s_1..s_d), near-cubic, product <= rows_target."""
    s = max(1, int(rows_target ** (1.0 / d)))
    shape = [s] * d
    for k in range(d):
        trial = list(shape)
        trial[k] += 1
        if np.prod(trial) <= rows_target:
            shape = trial
    return shape


def _pack_positions(pos):
    """[n_inst, m] int16 -> column-major 32-bit words [ceil(m/2), n_inst] viewed as int16 memory."""
    n, m = pos.shape
    if m % 2:
        pos = torch.cat([pos, torch.full((n, 1), -1, dtype=torch.int16, device=pos.device)], dim=1)
    words = pos.reshape(n, pos.shape[1] // 2, 2)          # [n, m/2, 2]: low half first (little endian)
    return words.permute(1, 0, 2).contiguous()            # [m/2, n, 2]


class BlockedPlan:
    def __init__(self, plan, capacity=DEFAULT_CAPACITY):
        mesh = plan.mesh
        dev = mesh.device
        d = mesh.gdim
        nv = d + 1
        n, nnz = plan.n_rows, plan.nnz
        self.plan, self.capacity_request = plan, int(capacity)
        i64 = dict(dtype=torch.int64, device=dev)
        cells_act = mesh.cells[plan.active.long()].long()                    # [Na, nv]
        g_c, g_g, g_b = plan.slots_cells.long(), plan.slots_ghost.long(), plan.slots_boundary.long()
        row_of_slot = torch.repeat_interleave(torch.arange(n, **i64), torch.diff(plan.indptr.long()))

        # ---- 1. rows -> blocks -------------------------------------------------------------------
        mult = torch.bincount(torch.cat([g_c.reshape(-1), g_g.reshape(-1), g_b.reshape(-1)]), minlength=nnz)
        crow = torch.zeros(n, **i64).index_add_(0, row_of_slot, mult)
        crow += torch.bincount(cells_act.reshape(-1), minlength=n)
        act_rows = torch.nonzero(crow > 0).reshape(-1)
        if act_rows.numel() == 0:
            raise ValueError("nothing to assemble: no active cell")
        max_row = int(crow.max())
        if max_row >= capacity or capacity > 32767:
            raise NotImplementedError(
                "owner-computes assembly: a single row gathers %d contributions (block capacity %d, "
                "hard limit 32767); use the atomic scatter kernels for this mesh" % (max_row, capacity))
        cap_eff = capacity - max_row
        cbar = float(crow[act_rows].double().mean())
        shape = _bin_shape(max(1.0, cap_eff / (1.10 * cbar)), d)
        lo, hi = mesh.x.min(dim=0).values, mesh.x.max(dim=0).values
        h = float(((hi - lo).clamp(min=1e-300).prod() / max(1, mesh.num_vertices)) ** (1.0 / d))
        side = torch.tensor([s * h for s in shape], dtype=torch.float64, device=dev)
        bins = torch.floor((mesh.x[act_rows] - lo) / side + 1e-6).long()
        nb_dim = bins.max(dim=0).values + 1
        bin_id = bins[:, 0]
        for k in range(1, d):
            bin_id = bin_id * nb_dim[k] + bins[:, k]
        order = torch.argsort(bin_id * n + act_rows)                          # by bin, then row id
        rows_s, bin_s, c_s = act_rows[order], bin_id[order], crow[act_rows][order]
        excl = torch.cumsum(c_s, 0) - c_s
        first = torch.ones_like(bin_s, dtype=torch.bool)
        first[1:] = bin_s[1:] != bin_s[:-1]
        bin_dense = torch.cumsum(first.long(), 0) - 1
        bin_start = excl[first]
        chunk = (excl - bin_start[bin_dense]) // cap_eff
        bkey = bin_dense * (int(chunk.max()) + 1) + chunk
        newb = torch.ones_like(bkey, dtype=torch.bool)
        newb[1:] = bkey[1:] != bkey[:-1]
        block_s = torch.cumsum(newb.long(), 0) - 1
        NB = int(block_s[-1]) + 1
        block_of_row = torch.full((n,), -1, **i64)
        block_of_row[rows_s] = block_s
        self.n_blocks, self.bin_shape, self.block_of_row = NB, shape, block_of_row

        # ---- 2. contributions -> sorted keys -> positions, segments ----------------------------------
        def mat_keys(g):
            return (block_of_row[row_of_slot[g]] << _KEY_SHIFT) | g

        keys_cell = torch.cat([mat_keys(g_c), (block_of_row[cells_act] << _KEY_SHIFT) | (1 << 32) | cells_act],
                              dim=1)                                           # [Na, nv*nv + nv]
        keys_g, keys_b = mat_keys(g_g), mat_keys(g_b)
        del g_c, g_g, g_b
        sizes = [keys_cell.numel(), keys_g.numel(), keys_b.numel()]
        allk = torch.cat([keys_cell.reshape(-1), keys_g.reshape(-1), keys_b.reshape(-1)])
        sorted_k, perm = torch.sort(allk, stable=True)
        rank = torch.empty_like(perm)
        rank[perm] = torch.arange(perm.numel(), **i64)
        del perm
        bounds = torch.arange(NB + 1, **i64) << _KEY_SHIFT
        block_first = torch.searchsorted(sorted_k, bounds)                     # first contribution of a block
        local_pos = rank - block_first[allk >> _KEY_SHIFT]
        del rank
        n_contrib = block_first[1:] - block_first[:-1]
        self.capacity = int(n_contrib.max())
        assert self.capacity <= capacity <= 32767, (self.capacity, capacity)
        uniq_k, counts = torch.unique_consecutive(sorted_k, return_counts=True)
        seg_global = torch.cumsum(counts, 0) - counts
        seg_block = uniq_k >> _KEY_SHIFT
        low = uniq_k & ((1 << _KEY_SHIFT) - 1)
        is_b = low >> 32
        dest = (low & 0xFFFFFFFF) - (is_b << 31)           # b rows: index - 2^31 (sign bit set, low bits = row)
        seg_first = torch.searchsorted(uniq_k, bounds)
        # per-block segment tables, padded to a multiple of 8 entries with at least one pad: the kernel
        # stages them with 16-byte cp.async chunks and reads start[s+1] as the end of segment s
        n_seg = seg_first[1:] - seg_first[:-1]
        n_pad = ((n_seg + 1 + 7) // 8) * 8
        pad_begin = torch.cumsum(n_pad, 0) - n_pad
        total_pad = int(n_pad.sum())
        self.max_segments = int(n_pad.max())
        where = pad_begin[seg_block] + (torch.arange(uniq_k.numel(), **i64) - seg_first[seg_block])
        seg_start = torch.repeat_interleave(n_contrib, n_pad).to(torch.int16)      # pads: n_contrib
        seg_start[where] = (seg_global - block_first[seg_block]).to(torch.int16)
        seg_dest = torch.zeros(total_pad, dtype=torch.int32, device=dev)
        seg_dest[where] = dest.to(torch.int32)
        self.seg_start, self.seg_dest = seg_start.contiguous(), seg_dest.contiguous()
        del sorted_k, uniq_k, counts
        pos_cell, pos_g, pos_b = torch.split(local_pos.to(torch.int16), sizes)
        blk_cell, blk_g, blk_b = torch.split(allk >> _KEY_SHIFT, sizes)
        del allk, local_pos

        # ---- 3. instances ---------------------------------------------------------------------------
        def instances(blk, pos, m):
            """blk/pos: flat [E*m] -> (entity of instance, block of instance, positions [n_inst, m])
            sorted by block, plus the per-block ranges."""
            E = blk.numel() // m
            if E == 0:
                z = torch.zeros(0, **i64)
                return z, torch.full((0, m), -1, dtype=torch.int16, device=dev), torch.zeros(NB + 1, **i64)
            ent = torch.arange(E, **i64).repeat_interleave(m)
            uk, inv = torch.unique(ent * NB + blk, return_inverse=True)
            inst_ent, inst_blk = uk // NB, uk % NB
            p = torch.full((uk.numel(), m), -1, dtype=torch.int16, device=dev)
            p[inv, torch.arange(m, **i64).repeat(E)] = pos
            o = torch.argsort(inst_blk, stable=True)
            return inst_ent[o], p[o], torch.searchsorted(inst_blk[o].contiguous(), torch.arange(NB + 1, **i64))

        m_c, m_g, m_b = nv * nv + nv, (nv + 1) ** 2, nv * nv
        ce, cp, c_rng = instances(blk_cell, pos_cell, m_c)
        ge, gp, g_rng = instances(blk_g, pos_g, m_g)
        be, bp, b_rng = instances(blk_b, pos_b, m_b)
        verts = torch.zeros((ce.numel(), 4), dtype=torch.int32, device=dev)
        verts[:, :nv] = cells_act[ce].to(torch.int32)
        cut = plan.cell_tags8[plan.active.long()[ce]] == 2
        verts[:, 0] |= torch.where(cut, torch.tensor(-2 ** 31, dtype=torch.int32, device=dev),
                                   torch.tensor(0, dtype=torch.int32, device=dev))
        self.cell_verts = verts.contiguous()
        self.cell_pos = _pack_positions(cp)
        self.ghost_facet = plan.ghost[ge].contiguous()
        self.ghost_pos = _pack_positions(gp)
        self.bnd_entity = plan.entities[be].contiguous()
        self.bnd_pos = _pack_positions(bp)
        self.n_cell_inst, self.n_ghost_inst, self.n_bnd_inst = int(ce.numel()), int(ge.numel()), int(be.numel())

        desc = torch.zeros((NB, DESC_INTS), dtype=torch.int32, device=dev)
        desc[:, 0] = n_contrib
        desc[:, 1], desc[:, 2] = pad_begin, n_pad
        desc[:, 3], desc[:, 4] = c_rng[:-1], c_rng[1:]
        desc[:, 5], desc[:, 6] = g_rng[:-1], g_rng[1:]
        desc[:, 7], desc[:, 8] = b_rng[:-1], b_rng[1:]
        self.block_desc = desc.contiguous()
        self.redundancy = self.n_cell_inst / max(1, plan.active.numel())
        self._c = None

    def c_struct(self):
        if self._c is None:
            p = _lib.ptr
            self._c = CBlockedPlan(self.n_blocks, self.capacity, self.max_segments, 0,
                                   p(self.block_desc), p(self.seg_start),
                                   p(self.seg_dest), self.n_cell_inst, p(self.cell_verts), p(self.cell_pos),
                                   self.n_ghost_inst, p(self.ghost_facet), p(self.ghost_pos),
                                   self.n_bnd_inst, p(self.bnd_entity), p(self.bnd_pos))
        return self._c

    def index_bytes(self):
        """Bytes of plan arrays one numeric pass streams (overhead on top of the algorithmic bytes)."""
        ts = (self.block_desc, self.seg_start, self.seg_dest, self.cell_verts, self.cell_pos,
              self.ghost_facet, self.ghost_pos, self.bnd_entity, self.bnd_pos)
        return int(sum(t.numel() * t.element_size() for t in ts))


def assemble_blocked_into(bplan, phi, f, sigma, data, b):
    """Numeric phase, one launch on the current stream.  `data` needs no zero-fill; `b` must have been
    zeroed once (rows without contributions are never written)."""
    mesh = bplan.plan.mesh
    _lib.require_cuda(mesh)
    _lib.check(_lib.load().phifem_assemble_blocked_p1(
        _lib.c_mesh(mesh), _lib.ptr(phi), _lib.ptr(f), float(sigma), ctypes.byref(bplan.c_struct()),
        _lib.ptr(data), _lib.ptr(b), _lib.stream()))
    return data, b
