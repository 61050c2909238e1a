"""Symbolic phase of the row-gather assembly (csrc/assemble_rows.cu, k_assemble_rows_p1).

Turns an `AssemblyPlan` (CSR pattern + entity lists + entity -> CSR-slot maps) into the arrays of
`phifem_rows_plan` (include/phifem_b200.h): for every row with pattern entries, the list of entities
touching its vertex, each entity written as the positions -- inside the row's own column list -- of its
other vertices.  Two lists -- cells (every pattern row) and surface (ghost-penalty facets and one-sided
entities, rows near the surface only) -- each stored as sliced ELLPACK over groups of 32 consecutive listed rows
so that a warp reads one coalesced 128-byte line per step.

Sort/scatter plumbing written with torch ops; runs on the device of the mesh (CPU tensors work too,
which is how the CPU tests check it against the oracle).
"""
import ctypes

import torch

from . import _lib

PAD = 0xFFFFFFFF
MAX_ROW_NNZ = 255            # positions are uint8
MAX_SMEM_BYTES = 200 * 1024  # accumulators of one CTA (128 threads x max_row_nnz doubles)
BLOCK = 128
BALANCE_CHUNK = 4096         # rows per spatially compact group of the surface lists


def morton_order(x, rows):
    """Listed rows sorted along a Morton curve of their vertex coordinates (21 bits per axis)."""
    from .mesh import morton_keys
    return rows[torch.argsort(morton_keys(x[rows]), stable=True)]


class RowList:
    """One record kind of the plan: the rows it touches (processing order), the position of each row's
    diagonal entry, and the records as sliced ELLPACK -- record k of lane l of slice s (= listed rows
    [32 s, 32 s + 32)) at rec[((ptr[s] + k) * 32 + l) * words]; pads are all-ones words; the records of a
    row keep their input order."""

    def __init__(self, rows, diag_slot, indptr, rec_rows, words, n_rows, balance=False):
        """rows [L] int64 (processing order); rec_rows [m] row id of each record; words [m, w].
        balance: re-sort the rows by record count (descending, stable) so that the 32 rows of a slice
        carry the same number of records -- for the short surface lists, whose data stays in L2 anyway."""
        dev = rows.device
        i64 = dict(dtype=torch.int64, device=dev)
        if balance and rows.numel():
            # `rows` arrive in a spatially compact order (Morton curve); inside chunks of BALANCE_CHUNK rows
            # they are re-sorted by record count, so that the 32 lanes of a slice carry (almost) the same number
            # of records while rows that share entities are still processed close in time (L2 reuse of the
            # facet-once records of the ghost-penalty pass)
            per_row = torch.bincount(rec_rows, minlength=n_rows)[rows]
            chunk = torch.arange(rows.numel(), **i64) // BALANCE_CHUNK
            key = chunk * (int(per_row.max()) + 1) + (int(per_row.max()) - per_row)
            rows = rows[torch.argsort(key, stable=True)]
        self.n_listed = int(rows.numel())
        self.n_slices = (self.n_listed + 31) // 32
        self.words = w = words.shape[1]
        self.n_records = int(rec_rows.numel())
        self.rows = rows.to(torch.int32).contiguous()
        self.diag_pos = (diag_slot[rows] - indptr[rows]).to(torch.uint8).contiguous()
        ns = self.n_slices
        li_of_row = torch.full((n_rows,), -1, **i64)
        li_of_row[rows] = torch.arange(self.n_listed, **i64)
        li = li_of_row[rec_rows]
        assert self.n_records == 0 or int(li.min()) >= 0, "a record names a row outside its list"
        counts = torch.bincount(li, minlength=ns * 32)[:ns * 32]
        width = counts.reshape(ns, 32).max(dim=1).values if ns else torch.zeros(0, **i64)
        ptr = torch.zeros(ns + 1, **i64)
        ptr[1:] = torch.cumsum(width, 0)
        total = int(ptr[-1])
        if total * 32 * w >= 2 ** 31:
            raise NotImplementedError("row-gather plan: record array exceeds 2^31 words")
        rec = torch.full((max(total, 1) * 32, w), PAD, **i64)
        if self.n_records:
            order = torch.argsort(li, stable=True)
            ls = li[order]
            first = torch.cumsum(counts, 0) - counts                 # first sorted record of each row
            k = torch.arange(ls.numel(), **i64) - first[ls]          # rank inside the row
            rec[(ptr[ls >> 5] + k) * 32 + (ls & 31)] = words[order]
        rec = rec.reshape(-1)
        rec = torch.where(rec >= 2 ** 31, rec - 2 ** 32, rec).to(torch.int32)   # same bits as uint32
        self.ptr, self.rec = ptr.to(torch.int32).contiguous(), rec.contiguous()

    def c_struct(self):
        p = _lib.ptr
        return _lib.CRowList(self.n_listed, p(self.rows), p(self.diag_pos), p(self.ptr), p(self.rec))

    def nbytes(self):
        return int(sum(t.numel() * t.element_size() for t in (self.rows, self.diag_pos, self.ptr, self.rec)))

    def padding(self):
        """Fraction of record slots that are pads (lanes idling in the record loop)."""
        return 1.0 - self.n_records * self.words / max(1, self.rec.numel())


def _pack_bytes(pos):
    """[m, k<=4] small non-negative ints -> one word per row, byte j = pos[:, j]."""
    word = torch.zeros(pos.shape[0], dtype=torch.int64, device=pos.device)
    for j in range(pos.shape[1]):
        word |= pos[:, j] << (8 * j)
    return word


def cell_geometry(x, cells):
    """[n, 8] fp64 geometry table of the cell pass (phifem_rows_plan.cell_geom): S_ab = |K| grad(lambda_a).grad(lambda_b)
    for a < b in lexicographic order, |K|, h_T^2 (CellDiameter^2 = largest squared vertex distance), zero padding.
    Elementwise torch ops only (one IEEE operation each, no library factorisation): the table of a cell depends on
    its own coordinates alone, so a rank of a sharded run tabulates bit for bit what a single GPU does."""
    n, nv = cells.shape
    d = nv - 1
    X = x[cells.long()]                                   # [n, nv, d]
    e = X[:, 1:, :] - X[:, :1, :]                         # edge k = X[k + 1] - X[0]
    if d == 2:
        R = [None, torch.stack([e[:, 1, 1], -e[:, 1, 0]], dim=1), torch.stack([-e[:, 0, 1], e[:, 0, 0]], dim=1)]
        det = e[:, 0, 0] * e[:, 1, 1] - e[:, 1, 0] * e[:, 0, 1]
        dfact = 2.0
    else:
        def cross(a, b):
            return torch.stack([a[:, 1] * b[:, 2] - a[:, 2] * b[:, 1], a[:, 2] * b[:, 0] - a[:, 0] * b[:, 2],
                                a[:, 0] * b[:, 1] - a[:, 1] * b[:, 0]], dim=1)
        R = [None, cross(e[:, 1], e[:, 2]), cross(e[:, 2], e[:, 0]), cross(e[:, 0], e[:, 1])]
        det = e[:, 0, 0] * R[1][:, 0] + e[:, 0, 1] * R[1][:, 1] + e[:, 0, 2] * R[1][:, 2]
        dfact = 6.0
    R[0] = -(R[1] + R[2]) if d == 2 else -(R[1] + R[2] + R[3])
    out = torch.zeros((n, 8), dtype=torch.float64, device=x.device)
    scale = 1.0 / (dfact * det.abs())                     # |K| / det^2
    k = 0
    for a in range(nv):
        for b in range(a + 1, nv):
            out[:, k] = (R[a] * R[b]).sum(dim=1) * scale
            k += 1
    out[:, k] = det.abs() / dfact
    h2 = torch.zeros(n, dtype=torch.float64, device=x.device)
    for a in range(nv):
        for b in range(a + 1, nv):
            h2 = torch.maximum(h2, ((X[:, a] - X[:, b]) ** 2).sum(dim=1))
    out[:, k + 1] = h2
    return out.contiguous()


class RowsPlan:
    """row_mask (bool [n_rows], optional): assemble only these rows -- the rows a rank owns in a multi-GPU
    run; records of other rows are dropped (their owner evaluates them from its own halo cells).
    geometry: tabulate the cells' P1 stiffness entries, |K| and h_T^2 once (`cell_geom`, 64 bytes per active cell) and
    let the cell pass read them instead of the coordinates (85 instead of 146 fp64 instructions per record; measured
    SLOWER on the B200 -- the table misses L1 where the coordinates hit -- hence off by default, see
    csrc/assemble_rows.cu)."""

    def __init__(self, plan, order="natural", row_mask=None, geometry=False, cell_pass="rows", rows_per_tile=256):
        """cell_pass: "rows" = the row-gather cell pass (one evaluation per (row, cell) record); "tiles" = the cell-once
        pass of csrc/assemble_tiles.cu (phifem_b200/tiles.py), whose rows are listed along the Morton curve unless
        `order` says otherwise or the mesh is already numbered along a space-filling curve (`mesh.sfc_ordered`)."""
        mesh = plan.mesh
        dev = mesh.device
        d = mesh.gdim
        nv = d + 1
        n = plan.n_rows
        i64 = dict(dtype=torch.int64, device=dev)
        self.plan = plan
        indptr = plan.indptr.long()
        row_nnz = indptr[1:] - indptr[:-1]
        has = row_nnz > 0
        if row_mask is not None:
            has = has & row_mask.to(dev)
        listed = torch.nonzero(has).reshape(-1)
        self.max_row_nnz = int(row_nnz[listed].max()) if listed.numel() else 0

        def owned(rec_rows, words):
            """Drop the records of rows outside the mask."""
            if row_mask is None or rec_rows.numel() == 0:
                return rec_rows, words
            keep = row_mask.to(dev)[rec_rows]
            return rec_rows[keep], words[keep]

        if self.max_row_nnz > MAX_ROW_NNZ or self.max_row_nnz * BLOCK * 8 > MAX_SMEM_BYTES:
            raise NotImplementedError(
                "row-gather assembly: a row holds %d entries (limit %d); use the atomic scatter kernels "
                "for this mesh" % (self.max_row_nnz, min(MAX_ROW_NNZ, MAX_SMEM_BYTES // (BLOCK * 8))))
        if order not in ("natural", "morton", "auto"):
            raise ValueError("order must be 'natural', 'morton' or 'auto'")
        if cell_pass not in ("rows", "tiles", "push"):
            raise ValueError("cell_pass must be 'rows', 'tiles' or 'push'")
        if cell_pass in ("tiles", "push") and (geometry or self.max_row_nnz > 128):
            cell_pass = "rows"        # positions of the tile records are 7 bits
        if order == "auto":
            order = "morton" if cell_pass in ("tiles", "push") and not getattr(mesh, "sfc_ordered", False) else "natural"
        self.order = order
        self.cell_pass = cell_pass

        def ordered(rows):
            """Unique row ids in processing order."""
            rows = torch.unique(rows)
            return morton_order(mesh.x, rows) if order == "morton" and rows.numel() else rows

        def surface_ordered(rows):
            """Rows of the surface lists along the Morton curve (then balanced chunk by chunk)."""
            rows = torch.unique(rows)
            return morton_order(mesh.x, rows) if rows.numel() else rows

        # diagonal slot: the pattern holds (r, r) for every listed row (each entity couples its vertices
        # with themselves)
        cols = plan.indices.long()
        row_of_slot = torch.repeat_interleave(torch.arange(n, **i64), row_nnz)
        is_diag = cols == row_of_slot
        dslot = torch.full((n,), -1, **i64)
        dslot[row_of_slot[is_diag]] = torch.nonzero(is_diag).reshape(-1)
        assert bool((dslot[listed] >= 0).all()), "a listed row has no diagonal entry"
        del row_of_slot, is_diag, cols

        # ---- cells: one record per (vertex of an active cell); list = every row of the pattern ------
        cells_act = mesh.cells[plan.active.long()].long()                       # [Na, nv]
        slots = plan.slots_cells.long().reshape(-1, nv, nv)                     # slot of (row i, col j)
        cut = (plan.cell_tags8[plan.active.long()] == 2).long()
        words = []
        cidx = torch.arange(cells_act.shape[0], **i64)
        for i in range(nv):
            others = [j for j in range(nv) if j != i]
            pos = slots[:, i, others] - indptr[cells_act[:, i]][:, None]
            w0 = _pack_bytes(pos) | (cut << 24)
            # geometry mode: the row's cell-local index (which row of the cached matrix) and the cell's table index
            words.append(torch.stack([w0 | (i << 25), cidx], dim=1) if geometry else w0[:, None])
        # interleaved so that the records of a row keep cell order (deterministic summation order)
        rec_rows, words = owned(cells_act.reshape(-1), torch.stack(words, dim=1).reshape(-1, 2 if geometry else 1))
        self.tiles = None
        if cell_pass in ("tiles", "push"):
            from .tiles import CellTiles
            trows = ordered(listed)
            self.tiles = CellTiles(mesh, plan.active, cut, plan.slots_cells, indptr, trows,
                                   (dslot[trows] - indptr[trows]).to(torch.uint8), rows_per_tile,
                                   push=cell_pass == "push")
            self.n_cell_records = int(rec_rows.numel())
            empty = torch.zeros(0, **i64)
            self.cells = RowList(empty, dslot, indptr, empty, torch.zeros((0, 1), **i64), n)
        else:
            self.cells = RowList(ordered(listed), dslot, indptr, rec_rows, words, n)
            self.n_cell_records = self.cells.n_records
        self.cell_geom = cell_geometry(mesh.x, cells_act) if geometry and cells_act.shape[0] else None
        del slots, cells_act, words

        # ---- ghost-penalty facets: one record per distinct vertex of the macro element ------------------
        from .assemble import ghost_macro_vertices
        ng = int(plan.ghost.numel())
        if ng:
            mac = ghost_macro_vertices(mesh, plan.ghost)                         # [ng, nv+1]
            gs = plan.slots_ghost.long().reshape(-1, nv + 1, nv + 1)
            if ng >= 2 ** 28:
                raise NotImplementedError("row-gather plan: more than 2^28 ghost-penalty facets")
            words = []
            gidx = torch.arange(ng, **i64)
            for a in range(nv + 1):     # the row's vertex is macro vertex a; others keep the macro order
                others = [j for j in range(nv + 1) if j != a]
                pos = gs[:, a, others] - indptr[mac[:, a]][:, None]
                words.append(torch.stack([_pack_bytes(pos), gidx | (a << 28)], dim=1))
            rec_rows, words = mac.reshape(-1), torch.stack(words, dim=1).reshape(-1, 2)
            self.ghost_macro = mac.to(torch.int32).contiguous()
        else:
            rec_rows, words = torch.zeros(0, **i64), torch.zeros((0, 2), **i64)
            self.ghost_macro = torch.zeros((0, nv + 1), dtype=torch.int32, device=dev)
        self.n_ghost_facets = ng
        ghost_rows, ghost_words = owned(rec_rows, words)

        # ---- one-sided facets: one record per facet vertex of each (cell, local facet) entity --------------
        ne = int(plan.entities.shape[0])
        self.n_entities = ne
        if ng + ne >= 2 ** 28:
            raise NotImplementedError("row-gather plan: more than 2^28 surface entities")
        if ne:
            ec = plan.entities[:, 0].long()
            eo = plan.entities[:, 1].long()
            ev = mesh.cells[ec].long()
            bs = plan.slots_boundary.long().reshape(-1, nv, nv)
            ar = torch.arange(ne, **i64)
            loc = torch.arange(nv, **i64)[None, :].expand(ne, nv)
            fac = loc[loc != eo[:, None]].reshape(ne, d)                # local ids of the facet vertices
            rec_rows, words = [], []
            for t in range(d):                      # the row's vertex = t-th facet vertex
                i = fac[:, t]
                rest = fac[:, [(t + u) % d for u in range(1, d)]]       # the others, rotated: (t + 1) % d, ...
                oth = torch.cat([eo[:, None], rest], dim=1)             # opposite vertex first
                row = ev[ar, i]
                pos = bs[ar[:, None], i[:, None], oth] - indptr[row][:, None]
                words.append(torch.stack([_pack_bytes(pos), (ng + ar) | (t << 28) | (1 << 31)], dim=1))
                rec_rows.append(row)
            rec_rows = torch.stack(rec_rows, dim=1).reshape(-1)
            words = torch.stack(words, dim=1).reshape(-1, 2)
            # vertex ids [facet vertices in ascending local order, opposite vertex]
            self.entity_macro = torch.cat([ev.gather(1, fac), ev.gather(1, eo[:, None])], dim=1) \
                .to(torch.int32).contiguous()
        else:
            rec_rows, words = torch.zeros(0, **i64), torch.zeros((0, 2), **i64)
            self.entity_macro = torch.zeros((0, nv), dtype=torch.int32, device=dev)
        ent_rows, ent_words = owned(rec_rows, words)
        # one surface list: a row's ghost-penalty records first, then its one-sided records
        rec_rows = torch.cat([ghost_rows, ent_rows])
        words = torch.cat([ghost_words, ent_words])
        self.n_ghost_records, self.n_entity_records = int(ghost_rows.numel()), int(ent_rows.numel())
        self.surface = RowList(surface_ordered(rec_rows), dslot, indptr, rec_rows, words, n, balance=True)
        # facet-once scratch of the surface pass (8 doubles = 64 bytes per ghost facet / entity), rewritten by
        # every assembly
        self.surface_work = torch.empty((max(ng + ne, 1), 8), dtype=torch.float64, device=dev)
        # mesh-only part of those records (csrc/assemble_rows.cu k_surface_static_p1), tabulated once below
        self.surface_static = torch.zeros((max(ng + ne, 1), 8), dtype=torch.float64, device=dev) if dev.type == "cuda" \
            else None
        self._c = None
        if self.surface_static is not None and ng + ne > 0:
            _lib.check(_lib.load().phifem_surface_static_p1(_lib.c_mesh(mesh), ctypes.byref(self.c_struct()),
                                                            _lib.stream()))

    def c_struct(self, passes=None):
        """`passes`: subset of ("cells", "surface") to run (bench.py times them one by one); the other list is
        passed empty."""
        if passes is not None:
            p = _lib.ptr
            empty = _lib.CRowList(0, None, None, None, None)
            lists = [getattr(self, nm).c_struct() if nm in passes else empty for nm in ("cells", "surface")]
            return _lib.CRowsPlan(p(self.plan.indptr), p(self.plan.indices), self.max_row_nnz, 0, *lists,
                                  *self._surface_fields(cells="cells" in passes))
        if self._c is None:
            p = _lib.ptr
            pl = self.plan
            self._c = _lib.CRowsPlan(p(pl.indptr), p(pl.indices), self.max_row_nnz, 0,
                                     self.cells.c_struct(), self.surface.c_struct(), *self._surface_fields())
        return self._c

    def _surface_fields(self, cells=True):
        """Trailing fields of phifem_rows_plan: ghost facet / entity vertex lists, the facet-once scratch, the cached
        cell geometry."""
        if self.surface_work.device.type != "cuda":
            return 0, None, 0, None, None, None, None, None
        return (self.n_ghost_facets, _lib.ptr(self.ghost_macro) if self.n_ghost_facets else None,
                self.n_entities, _lib.ptr(self.entity_macro) if self.n_entities else None,
                _lib.ptr(self.surface_work), _lib.ptr(self.cell_geom),
                ctypes.pointer(self.tiles.c_struct()) if self.tiles is not None and cells else None,
                _lib.ptr(self.surface_static))

    def index_bytes(self):
        """Bytes of plan arrays one numeric pass streams besides the CSR pattern itself."""
        return (self.cells.nbytes() + (self.tiles.nbytes() if self.tiles is not None else 0)
                + self.surface.nbytes() + self.ghost_macro.numel() * 4
                + self.entity_macro.numel() * 4 + self.surface_work.numel() * 8
                + (self.surface_static.numel() * 8 if self.surface_static is not None else 0)
                + (self.cell_geom.numel() * 8 if self.cell_geom is not None else 0))


class _NativeList:
    """A row list of a natively built plan, with the attributes of `RowList` that callers read."""

    def __init__(self, c_list, n_records, slots, words, device, keep):
        self.c = c_list
        self.n_listed = int(c_list.n_listed)
        self.n_slices = (self.n_listed + 31) // 32
        self.words, self.n_records, self._slots = words, int(n_records), int(slots)
        def dv(*a):
            return _lib.device_view(*a, owner=keep)
        self.rows = dv(c_list.rows, (self.n_listed,), torch.int32, device)
        self.diag_pos = dv(c_list.diag_pos, (self.n_listed,), torch.uint8, device)
        self.ptr = dv(c_list.ptr, (self.n_slices + 1,), torch.int32, device)
        self.rec = dv(c_list.rec, (max(self._slots, 32) * words,), torch.int32, device)
        self._keep = keep

    def c_struct(self):
        return self.c

    def nbytes(self):
        return self.n_listed * 5 + (self.n_slices + 1) * 4 + max(self._slots, 32) * self.words * 4

    def padding(self):
        return 1.0 - self.n_records / max(1, self._slots)


class _Handle:
    def __init__(self, pointer):
        self.pointer = pointer

    def __del__(self):
        try:
            if self.pointer:
                _lib.load().phifem_rows_plan_destroy(self.pointer)
        except Exception:   # interpreter shutdown
            pass
        self.pointer = None


class NativeRowsPlan:
    """The arrays of `RowsPlan` built on the device behind the C ABI (csrc/rows_plan.cu, phifem_rows_plan_create): the
    CSR pattern and both row lists from the vertex -> cell adjacency, bit for bit what the torch passes of
    `AssemblyPlan` + `RowsPlan` produce (tests/test_gpu_rows_plan_capi.py)."""

    def __init__(self, mesh, cell_tags8, facet_tags8, entities, order="natural", row_mask=None):
        _lib.require_cuda(mesh)
        lib = _lib.load()
        dev = mesh.device
        ents = entities.reshape(-1, 2).to(torch.int32).contiguous()
        mask = None if row_mask is None else row_mask.to(dev).to(torch.uint8).contiguous()
        out = ctypes.c_void_p()
        _lib.check(lib.phifem_rows_plan_create(_lib.c_mesh(mesh), _lib.ptr(cell_tags8), _lib.ptr(facet_tags8),
                                               _lib.ptr(ents) if ents.numel() else None, ents.shape[0],
                                               _lib.ptr(mask), 1 if order == "morton" else 0, ctypes.byref(out),
                                               _lib.stream()))
        self._handle = _Handle(out.value)
        self._c, info = _lib.CRowsPlan(), _lib.CRowsPlanInfo()
        _lib.check(lib.phifem_rows_plan_view(out.value, ctypes.byref(self._c), ctypes.byref(info)))
        self.info = info
        self.mesh, self.order, self.cell_pass, self.tiles, self.cell_geom = mesh, order, "rows", None, None
        nv = mesh.cells.shape[1]

        def dv(*a):
            return _lib.device_view(*a, owner=self._handle)
        self.n_rows, self.nnz = int(info.n_rows), int(info.nnz)
        self.indptr = dv(self._c.indptr, (self.n_rows + 1,), torch.int32, dev)
        self.indices = dv(self._c.indices, (self.nnz,), torch.int32, dev)
        self.active = dv(info.active, (int(info.n_active),), torch.int32, dev)
        self.ghost = dv(info.ghost, (int(info.n_ghost),), torch.int32, dev)
        self.max_row_nnz = int(self._c.max_row_nnz)
        self.n_ghost_facets, self.n_entities = int(info.n_ghost), int(info.n_entities)
        self.ghost_macro = dv(self._c.ghost_macro, (self.n_ghost_facets, nv + 1), torch.int32, dev)
        self.entity_macro = dv(self._c.entity_macro, (self.n_entities, nv), torch.int32, dev)
        self.surface_work = dv(self._c.surface_work, (max(self.n_ghost_facets + self.n_entities, 1), 8), torch.float64,
                               dev)
        self.surface_static = dv(self._c.surface_static, (max(self.n_ghost_facets + self.n_entities, 1), 8),
                                 torch.float64, dev)
        n_cell_rec = int(info.n_cell_records)
        if mask is not None:     # records of the listed rows only
            cnt = (self.indptr[1:] > self.indptr[:-1]) & mask.bool()
            deg = torch.bincount(mesh.cells[self.active.long()].long().reshape(-1), minlength=self.n_rows)
            n_cell_rec = int(deg[cnt].sum())
        self.n_cell_records = n_cell_rec
        self.cells = _NativeList(self._c.cells, n_cell_rec, info.cells_record_slots, 1, dev, self._handle)
        self.surface = _NativeList(self._c.surface, info.n_surface_records, info.surface_record_slots, 2, dev,
                                   self._handle)
        self.n_ghost_records = self.n_ghost_facets * (nv + 1) if mask is None else None
        self.n_entity_records = self.n_entities * (nv - 1) if mask is None else None

    def c_struct(self, passes=None):
        if passes is None:
            return self._c
        c = _lib.CRowsPlan()
        ctypes.pointer(c)[0] = self._c
        empty = _lib.CRowList(0, None, None, None, None)
        if "cells" not in passes:
            c.cells = empty
        if "surface" not in passes:
            c.surface = empty
        return c

    def index_bytes(self):
        return (self.cells.nbytes() + self.surface.nbytes() + self.ghost_macro.numel() * 4
                + self.entity_macro.numel() * 4 + self.surface_work.numel() * 8 + self.surface_static.numel() * 8)


def assemble_rows_into(rplan, phi, f, sigma, data, b, passes=None):
    """Numeric phase on the current stream (facet-once kernel, cell pass, surface pass).  `data` needs
    no zero-fill; `b` must have been zeroed once (rows without pattern entries are never written)."""
    mesh = rplan.mesh if isinstance(rplan, NativeRowsPlan) else rplan.plan.mesh
    _lib.require_cuda(mesh)
    _lib.check(_lib.load().phifem_assemble_rows_p1(
        _lib.c_mesh(mesh), _lib.ptr(phi), _lib.ptr(f), float(sigma),
        ctypes.byref(rplan.c_struct(passes)), _lib.ptr(data), _lib.ptr(b), _lib.stream()))
    return data, b
