"""Multi-GPU sharding of the hot path: one process per GPU, owned CSR rows, halo-row exchange.

The reference has no communication code of its own (dolfinx/PETSc do the MPI work implicitly;
reference src/phifem/mesh_scripts.py:264 even carries a TODO for parallel tag transfer).  Here the
path shards the way SURVEY.md section 8(e) lays out:

  * cells are split into contiguous ranges of the global cell order (slabs of a box for the
    synthetic benchmark), one range per rank; every rank also holds ONE ghost layer of cells, which
    it classifies redundantly so that the facets on the partition boundary see both cell tags
    without communication;
  * one 8-byte all-reduce makes "is there any exterior cell" global (the facet algebra changes when
    there is none, mesh_scripts.py:469-474);
  * rows are owned by the lowest rank touching them.  Two numeric strategies:
    mode="rows" (default of bench.py): OWNER COMPUTES -- every rank runs the row-gather kernels on the
      rows it owns; the entities touching those rows reach at most two cell layers into the next slab
      (a cell holding an owned boundary vertex, and its facet neighbour for the ghost penalty), so the
      rank keeps two redundantly classified ghost layers on that side and NOTHING is exchanged in the
      numeric phase: no halo buffers, no atomics, bitwise identical to the single-GPU rows;
    mode="exchange": every rank assembles its owned cells, the ghost-penalty facets whose first cell it
      owns and the one-sided entities of its owned cells with the per-entity kernels; contributions to
      rows owned by another rank are written straight into that rank's send segment (the slot maps point
      there), so the exchange is pack-free: grouped send/recv of the segments (NCCL over NVLink), then
      one indexed add on the owner.

The symbolic phase (pattern union across ranks, slot maps, send/recv lists) uses point-to-point
transfers and works on the `gloo` backend too, which is how the CPU tests cover it.
"""
from types import SimpleNamespace

import torch
import torch.distributed as dist

from . import assemble, mesh_scripts, synthetic
from .mesh import Mesh

TUBE_RADIUS = 0.15


def _exchange(tensors_out, dtype, device, group=None):
    """All-to-all of variable-length 1-D tensors with point-to-point ops (NCCL and gloo)."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    sizes = torch.tensor([int(t.numel()) for t in tensors_out], dtype=torch.int64, device=device)
    all_sizes = [torch.empty_like(sizes) for _ in range(world)]
    dist.all_gather(all_sizes, sizes, group=group)
    recv = [torch.empty(int(all_sizes[p][rank]), dtype=dtype, device=device) for p in range(world)]
    ops = []
    for p in range(world):
        if p == rank:
            continue
        if tensors_out[p].numel():
            ops.append(dist.P2POp(dist.isend, tensors_out[p].contiguous(), p, group=group))
        if recv[p].numel():
            ops.append(dist.P2POp(dist.irecv, recv[p], p, group=group))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    recv[rank] = tensors_out[rank]
    return recv


def entities_host(mesh, cell_tags8, facet_tags8):
    """CPU stand-in of the ds(100) entity search for the gloo tests (same ordering rules, torch ops)."""
    f = torch.nonzero(facet_tags8 == 4).reshape(-1)
    pairs = []
    for col in (1, 0):                    # reversed link order (mesh_scripts.py:195-214)
        c = mesh.f2c[f, col].long()
        ok = c >= 0
        ok &= ((cell_tags8[c.clamp(min=0)] == 1) | (cell_tags8[c.clamp(min=0)] == 2))
        cc, ff = c[ok], f[ok]
        lf = (mesh.c2f[cc].long() == ff[:, None]).long().argmax(dim=1)
        pairs.append(torch.stack([cc, lf], dim=1))
    return torch.cat(pairs).to(torch.int32)


class SlabProblem:
    """Synthetic weak-scaling problem: the box [0, world] x [0,1]^2 of world*n x n x n cubes (6 Kuhn
    tetrahedra each), rank r owning the cubes of [r, r+1].  Level set: one sphere (radius 0.45) per
    unit cube joined by a thin tube along x, so the partition boundaries cut through active cells and
    the halo exchange carries real entries."""

    def __init__(self, n, rank, world, device, group=None, mode="exchange"):
        if mode not in ("exchange", "rows"):
            raise ValueError("mode must be 'exchange' or 'rows'")
        self.n, self.rank, self.world, self.group, self.mode = n, rank, world, group, mode
        self.device = torch.device(device)
        gl = 1 if rank > 0 else 0
        gr = (min(2, n) if mode == "rows" else 1) if rank < world - 1 else 0
        nx = n + gl + gr
        i0 = rank * n - gl
        self.mesh = self._slab_mesh(nx, n, i0)
        dev = self.device
        # global vertex ids, owners
        plane = (n + 1) * (n + 1)
        iloc = torch.arange(nx + 1, device=dev, dtype=torch.int64).repeat_interleave(plane)
        gi = iloc + i0
        self.global_vertex = gi * plane + torch.arange(plane, device=dev).repeat(nx + 1)
        self.n_global_vertices = (world * n + 1) * plane
        self.vertex_owner = torch.where(gi == 0, torch.zeros_like(gi), (gi - 1) // n)
        self.row_lo = 0 if rank == 0 else (rank * n + 1) * plane
        self.row_hi = ((rank + 1) * n + 1) * plane          # exclusive
        # owned cells: cube layers [gl, gl + n)
        cube = torch.arange(self.mesh.num_cells, device=dev, dtype=torch.int64) // (6 * n * n)
        self.cell_owned = (cube >= gl) & (cube < gl + n)
        self.n_owned_cells = int(self.cell_owned.sum())
        self.phi = self._levelset(self.mesh.x)
        self.f = self._source(self.mesh.x)
        self.plan = None
        self.peer = None

    @classmethod
    def global_reference(cls, n, world, device="cpu"):
        """(mesh, phi, f) of the WHOLE box on one device: what the ranks' results are checked against.
        Vertex / cell numbering equals the ranks' global numbering."""
        self = cls.__new__(cls)
        self.n, self.rank, self.world, self.group, self.device = n, 0, world, None, torch.device(device)
        mesh = self._slab_mesh(world * n, n, 0)
        return mesh, self._levelset(mesh.x), self._source(mesh.x)

    def _slab_mesh(self, nx, n, i0):
        base = synthetic.box_mesh((nx, n, n), device=self.device)
        x = base.x.clone()
        idx = torch.arange(nx + 1, device=self.device, dtype=torch.float64).repeat_interleave((n + 1) ** 2)
        x[:, 0] = (idx + i0) / n          # exact function of the GLOBAL index: identical on every rank
        m = Mesh(x, base.cells, "tetrahedron", self.device)
        return m

    def _levelset(self, x):
        cx, cy, cz = synthetic.SPHERE_CENTER
        r2 = synthetic.SPHERE_RADIUS ** 2
        slab = torch.clamp(torch.floor(x[:, 0]), 0, self.world - 1)
        out = None
        for off in (-1.0, 0.0, 1.0):
            c = torch.clamp(slab + off, 0, self.world - 1)
            d = (x[:, 0] - (c + cx)) ** 2 + (x[:, 1] - cy) ** 2 + (x[:, 2] - cz) ** 2 - r2
            out = d if out is None else torch.minimum(out, d)
        tube = (x[:, 1] - cy) ** 2 + (x[:, 2] - cz) ** 2 - TUBE_RADIUS ** 2
        return torch.minimum(out, tube).contiguous()

    def _source(self, x):
        frac = x.clone()
        frac[:, 0] = x[:, 0] - torch.clamp(torch.floor(x[:, 0]), 0, self.world - 1)
        return synthetic.ball_source(frac)

    # ---- tags --------------------------------------------------------------------------------------
    def classify(self, dls, ws, mark=None):
        """Cells, interior facets overlapped with the 8-byte all-reduce of the exterior-cell count, mesh-boundary
        facets (all on the current stream; mesh_scripts.classify_sharded)."""
        return mesh_scripts.classify_sharded(self.mesh, dls, ws, group=self.group, world=self.world, mark=mark,
                                             peer=getattr(self, "peer", None))
        # (`single_layer_cut` needs one more ghost layer: offered by partition.PartitionedProblem, not by the slabs)

    def enable_peer_flags(self):
        """Exterior-cell counts through NVLink peer memory instead of an all-reduce (phifem_b200/peer.py)."""
        from . import peer
        self.peer = peer.try_create(self.rank, self.world, self.group)
        return self.peer is not None

    # ---- symbolic phase ----------------------------------------------------------------------------
    def build_plan(self, cell_tags8, facet_tags8):
        if self.mode == "rows":
            return self._build_rows_plan(cell_tags8, facet_tags8)
        mesh, dev, rank, world = self.mesh, self.device, self.rank, self.world
        NG = self.n_global_vertices
        gv = self.global_vertex
        nv = 4
        owned_c = self.cell_owned
        active = torch.nonzero(owned_c & ((cell_tags8 == 1) | (cell_tags8 == 2))).reshape(-1)
        interior = mesh.f2c[:, 1] >= 0
        first_owned = owned_c[mesh.f2c[:, 0].long()]
        ghost = torch.nonzero(((facet_tags8 == 2) | (facet_tags8 == 3)) & interior & first_owned).reshape(-1)
        if rank >= 0 and mesh.device.type == "cuda":
            ents = mesh_scripts._integration_entities_dev(mesh, cell_tags8, facet_tags8, 4, (1, 2)).reshape(-1, 2)
        else:
            ents = self._entities_host(cell_tags8, facet_tags8)
        ents = ents[owned_c[ents[:, 0].long()]].contiguous()

        def pair_keys(dm):
            g = gv[dm]
            return (g[:, :, None] * NG + g[:, None, :]).reshape(dm.shape[0], dm.shape[1] ** 2)

        dm_c = mesh.cells[active].long()
        mac = assemble.ghost_macro_vertices(mesh, ghost)
        dm_b = mesh.cells[ents[:, 0].long()].long()
        keys = [pair_keys(dm_c), pair_keys(mac), pair_keys(dm_b)]
        flat = torch.cat([k.reshape(-1) for k in keys])
        plane = (self.n + 1) ** 2
        gi = (flat // NG) // plane
        owner = torch.where(gi == 0, torch.zeros_like(gi), (gi - 1) // self.n)
        send_keys = [torch.unique(flat[owner == q]) if q != rank else flat.new_zeros(0) for q in range(world)]
        recv_keys = _exchange(send_keys, torch.int64, dev, self.group) if world > 1 else [flat.new_zeros(0)]
        own = torch.unique(torch.cat([flat[owner == rank]] + [recv_keys[p] for p in range(world) if p != rank]))
        nnz = int(own.numel())
        send_off, off = [], nnz
        for q in range(world):
            send_off.append(off)
            off += int(send_keys[q].numel())
        total = off
        slot = torch.empty_like(flat)
        mine = owner == rank
        slot[mine] = torch.searchsorted(own, flat[mine])
        for q in range(world):
            if q != rank and send_keys[q].numel():
                sel = owner == q
                slot[sel] = send_off[q] + torch.searchsorted(send_keys[q], flat[sel])
        sizes = [k.numel() for k in keys]
        # clone: the kernels read slot rows with 16-byte vector loads, split() views may start unaligned
        s_c, s_g, s_b = (t.clone() for t in torch.split(slot.to(torch.int32), sizes))
        rows = own // NG
        n_rows = self.row_hi - self.row_lo
        counts = torch.bincount(rows - self.row_lo, minlength=n_rows)
        indptr = torch.zeros(n_rows + 1, dtype=torch.int64, device=dev)
        indptr[1:] = torch.cumsum(counts, dim=0)
        # load vector: halo vertices travel as (global id) lists
        vo = self.vertex_owner
        touched = torch.zeros(mesh.num_vertices, dtype=torch.bool, device=dev)
        touched[dm_c.reshape(-1)] = True
        b_send_local = [torch.nonzero(touched & (vo == q)).reshape(-1) if q != rank else flat.new_zeros(0)
                        for q in range(world)]
        b_recv_global = _exchange([gv[v] for v in b_send_local], torch.int64, dev, self.group) \
            if world > 1 else [flat.new_zeros(0)]
        self.plan = SimpleNamespace(
            mesh=mesh, cell_tags8=cell_tags8, active=active.to(torch.int32).contiguous(),
            ghost=ghost.to(torch.int32).contiguous(), entities=ents.to(torch.int32).contiguous(),
            slots_cells=s_c.reshape(-1, nv * nv).contiguous(), slots_ghost=s_g.reshape(-1, (nv + 1) ** 2).contiguous(),
            slots_boundary=s_b.reshape(-1, nv * nv).contiguous(), nnz=nnz, total=total, n_rows=n_rows,
            indptr=indptr.to(torch.int64), indices=(own - rows * NG).contiguous(),
            send_ranges=[(send_off[q], send_off[q] + int(send_keys[q].numel())) for q in range(world)],
            recv_slots=[torch.searchsorted(own, recv_keys[p]) if p != rank else None for p in range(world)],
            b_send_local=b_send_local,
            b_recv_rows=[(b_recv_global[p] - self.row_lo) if p != rank else None for p in range(world)],
            owned_vertices=torch.nonzero(vo == rank).reshape(-1))
        self.data = torch.zeros(total, dtype=torch.float64, device=dev)
        self.b_local = torch.zeros(mesh.num_vertices, dtype=torch.float64, device=dev)
        self._recv_data = [torch.empty(0 if s is None else s.numel(), dtype=torch.float64, device=dev)
                           for s in self.plan.recv_slots]
        self._recv_b = [torch.empty(0 if s is None else s.numel(), dtype=torch.float64, device=dev)
                        for s in self.plan.b_recv_rows]
        return self.plan

    def _build_rows_plan(self, cell_tags8, facet_tags8):
        """Owner-computes plan: the single-GPU symbolic phase on the local mesh (slab + ghost layers),
        restricted to the rows this rank owns.  The owned rows are a contiguous range of local rows."""
        mesh = self.mesh
        if mesh.device.type == "cuda":
            ents = mesh_scripts._integration_entities_dev(mesh, cell_tags8, facet_tags8, 4, (1, 2))
        else:
            ents = self._entities_host(cell_tags8, facet_tags8)
        owned = self.vertex_owner == self.rank
        plan = assemble.AssemblyPlan(mesh, cell_tags8.contiguous(), facet_tags8.contiguous(),
                                     ents.reshape(-1, 2), method="rows", row_mask=owned)
        if plan.method != "rows":
            raise NotImplementedError("owner-computes sharding needs the row-gather plan")
        ov = torch.nonzero(owned).reshape(-1)
        lo, hi = int(ov[0]), int(ov[-1]) + 1
        assert hi - lo == ov.numel() == self.row_hi - self.row_lo, "owned rows are not contiguous"
        plan.owned_lo, plan.owned_hi = lo, hi
        plan.owned_vertices = ov
        ip = plan.indptr.long()
        plan.owned_slots = (int(ip[lo]), int(ip[hi]))      # CSR range of the owned rows (host ints: the
        plan.owned_indptr = (ip[lo:hi + 1] - ip[lo])       # numeric phase must not synchronise)
        plan.owned_cols = self.global_vertex[plan.indices[plan.owned_slots[0]:plan.owned_slots[1]].long()]
        plan.send_ranges = []
        self.plan = plan
        self.data, self.b_local = plan.new_outputs()
        return plan

    def owned_csr(self):
        """(indptr, global column ids, data view, b view) of the owned rows (mode="rows")."""
        p = self.plan
        a, b = p.owned_slots
        return p.owned_indptr, p.owned_cols, self.data[a:b], self.b_local[p.owned_lo:p.owned_hi]

    def _entities_host(self, cell_tags8, facet_tags8):
        return entities_host(self.mesh, cell_tags8, facet_tags8)

    # ---- numeric phase -----------------------------------------------------------------------------
    def assemble(self, sigma=1.0, marks=None, local_kernels=None):
        """Owned CSR values + owned load vector entries.  `local_kernels(plan, phi, f, sigma, data, b)`
        replaces the CUDA kernels in the CPU (gloo) tests only."""
        p = self.plan
        run = local_kernels or assemble.assemble_into
        if self.mode == "rows":
            if local_kernels is None:
                run(p, self.phi, self.f, sigma, self.data, self.b_local, marks=marks)
            else:
                run(p, self.phi, self.f, sigma, self.data, self.b_local)
            _, _, data, b = self.owned_csr()
            return data, b
        if local_kernels is None:
            run(p, self.phi, self.f, sigma, self.data, self.b_local, marks=marks)
        else:
            self.data.zero_()
            self.b_local.zero_()
            run(p, self.phi, self.f, sigma, self.data, self.b_local)
        self.exchange()
        return self.data[:p.nnz], self.b_owned()

    def enable_peer_halo(self):
        """mode "exchange": the halo contributions travel as stores into the owners' HBM over NVLink and are added there
        by the same kernel (phifem_b200/peer.py HaloExchange, csrc/peer.cu) instead of grouped NCCL send / recv + eager
        index_add_.  Call after build_plan; returns False where peer mapping is not available (every rank agrees)."""
        from . import peer
        if self.world == 1 or self.mode != "exchange" or self.plan is None or self.device.type != "cuda":
            return False
        p, world, rank = self.plan, self.world, self.rank
        try:
            halo_a = peer.HaloExchange(rank, world,
                                       [hi - lo for lo, hi in p.send_ranges],
                                       [0 if s is None else s.numel() for s in p.recv_slots],
                                       send_src=[lo for lo, _ in p.send_ranges], group=self.group)
            halo_b = peer.HaloExchange(rank, world, [t.numel() for t in p.b_send_local],
                                       [0 if s is None else s.numel() for s in p.b_recv_rows], group=self.group)
            ok = 1
        except Exception:   # noqa: BLE001 -- every rank must agree, see below
            halo_a = halo_b = None
            ok = 0
        t = torch.tensor([ok], device=self.device)
        dist.all_reduce(t, op=dist.ReduceOp.MIN, group=self.group)
        if int(t.item()) == 0:
            for h in (halo_a, halo_b):
                if h is not None:
                    h.close()
            return False
        empty = torch.zeros(0, dtype=torch.int64, device=self.device)
        self._halo = SimpleNamespace(
            a=halo_a, b=halo_b,
            a_recv=torch.cat([s for s in p.recv_slots if s is not None] or [empty]).contiguous(),
            b_send=torch.cat([t_ for t_ in p.b_send_local] or [empty]).contiguous(),
            b_recv=torch.cat([s for s in p.b_recv_rows if s is not None] or [empty]).contiguous())
        return True

    def exchange(self):
        """Halo rows -> owners: over NVLink peer memory when enable_peer_halo() succeeded, else a grouped send/recv of
        the send segments and an indexed add on arrival."""
        if self.world == 1:
            return
        halo = getattr(self, "_halo", None)
        if halo is not None:
            p = self.plan
            self._b_owned = self.b_local[p.owned_vertices].clone()
            halo.a.exchange(self.data, halo.a_recv, self.data)
            halo.b.exchange(self.b_local, halo.b_recv, self._b_owned, send_index=halo.b_send)
            return
        p, ops, bsend = self.plan, [], []
        for q in range(self.world):
            if q == self.rank:
                continue
            lo, hi = p.send_ranges[q]
            if hi > lo:
                ops.append(dist.P2POp(dist.isend, self.data[lo:hi], q, group=self.group))
            if self._recv_data[q].numel():
                ops.append(dist.P2POp(dist.irecv, self._recv_data[q], q, group=self.group))
            if p.b_send_local[q].numel():
                bsend.append(self.b_local[p.b_send_local[q]])
                ops.append(dist.P2POp(dist.isend, bsend[-1], q, group=self.group))
            if self._recv_b[q].numel():
                ops.append(dist.P2POp(dist.irecv, self._recv_b[q], q, group=self.group))
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
        self._b_owned = self.b_local[p.owned_vertices].clone()
        for q in range(self.world):
            if q == self.rank:
                continue
            if self._recv_data[q].numel():
                self.data.index_add_(0, p.recv_slots[q], self._recv_data[q])
            if self._recv_b[q].numel():
                self._b_owned.index_add_(0, p.b_recv_rows[q], self._recv_b[q])

    def b_owned(self):
        if self.world == 1:
            return self.b_local[self.plan.owned_vertices]
        return self._b_owned
