"""Owner-computes sharding of an ARBITRARY simplex mesh across ranks.

`dist.SlabProblem` shards the synthetic weak-scaling box by construction.  This module shards any
triangle / tetrahedron mesh the way the north star describes: cells are sorted along a space-filling
(Morton) curve of their centroids and cut into `world` contiguous ranges; a CSR row (= vertex) is owned by
the lowest rank owning a cell that touches it; every rank keeps its own range, the cells touching its rows and
their facet neighbours (the reach of the ghost penalty) -- plus, with `single_layer_cut`, the vertex neighbours of
all of those --, classifies them redundantly and runs the row-gather kernels on its own rows.  Nothing is exchanged in the numeric phase except the 8-byte "any exterior cell" all-reduce of
the facet algebra (reference src/phifem/mesh_scripts.py:469-474).

The local mesh keeps the global relative order of cells, vertices and hence facets (order-preserving
relabelling), so every owned row sums the same contributions in the same order as on one GPU: the stacked
ranks are bitwise identical to the single-GPU operator.

The partitioner is symbolic-phase plumbing (torch ops on the mesh's device; every rank derives its part from
the same global arrays, no communication), so it also runs on CPU tensors under `gloo`, which is how the CPU
tests cover it.  The reference has no counterpart: dolfinx partitions with a graph partitioner behind
`create_rectangle` / `read_mesh` (demo/strong-dirichlet/flower/main.py:48-49) and phifem carries a TODO for
parallel tag transfer (src/phifem/mesh_scripts.py:264).
"""
import torch

from . import assemble, mesh_scripts
from .mesh import Mesh


def morton_keys(pts):
    """63-bit Morton keys of points [n, d] (21 bits per axis over the bounding box)."""
    lo, hi = pts.min(dim=0).values, pts.max(dim=0).values
    q = ((pts - lo) / (hi - lo).clamp(min=1e-300) * (2 ** 21 - 1)).long().clamp_(0, 2 ** 21 - 1)
    d = pts.shape[1]
    key = torch.zeros(pts.shape[0], dtype=torch.int64, device=pts.device)
    for bit in range(21):
        for k in range(d):
            key |= ((q[:, k] >> bit) & 1) << (bit * d + k)
    return key


def partition_cells(mesh, world, weights=None):
    """cell_owner [Nc] int64: contiguous ranges of the Morton order of the cell centroids, of equal total
    weight (default weight 1 per cell)."""
    cent = mesh.x[mesh.cells.long()].mean(dim=1)
    order = torch.argsort(morton_keys(cent), stable=True)
    w = torch.ones(mesh.num_cells, dtype=torch.float64, device=mesh.device) if weights is None \
        else weights.to(mesh.device, torch.float64)
    cum = torch.cumsum(w[order], 0)
    part = torch.clamp(((cum - 0.5 * w[order]) / cum[-1] * world).long(), 0, world - 1)
    owner = torch.empty(mesh.num_cells, dtype=torch.int64, device=mesh.device)
    owner[order] = part
    return owner


class PartitionedProblem:
    """One rank's share of (mesh, phi, f): local mesh = cells touching owned rows + their facet neighbours.

    Attributes: `mesh` (local), `phi`, `f` (local vertex values), `global_vertex` / `global_cell` (ids of the
    local entities, ascending), `row_mask` (local vertices whose rows this rank owns), `cell_owned`."""

    def __init__(self, mesh, phi, f, rank, world, group=None, weights=None, single_layer_cut=False):
        if mesh.cell_type not in ("triangle", "tetrahedron"):
            raise NotImplementedError("sharding supports triangles and tetrahedra")
        self.rank, self.world, self.group = rank, world, group
        dev = mesh.device
        cells = mesh.cells.long()
        cell_owner = partition_cells(mesh, world, weights)
        vowner = torch.full((mesh.num_vertices,), world, dtype=torch.int64, device=dev)
        vowner.scatter_reduce_(0, cells.reshape(-1), cell_owner.repeat_interleave(cells.shape[1]), reduce="amin")
        self.vertex_owner_global = vowner
        owned_v = vowner == rank
        keep = owned_v[cells].any(dim=1)                         # cells touching an owned row
        # + their facet neighbours (ghost-penalty macro elements, tags on both sides of every relevant facet)
        fac = mesh.c2f[keep].long().reshape(-1)
        nb = mesh.f2c[fac].long().reshape(-1)
        keep = keep.clone()
        keep[nb[nb >= 0]] = True
        keep |= cell_owner == rank     # plus the rank's own range, so that every cell's tag has one home
        self.single_layer_cut = bool(single_layer_cut)
        if single_layer_cut:
            # `single_layer_cut` (reference :349-358) re-tags a cut cell from the tags of every cell sharing a VERTEX
            # with it: one more layer, so that the cells above see all their vertex neighbours
            vmark = torch.zeros(mesh.num_vertices, dtype=torch.bool, device=dev)
            vmark[cells[keep].reshape(-1)] = True
            keep = keep | vmark[cells].any(dim=1)
        gc = torch.nonzero(keep).reshape(-1)                      # ascending global cell ids
        gv = torch.unique(cells[gc].reshape(-1))                  # ascending global vertex ids
        relabel = torch.full((mesh.num_vertices,), -1, dtype=torch.int64, device=dev)
        relabel[gv] = torch.arange(gv.numel(), device=dev)
        self.mesh = Mesh(mesh.x[gv], relabel[cells[gc]].to(torch.int32), mesh.cell_type, dev)
        self.global_cell, self.global_vertex = gc, gv
        self.n_global_vertices = mesh.num_vertices
        self.row_mask = owned_v[gv]
        self.cell_owned = cell_owner[gc] == rank
        self.n_owned_cells = int(self.cell_owned.sum())
        self.phi = phi.to(dev)[gv].contiguous()
        self.f = f.to(dev)[gv].contiguous()
        self.plan = None

    # ---- tags --------------------------------------------------------------------------------------
    def classify(self, dls, ws, mark=None):
        """Cells, interior facets overlapped with the 8-byte all-reduce of the exterior-cell count, mesh-boundary
        facets (all on the current stream; mesh_scripts.classify_sharded)."""
        return mesh_scripts.classify_sharded(self.mesh, dls, ws, group=self.group, world=self.world, mark=mark,
                                             single_layer_cut=self.single_layer_cut)

    # ---- symbolic phase ----------------------------------------------------------------------------
    def build_plan(self, cell_tags8, facet_tags8, entities=None):
        mesh = self.mesh
        if entities is None:
            entities = mesh_scripts._integration_entities_dev(mesh, cell_tags8, facet_tags8, 4, (1, 2))
        plan = assemble.AssemblyPlan(mesh, cell_tags8.contiguous(), facet_tags8.contiguous(),
                                     entities.reshape(-1, 2), method="rows", row_mask=self.row_mask)
        if plan.method != "rows":
            raise NotImplementedError("owner-computes sharding needs the row-gather plan")
        ip = plan.indptr.long()
        rows = torch.nonzero(self.row_mask).reshape(-1)
        cnt = ip[rows + 1] - ip[rows]
        plan.owned_rows_local = rows
        plan.owned_rows_global = self.global_vertex[rows]
        plan.owned_indptr = torch.zeros(rows.numel() + 1, dtype=torch.int64, device=mesh.device)
        plan.owned_indptr[1:] = torch.cumsum(cnt, 0)
        # CSR slots of the owned rows, row by row
        start = torch.repeat_interleave(ip[rows] - plan.owned_indptr[:-1], cnt)
        plan.owned_slots = start + torch.arange(int(plan.owned_indptr[-1]), device=mesh.device)
        plan.owned_cols = self.global_vertex[plan.indices[plan.owned_slots].long()]
        self.plan = plan
        self.data, self.b_local = plan.new_outputs()
        return plan

    # ---- numeric phase -----------------------------------------------------------------------------
    def assemble(self, sigma=1.0, marks=None, local_kernels=None):
        """Runs the row-gather kernels on the owned rows.  `local_kernels(plan, phi, f, sigma, data, b)`
        replaces the CUDA kernels in the CPU (gloo) tests only."""
        p = self.plan
        if local_kernels is None:
            assemble.assemble_into(p, self.phi, self.f, sigma, self.data, self.b_local, marks=marks)
        else:
            local_kernels(p, self.phi, self.f, sigma, self.data, self.b_local)
        return self.data, self.b_local

    def owned_csr(self):
        """(global row ids [m] ascending, indptr [m+1], global column ids, data, b) of the owned rows."""
        p = self.plan
        return (p.owned_rows_global, p.owned_indptr, p.owned_cols, self.data[p.owned_slots],
                self.b_local[p.owned_rows_local])
