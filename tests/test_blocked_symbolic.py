"""CPU-only: the symbolic phase of the owner-computes assembly (phifem_b200/blocked.py).  The CUDA kernel
is emulated in numpy from the plan arrays (buffer scatter + segment sums, with the oracle's element
tensors as the per-entity values) and must reproduce the oracle's assembled CSR operator."""
import numpy as np
import pytest
import torch

from oracle import assembly as OA
from oracle import tags as OT
from phifem_b200 import assemble, synthetic
from phifem_b200.mesh import MeshTags


def _emulate(bp, cell_t, cell_b, ghost_t, bnd_t, nnz, n):
    """cell_t[e] etc. are the element tensors of plan.active[e] / plan.ghost[e] / plan.entities[e]."""
    data = np.full(nnz, np.nan)
    b = np.zeros(n)
    desc = bp.block_desc.numpy()
    seg_start, seg_dest = bp.seg_start.numpy(), bp.seg_dest.numpy()
    cpos, gpos, bpos = bp.cell_pos.numpy(), bp.ghost_pos.numpy(), bp.bnd_pos.numpy()
    for blk in range(bp.n_blocks):
        ncon, s0, s1, c0, c1, g0, g1, b0, b1 = desc[blk, :9]
        buf = np.full(ncon, np.nan)

        def put(pos, i, vals):
            for e, v in enumerate(vals):
                p = pos[e // 2, i, e % 2]
                if p >= 0:
                    assert np.isnan(buf[p]), "two contributions share a buffer position"
                    buf[p] = v
        for i in range(c0, c1):
            put(cpos, i, np.concatenate([cell_t[i].ravel(), cell_b[i]]))
        for i in range(g0, g1):
            put(gpos, i, ghost_t[i].ravel())
        for i in range(b0, b1):
            put(bpos, i, bnd_t[i].ravel())
        assert not np.isnan(buf).any(), "unfilled buffer position"
        assert s0 % 8 == 0 and s1 % 8 == 0 and s1 <= bp.max_segments
        n_real = 0
        for s in range(s0, s0 + s1 - 1):       # s1 = padded count; pads are empty segments
            lo, hi = seg_start[s], seg_start[s + 1]
            if hi <= lo:
                assert lo == ncon
                continue
            n_real += 1
            dst = int(seg_dest[s])
            if dst >= 0:
                assert np.isnan(data[dst])
                data[dst] = buf[lo:hi].sum()
            else:
                b[dst & 0x7FFFFFFF] = buf[lo:hi].sum()
    return data, b


@pytest.mark.parametrize("d,n,cap", [(2, 12, 300), (3, 5, 2000), (3, 6, 27000)])
def test_blocked_plan_reproduces_the_oracle_operator(d, n, cap):
    m = synthetic.rectangle_mesh(n, device="cpu") if d == 2 else synthetic.box_mesh(n, device="cpu")
    m = synthetic.unstructured_variant(m, jitter=0.15, seed=11)
    x, cells = m.x.numpy(), m.cells.numpy().astype(np.int64)
    center = np.array([0.1, 0.05]) if d == 2 else np.array([0.52, 0.49, 0.51])
    r = 0.6 if d == 2 else 0.33
    phi = ((x - center) ** 2).sum(axis=1) - r * r
    f = np.random.default_rng(5).uniform(-1, 1, len(x))
    ct = m.cell_type
    pts = OT.cell_detection_points(ct, 1)
    ftab = np.asarray([OT.coordinate_basis(ct, p)[0] for p in OT.facet_points_in_cell(ct, 1)])
    out = OT.compute_tags_measures(x, cells, ct, phi[cells], OT.point_values_function(phi, cells, ftab),
                                   box_mode=True, detection_points=pts)
    plan = assemble.build_plan(m, MeshTags(m, d, torch.from_numpy(out["cell_tags"])),
                               MeshTags(m, d - 1, torch.from_numpy(out["facet_tags"])), out["ds100"],
                               method="blocked", capacity=cap)
    bp = plan.blocked
    assert plan.method == "blocked" and bp.capacity <= cap
    if cap < 3000:
        assert bp.n_blocks > 3
    # every active row belongs to exactly one block; contributions per block within the capacity
    assert int(bp.block_desc[:, 0].max()) == bp.capacity
    # element tensors of every instance from the oracle's closed forms
    cv = bp.cell_verts.numpy().astype(np.int64)
    cut = cv[:, 0] < 0
    cv[:, 0] &= 0x7FFFFFFF
    At, bt = OA.cell_tensors_closed_form(x, cv[:, :d + 1], phi, f, cut, 1.0)
    G8, macro = OA.ghost_tensors_closed_form(x, cells, phi, out["c2f"], out["f2c"], bp.ghost_facet.numpy(), 1.0)
    # fold the (2nv)^2 macro tensors onto the nv+1 distinct vertices, in the kernel's vertex order
    from phifem_b200.assemble import ghost_macro_vertices
    mv = ghost_macro_vertices(m, bp.ghost_facet).numpy()
    Gt = np.zeros((len(mv), d + 2, d + 2))
    for e in range(len(mv)):
        t = [int(np.nonzero(mv[e] == v)[0][0]) for v in macro[e]]
        np.add.at(Gt[e], (np.repeat(t, len(t)), np.tile(t, len(t))), G8[e].ravel())
    Bt = OA.boundary_tensors_closed_form(x, cells, phi, bp.bnd_entity.numpy())
    data, b = _emulate(bp, At, bt, Gt, Bt, plan.nnz, len(x))
    ip, ix, want, wb = OA.assemble_strong_dirichlet(x, cells, cells, len(x), phi, f, out["cell_tags"],
                                                    out["facet_tags"], out["c2f"], out["f2c"],
                                                    out["ds100"], sigma=1.0)
    assert not np.isnan(data).any()                         # every CSR entry written exactly once
    assert np.abs(data - want).max() <= 1e-13 * np.abs(want).max()
    assert np.abs(b - wb).max() <= 1e-13 * np.abs(wb).max()
    assert 1.0 <= bp.redundancy < 4.0


def test_blocked_plan_falls_back_when_a_row_is_too_dense():
    m = synthetic.box_mesh(3, device="cpu")
    tags = torch.ones(m.num_cells, dtype=torch.int32)
    ft = torch.ones(m.num_facets, dtype=torch.int32)
    plan = assemble.build_plan(m, MeshTags(m, 3, tags), MeshTags(m, 2, ft), None, method="blocked",
                               capacity=50)
    assert plan.method == "atomic" and plan.blocked is None
