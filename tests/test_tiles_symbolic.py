"""CPU-only: the symbolic phase of the cell-once cell pass (phifem_b200/tiles.py).  The CUDA kernel
(csrc/assemble_tiles.cu) is emulated in numpy from the plan arrays -- tile by tile, chunk by chunk: the cell in every
slot is evaluated with the oracle's closed-form element tensor, every row then pulls its records of the chunk -- and
must reproduce the cell part of the oracle's operator (reference demo/strong-dirichlet/flower/main.py:105,107-112,
126-128): every (row, cell) contribution exactly once, at the right position, in ascending cell order."""
import numpy as np
import pytest
import torch

from oracle import assembly as OA
from oracle import tags as OT
from phifem_b200 import assemble, synthetic
from phifem_b200.mesh import MeshTags


def _problem(d, n, seed=11):
    m = synthetic.rectangle_mesh(n, device="cpu") if d == 2 else synthetic.box_mesh(n, device="cpu")
    m = synthetic.unstructured_variant(m, jitter=0.15, seed=seed)
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
    return m, x, cells, phi, f, out


def _emulate_cell_pass(tl, plan, x, cells, phi, f, ct, sigma):
    R, nv = tl.rows_per_tile, cells.shape[1]
    indptr, indices = plan.indptr.numpy().astype(np.int64), plan.indices.numpy().astype(np.int64)
    rows, dpos = tl.rows.numpy().astype(np.int64), tl.diag_pos.numpy().astype(np.int64)
    chunk_ptr, sv = tl.chunk_ptr.numpy(), tl.slot_verts.numpy().view(np.uint32).reshape(-1, 4)
    rec_base, rec_off = tl.rec_base.numpy(), tl.rec_off.numpy().view(np.uint16).reshape(-1, R + 1)
    rec = tl.rec.numpy().view(np.uint32)
    cell_of = {tuple(c): i for i, c in enumerate(cells)}
    data = np.full(plan.nnz, np.nan)
    b = np.zeros(len(x))
    seen, last_cell, evaluated = set(), {}, 0
    assert len(chunk_ptr) == tl.n_tiles + 1 and tl.n_tiles == (len(rows) + R - 1) // R
    for t in range(tl.n_tiles):
        mine = rows[t * R:(t + 1) * R]
        for r in mine:
            data[indptr[r]:indptr[r + 1]] = 0.0
        prev = -1
        for c in range(chunk_ptr[t], chunk_ptr[t + 1]):
            tensors = {}
            for lane in range(R):
                v = sv[c * R + lane]
                if v[0] == 0xFFFFFFFF:
                    continue
                verts = [int(v[0]), int(v[1] & 0x7FFFFFFF)] + [int(q) for q in v[2:nv]]
                k = cell_of[tuple(verts)]                     # cell-local order kept
                assert k > prev, "cells of a tile ascend"
                prev = k
                assert ct[k] in (1, 2) and bool(v[1] >> 31) == (ct[k] == 2)
                assert set(verts) & set(mine), "a slot holds a cell touching the tile"
                tensors[lane] = (k,) + OA.cell_tensors_closed_form(x, cells[k:k + 1], phi, f, np.array([ct[k] == 2]),
                                                                   sigma)
                evaluated += 1
            for l, r in enumerate(mine):
                for q in range(rec_base[c] + rec_off[c, l], rec_base[c] + rec_off[c, l + 1]):
                    w = int(rec[q])
                    k, At, bt = tensors[w & 0xFF]
                    i = (w >> 8) & 3
                    assert cells[k][i] == r and (r, k) not in seen
                    assert last_cell.get(r, -1) < k, "a row sums its cells in ascending order"
                    last_cell[r] = k
                    seen.add((r, k))
                    assert indices[indptr[r] + dpos[t * R + l]] == r
                    data[indptr[r] + dpos[t * R + l]] += At[0, i, i]
                    b[r] += bt[0, i]
                    for m in range(nv - 1):
                        j = m + (m >= i)
                        p = (w >> (10 + 7 * m)) & 0x7F
                        assert indices[indptr[r] + p] == cells[k][j]
                        data[indptr[r] + p] += At[0, i, j]
            assert rec_off[c, len(mine):].max() == rec_off[c, len(mine)], "no records beyond the tile's rows"
    assert evaluated == tl.n_cell_slots
    return data, b, seen


@pytest.mark.parametrize("d,n,R,order", [(2, 14, 128, "natural"), (2, 20, 256, "morton"), (3, 5, 128, "morton"),
                                         (3, 6, 256, "auto"), (3, 4, 128, "natural")])
def test_cell_tiles_reproduce_the_oracle_cell_operator(d, n, R, order):
    m, x, cells, phi, f, out = _problem(d, n)
    plan = assemble.build_plan(m, MeshTags(m, d, torch.from_numpy(out["cell_tags"])),
                               MeshTags(m, d - 1, torch.from_numpy(out["facet_tags"])), out["ds100"],
                               method="rows", order=order, cell_pass="tiles", rows_per_tile=R)
    rp = plan.rowsplan
    tl = rp.tiles
    assert rp.cell_pass == "tiles" and tl is not None and rp.cells.n_listed == 0
    assert rp.order == ("morton" if order == "auto" else order)
    nnz_row = np.diff(plan.indptr.numpy())
    assert sorted(tl.rows.numpy()) == list(np.nonzero(nnz_row > 0)[0])
    ct = out["cell_tags"]
    data, b, seen = _emulate_cell_pass(tl, plan, x, cells, phi, f, ct, 1.0)
    na = int(np.isin(ct, (1, 2)).sum())
    assert len(seen) == (d + 1) * na == tl.n_records == rp.n_cell_records
    assert 1.0 <= tl.recompute <= d + 1
    # the cell part of the oracle's operator: no boundary entities, no ghost facets
    ft0 = np.where(np.isin(out["facet_tags"], (2, 3)), 1, out["facet_tags"])
    ip, ix, want, wb = OA.assemble_strong_dirichlet(x, cells, cells, len(x), phi, f, ct, ft0, out["c2f"], out["f2c"],
                                                    np.zeros(0, dtype=np.int32), sigma=1.0)
    # same values on the (larger) pattern of the full operator
    import scipy.sparse as sp
    full = sp.csr_matrix((np.nan_to_num(data), plan.indices.numpy(), plan.indptr.numpy()), shape=(len(x), len(x)))
    cellop = sp.csr_matrix((want, ix, ip), shape=(len(x), len(x)))
    assert not np.isnan(data).any()
    assert abs(full - cellop).max() <= 1e-13 * np.abs(want).max()
    assert np.abs(b - wb).max() <= 1e-13 * np.abs(wb).max()


def test_cell_tiles_with_a_row_mask_list_only_owned_rows():
    m, x, cells, phi, f, out = _problem(3, 5, seed=3)
    mask = torch.zeros(m.num_vertices, dtype=torch.bool)
    mask[: m.num_vertices // 2] = True
    from phifem_b200.assemble import AssemblyPlan, _plan_inputs
    c8, f8, ents = _plan_inputs(m, MeshTags(m, 3, torch.from_numpy(out["cell_tags"])),
                                MeshTags(m, 2, torch.from_numpy(out["facet_tags"])), out["ds100"])
    plan = AssemblyPlan(m, c8, f8, ents, row_mask=mask, cell_pass="tiles", rows_per_tile=128)
    tl = plan.rowsplan.tiles
    nnz_row = np.diff(plan.indptr.numpy())
    assert sorted(tl.rows.numpy()) == list(np.nonzero((nnz_row > 0) & mask.numpy())[0])
    data, b, seen = _emulate_cell_pass(tl, plan, x, cells, phi, f, out["cell_tags"], 1.0)
    act = np.isin(out["cell_tags"], (1, 2))
    assert len(seen) == int(mask.numpy()[cells[act]].sum())


def _emulate_push_pass(tl, plan, x, cells, phi, f, ct, sigma):
    """k_assemble_push_p1 from the plan arrays: the cell of every slot is evaluated once and its rows are added to the
    accumulators of the tile's own rows through the push words."""
    R, nv = tl.rows_per_tile, cells.shape[1]
    indptr, indices = plan.indptr.numpy().astype(np.int64), plan.indices.numpy().astype(np.int64)
    rows, dpos = tl.rows.numpy().astype(np.int64), tl.diag_pos.numpy().astype(np.int64)
    chunk_ptr, sv = tl.chunk_ptr.numpy(), tl.slot_verts.numpy().view(np.uint32).reshape(-1, 4)
    push = tl.push.numpy().view(np.uint32).reshape(-1, 4)
    assert push.shape == sv.shape
    cell_of = {tuple(c): i for i, c in enumerate(cells)}
    data = np.full(plan.nnz, np.nan)
    b = np.zeros(len(x))
    seen = set()
    for t in range(tl.n_tiles):
        mine = rows[t * R:(t + 1) * R]
        for r in mine:
            data[indptr[r]:indptr[r + 1]] = 0.0
        for q in range(chunk_ptr[t] * R, chunk_ptr[t + 1] * R):
            v = sv[q]
            if v[0] == 0xFFFFFFFF:
                assert not push[q].any()
                continue
            verts = [int(v[0]), int(v[1] & 0x7FFFFFFF)] + [int(u) for u in v[2:nv]]
            k = cell_of[tuple(verts)]
            _, At, bt = (k,) + OA.cell_tensors_closed_form(x, cells[k:k + 1], phi, f, np.array([ct[k] == 2]), sigma)
            pushed = 0
            for i in range(nv):
                w = int(push[q, i])
                r = verts[i]
                if not (w >> 8) & 1:
                    assert w == 0 and r not in set(mine), "a row of the tile is always pushed"
                    continue
                l = w & 0xFF
                assert mine[l] == r and (r, k) not in seen
                seen.add((r, k))
                pushed += 1
                assert indices[indptr[r] + dpos[t * R + l]] == r
                data[indptr[r] + dpos[t * R + l]] += At[0, i, i]
                b[r] += bt[0, i]
                for m in range(nv - 1):
                    j = m + (m >= i)
                    p = (w >> (10 + 7 * m)) & 0x7F
                    assert indices[indptr[r] + p] == verts[j]
                    data[indptr[r] + p] += At[0, i, j]
            assert pushed >= 1, "a slot holds a cell touching the tile"
            assert not push[q, nv:].any()
    return data, b, seen


@pytest.mark.parametrize("d,n,R,order", [(2, 14, 128, "natural"), (2, 20, 256, "morton"), (3, 5, 128, "morton"),
                                         (3, 6, 256, "auto")])
def test_push_words_reproduce_the_oracle_cell_operator(d, n, R, order):
    m, x, cells, phi, f, out = _problem(d, n)
    plan = assemble.build_plan(m, MeshTags(m, d, torch.from_numpy(out["cell_tags"])),
                               MeshTags(m, d - 1, torch.from_numpy(out["facet_tags"])), out["ds100"],
                               method="rows", order=order, cell_pass="push", rows_per_tile=R)
    rp = plan.rowsplan
    tl = rp.tiles
    assert rp.cell_pass == "push" and tl.push is not None and rp.cells.n_listed == 0
    ct = out["cell_tags"]
    data, b, seen = _emulate_push_pass(tl, plan, x, cells, phi, f, ct, 1.0)
    na = int(np.isin(ct, (1, 2)).sum())
    assert len(seen) == (d + 1) * na
    ft0 = np.where(np.isin(out["facet_tags"], (2, 3)), 1, out["facet_tags"])
    ip, ix, want, wb = OA.assemble_strong_dirichlet(x, cells, cells, len(x), phi, f, ct, ft0, out["c2f"], out["f2c"],
                                                    np.zeros(0, dtype=np.int32), sigma=1.0)
    import scipy.sparse as sp
    assert not np.isnan(data).any()
    full = sp.csr_matrix((data, plan.indices.numpy(), plan.indptr.numpy()), shape=(len(x), len(x)))
    cellop = sp.csr_matrix((want, ix, ip), shape=(len(x), len(x)))
    assert abs(full - cellop).max() <= 1e-13 * np.abs(want).max()
    assert np.abs(b - wb).max() <= 1e-13 * np.abs(wb).max()


def test_push_words_with_a_row_mask_push_only_owned_rows():
    m, x, cells, phi, f, out = _problem(3, 5, seed=3)
    mask = torch.zeros(m.num_vertices, dtype=torch.bool)
    mask[: m.num_vertices // 2] = True
    from phifem_b200.assemble import AssemblyPlan, _plan_inputs
    c8, f8, ents = _plan_inputs(m, MeshTags(m, 3, torch.from_numpy(out["cell_tags"])),
                                MeshTags(m, 2, torch.from_numpy(out["facet_tags"])), out["ds100"])
    plan = AssemblyPlan(m, c8, f8, ents, row_mask=mask, cell_pass="push", rows_per_tile=128)
    data, b, seen = _emulate_push_pass(plan.rowsplan.tiles, plan, x, cells, phi, f, out["cell_tags"], 1.0)
    act = np.isin(out["cell_tags"], (1, 2))
    assert len(seen) == int(mask.numpy()[cells[act]].sum())
