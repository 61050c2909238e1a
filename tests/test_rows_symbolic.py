"""CPU-only: the symbolic phase of the row-gather assembly (phifem_b200/rows.py).  The CUDA kernel is
emulated in numpy from the plan arrays: every record is decoded through the row's own column list back
to vertex ids, the entity it names is looked up in the mesh, and the matching row of the oracle's
element tensor is added at the recorded positions.  The result must be the oracle's assembled CSR
operator -- i.e. every (row, entity) contribution is listed exactly once with the right positions."""
import numpy as np
import pytest
import torch

from oracle import assembly as OA
from oracle import tags as OT
from phifem_b200 import assemble, synthetic
from phifem_b200.mesh import MeshTags

PAD = 0xFFFFFFFF


def _records(rl, indptr, indices):
    """Yield (row, position of its diagonal, record words) in the order the kernel visits them."""
    ptr = rl.ptr.numpy().astype(np.int64)
    rec = rl.rec.numpy().view(np.uint32).reshape(-1, 32, rl.words)
    rows = rl.rows.numpy().astype(np.int64)
    dpos = rl.diag_pos.numpy().astype(np.int64)
    assert len(set(rows)) == len(rows) == rl.n_listed and len(ptr) == rl.n_slices + 1
    for li, r in enumerate(rows):
        assert indices[indptr[r] + dpos[li]] == r
        s, lane = li >> 5, li & 31
        for k in range(ptr[s], ptr[s + 1]):
            w = rec[k, lane]
            if w[-1] == PAD:
                continue
            yield r, dpos[li], w


def _emulate(rp, m, x, cells, phi, f, out, sigma):
    d = x.shape[1]
    nv = d + 1
    plan = rp.plan
    indptr, indices = plan.indptr.numpy().astype(np.int64), plan.indices.numpy().astype(np.int64)
    data = np.full(plan.nnz, np.nan)
    b = np.zeros(len(x))
    for r in rp.cells.rows.numpy():          # the cell pass writes every row of its list
        data[indptr[r]:indptr[r + 1]] = 0.0
    cell_of = {tuple(sorted(c)): i for i, c in enumerate(cells)}
    ct = out["cell_tags"]

    # cells
    seen = set()
    for r, dp, w in _records(rp.cells, indptr, indices):
        pos = [(int(w[0]) >> (8 * j)) & 0xFF for j in range(d)]
        others = [indices[indptr[r] + p] for p in pos]
        c = cell_of[tuple(sorted([r] + others))]
        assert ct[c] in (1, 2) and bool((int(w[0]) >> 24) & 1) == (ct[c] == 2)
        assert (r, c) not in seen
        seen.add((r, c))
        At, bt = OA.cell_tensors_closed_form(x, cells[c:c + 1], phi, f, np.array([ct[c] == 2]), sigma)
        i = list(cells[c]).index(r)
        if rp.cell_geom is not None:
            # geometry mode: word 0 also names the row's cell-local index, word 1 the cell's entry of the geometry
            # table; the other vertices come in ascending cell-local order
            assert (int(w[0]) >> 25) & 3 == i and int(plan.active[int(w[1])]) == c
            assert others == [int(v) for k, v in enumerate(cells[c]) if k != i]
        else:
            assert len(w) == 1
        data[indptr[r] + dp] += At[0, i, i]
        for p, v in zip(pos, others):
            data[indptr[r] + p] += At[0, i, list(cells[c]).index(v)]
        b[r] += bt[0, i]
    assert len(seen) == nv * int(np.isin(ct, (1, 2)).sum())
    if rp.cell_geom is not None:
        # the table: S_ab = |K| grad(lambda_a).grad(lambda_b) for a < b, |K|, h_T^2
        act = plan.active.numpy()
        G, vol, h = OA.simplex_geometry(x, cells[act])
        tab = rp.cell_geom.numpy()
        k = 0
        for a in range(nv):
            for c2 in range(a + 1, nv):
                want = vol * (G[:, a] * G[:, c2]).sum(axis=1)
                assert np.abs(tab[:, k] - want).max() <= 1e-13 * np.abs(want).max()
                k += 1
        assert np.allclose(tab[:, k], vol, rtol=1e-14, atol=0) and np.allclose(tab[:, k + 1], h * h, rtol=1e-14, atol=0)
        assert np.all(tab[:, k + 2:] == 0.0)

    # ghost-penalty facets
    ghost = plan.ghost.numpy()
    if len(ghost):
        G8, macro = OA.ghost_tensors_closed_form(x, cells, phi, out["c2f"], out["f2c"], ghost, sigma)
        by_set = {frozenset(int(v) for v in macro[e]): e for e in range(len(ghost))}
        assert len(by_set) == len(ghost)
    seen = set()
    gm = rp.ghost_macro.numpy()
    ents = plan.entities.numpy()
    if len(ents):
        Bt = OA.boundary_tensors_closed_form(x, cells, phi, ents)
        em = rp.entity_macro.numpy()
    seen_b = set()
    for r, dp, w in _records(rp.surface, indptr, indices):
        if int(w[1]) >> 31:          # one-sided entity e, the row's vertex = facet vertex t
            e, t = (int(w[1]) & 0x0FFFFFFF) - len(ghost), (int(w[1]) >> 28) & 7
            pos = [(int(w[0]) >> (8 * j)) & 0xFF for j in range(d)]
            others = [int(indices[indptr[r] + p]) for p in pos]
            c, o = int(ents[e, 0]), int(ents[e, 1])
            fac = [int(v) for k, v in enumerate(cells[c]) if k != o]         # facet vertices, ascending local order
            assert [int(v) for v in em[e]] == fac + [int(cells[c][o])]
            assert fac[t] == r and others == [int(cells[c][o])] + [fac[(t + u) % d] for u in range(1, d)]
            assert (r, e) not in seen_b
            seen_b.add((r, e))
            i = list(cells[c]).index(r)
            data[indptr[r] + dp] += Bt[e, i, i]
            for p, v in zip(pos, others):
                data[indptr[r] + p] += Bt[e, i, list(cells[c]).index(v)]
            continue
        e, a = int(w[1]) & 0x0FFFFFFF, int(w[1]) >> 28       # ghost facet index, macro index of the row's vertex
        pos = [(int(w[0]) >> (8 * j)) & 0xFF for j in range(nv)]
        others = [int(indices[indptr[r] + p]) for p in pos]
        assert e == by_set[frozenset([int(r)] + others)]
        assert (r, e) not in seen
        seen.add((r, e))
        cplus, cminus = out["f2c"][ghost[e]]
        fv = set(cells[cplus]) & set(cells[cminus])
        # macro order: facet vertices as ordered in cell +, opposite vertex of cell +, of cell -
        assert int(gm[e][a]) == r and [int(v) for k, v in enumerate(gm[e]) if k != a] == others
        assert set(int(v) for v in gm[e][:d]) == fv
        assert int(gm[e][d]) in cells[cplus] and int(gm[e][d + 1]) in cells[cminus]
        assert [v for v in cells[cplus] if v in fv] == [int(v) for v in gm[e][:d]]
        mv = [int(v) for v in macro[e]]
        rsel = [k for k, v in enumerate(mv) if v == r]
        for col, p in zip([int(r)] + others, [dp] + pos):
            csel = [k for k, v in enumerate(mv) if v == col]
            data[indptr[r] + p] += G8[e][np.ix_(rsel, csel)].sum()
    assert len(seen) == (d + 2) * len(ghost)
    assert len(seen_b) == d * len(ents)
    assert rp.n_ghost_records == len(seen) and rp.n_entity_records == len(seen_b)
    return data, b


@pytest.mark.parametrize("d,n,order,geometry", [(2, 12, "natural", True), (2, 9, "morton", False), (3, 5, "natural", True),
                                                (3, 4, "morton", True), (3, 4, "natural", False)])
def test_rows_plan_reproduces_the_oracle_operator(d, n, order, geometry):
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
                               method="rows", order=order, geometry=geometry)
    rp = plan.rowsplan
    assert plan.method == "rows" and rp.order == order and (rp.cell_geom is not None) == geometry
    assert rp.cells.words == (2 if geometry else 1)
    assert plan.ghost.numel() > 0 and plan.entities.shape[0] > 0
    # listed rows = rows with pattern entries, each once; slices cover them
    rows = rp.cells.rows.numpy()
    nnz_row = np.diff(plan.indptr.numpy())
    assert sorted(rows) == list(np.nonzero(nnz_row > 0)[0])
    assert rp.max_row_nnz == nnz_row.max()
    assert 0 < rp.surface.n_listed <= rp.cells.n_listed
    if order == "natural":
        assert np.all(np.diff(rows) > 0)
    cnt = np.diff(rp.surface.ptr.numpy())   # surface list: rows sorted by record count, descending (one chunk here)
    assert np.all(np.diff(cnt) <= 0) and rp.surface.padding() < 0.5
    data, b = _emulate(rp, m, x, cells, phi, f, out, 1.0)
    ip, ix, want, wb = OA.assemble_strong_dirichlet(x, cells, cells, len(x), phi, f, out["cell_tags"],
                                                    out["facet_tags"], out["c2f"], out["f2c"],
                                                    out["ds100"], sigma=1.0)
    assert np.array_equal(plan.indptr.numpy(), ip) and np.array_equal(plan.indices.numpy(), ix)
    assert not np.isnan(data).any()                         # every CSR entry belongs to a listed row
    assert np.abs(data - want).max() <= 1e-13 * np.abs(want).max()
    assert np.abs(b - wb).max() <= 1e-13 * np.abs(wb).max()
    assert 0.0 <= rp.cells.padding() < 0.9


def test_rows_plan_falls_back_when_a_row_is_too_long(monkeypatch):
    from phifem_b200 import rows as rows_mod
    monkeypatch.setattr(rows_mod, "MAX_ROW_NNZ", 8)
    m = synthetic.box_mesh(3, device="cpu")
    tags = torch.ones(m.num_cells, dtype=torch.int32)
    ft = torch.ones(m.num_facets, dtype=torch.int32)
    plan = assemble.build_plan(m, MeshTags(m, 3, tags), MeshTags(m, 2, ft), None, method="rows")
    assert plan.method == "atomic" and plan.rowsplan is None
