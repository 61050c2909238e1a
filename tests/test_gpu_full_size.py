"""Parity at BASELINE.json's full size (config E: 6 * 204^3 = 50 937 984 tetrahedra, sphere level set) through
size-independent properties, checked with plain torch ops on the device (the oracle cannot run this size):

  * cell tags: for a non-degenerate P1 level set, num == +-den exactly iff the vertex values share a sign
    (mesh_scripts.py:124-128,343-347), so tag = 3 / 1 / 2 for all-positive / all-negative / mixed;
  * facet tags: the decision table of `_tag_facets` (:454-496) restated with torch ops;
  * known counts of this configuration (probed);
  * operator: run-to-run bitwise identical; with phi == 1 the form is int grad w.grad v + jumps: constants in
    the kernel, x^T A y == y^T A x, u^T A u = int |grad u|^2 for linear u, sum(b) = |Omega_h|.
"""
import warnings

import numpy as np
import pytest
import torch

from phifem_b200 import assemble, fem, mesh_scripts, synthetic

pytestmark = pytest.mark.gpu
N = 204


@pytest.fixture(scope="module")
def problem():
    mesh = synthetic.box_mesh(N, device="cuda")
    phi = synthetic.sphere_levelset(mesh.x)
    assert float(phi.abs().min()) > 1e-9            # non-degenerate: no vertex value near zero
    fn = fem.Function(fem.functionspace_p1_device(mesh), phi)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", RuntimeWarning)
        ctags, ftags, _, ds, _ = mesh_scripts.compute_tags_measures(mesh, fn, 1, box_mode=True)
    return mesh, phi, ctags, ftags, ds


def test_full_size_cell_tags(problem):
    mesh, phi, ctags, _, _ = problem
    assert mesh.num_cells == 50937984
    pos = (phi > 0)[mesh.cells.long()]
    want = torch.where(pos.all(dim=1), 3, torch.where((~pos).all(dim=1), 1, 2)).to(torch.int32)
    assert torch.equal(ctags.values_dev, want)
    hist = torch.bincount(ctags.values_dev, minlength=4).tolist()
    assert hist == [0, 19081700, 725850, 31130434]   # SURVEY.md section 8(d), exact counts at n = 204


def test_full_size_facet_tags_and_entities(problem):
    mesh, _, ctags, ftags, ds = problem
    assert mesh.num_facets == 102125664
    t = ctags.values_dev
    a = t[mesh.f2c[:, 0].long()]
    interior = mesh.f2c[:, 1] >= 0
    b = torch.where(interior, t[mesh.f2c[:, 1].clamp(min=0).long()], torch.zeros_like(a))
    lo, hi = torch.minimum(a, b), torch.maximum(a, b)
    # interior facets, cell tags {lo, hi}: {1,1}->1 {1,2}->3 {2,2}->2 {2,3}->4 {3,3}->5 {1,3}->6
    table = torch.tensor([[0, 0, 0, 0], [0, 1, 3, 6], [0, 0, 2, 4], [0, 0, 0, 5]], dtype=torch.int32, device="cuda")
    want = table[lo.long(), hi.long()]
    # the sphere stays inside the box: every mesh-boundary facet belongs to an exterior cell -> tag 5
    assert bool((a[~interior] == 3).all())
    want = torch.where(interior, want, torch.full_like(want, 5))
    assert torch.equal(ftags.values_dev, want)
    assert int((want == 6).sum()) == 0               # P1 on a conforming mesh: no direct in/out interface
    # ds(100): every tag-4 facet once, seen from its cell tagged 1 or 2 (mesh_scripts.py:619-622)
    ents = ds(100).integration_entities_dev.reshape(-1, 2).long()
    f4 = torch.nonzero(want == 4).reshape(-1)
    assert ents.shape[0] == f4.numel()
    facets_of_ents = mesh.c2f[ents[:, 0], ents[:, 1]].long()
    assert torch.equal(torch.sort(facets_of_ents).values, f4)
    assert bool(((t[ents[:, 0]] == 1) | (t[ents[:, 0]] == 2)).all())


def _spmv(plan, data, x):
    rows = torch.repeat_interleave(torch.arange(plan.n_rows, device="cuda"), (plan.indptr[1:] - plan.indptr[:-1]).long())
    y = torch.zeros(plan.n_rows, dtype=torch.float64, device="cuda")
    y.index_add_(0, rows, data * x[plan.indices.long()])
    return y


def test_full_size_operator_properties(problem):
    mesh, phi, ctags, ftags, ds = problem
    f = synthetic.ball_source(mesh.x)
    plan = assemble.build_plan(mesh, ctags, ftags, ds(100))
    assert plan.method == "rows" and plan.nnz == 51669474
    A1, b1 = assemble.assemble_strong_dirichlet(plan, phi, f, stab_coef=1.0)
    A2, b2 = assemble.assemble_strong_dirichlet(plan, phi, f, stab_coef=1.0)
    assert torch.equal(A1.data, A2.data) and torch.equal(b1, b2)          # no atomics: bitwise reproducible
    assert bool(torch.isfinite(A1.data).all()) and float(A1.data.abs().max()) > 0
    del A1, A2
    plan0 = assemble.build_plan(mesh, ctags, ftags, None)                  # without the one-sided term
    one = torch.ones(mesh.num_vertices, dtype=torch.float64, device="cuda")
    A, b = assemble.assemble_strong_dirichlet(plan0, one, one, stab_coef=1.0)
    scale = float(A.data.abs().max())
    assert float(_spmv(plan0, A.data, one).abs().max()) <= 1e-10 * scale   # constants in the kernel
    g = torch.Generator(device="cuda").manual_seed(5)
    xv = torch.rand(mesh.num_vertices, dtype=torch.float64, device="cuda", generator=g)
    yv = torch.rand(mesh.num_vertices, dtype=torch.float64, device="cuda", generator=g)
    xay, yax = float(xv @ _spmv(plan0, A.data, yv)), float(yv @ _spmv(plan0, A.data, xv))
    assert abs(xay - yax) <= 1e-11 * abs(xay)                              # symmetric
    t = ctags.values_dev
    vol = float(((t == 1) | (t == 2)).sum()) * (1.0 / N) ** 3 / 6.0
    assert abs(float(b.sum()) - vol) <= 1e-10 * vol                        # sum b = int phi f = |Omega_h|
    u = mesh.x[:, 0] + 2.0 * mesh.x[:, 1] - mesh.x[:, 2]                   # linear: |grad u|^2 = 6, no jumps
    energy = float(u @ _spmv(plan0, A.data, u))
    assert abs(energy - 6.0 * vol) <= 1e-9 * 6.0 * vol
    assert abs(vol - 4.0 / 3.0 * np.pi * 0.45 ** 3) < 0.02                 # slightly above the ball volume


def test_full_size_submesh_route(problem):
    """`box_mode=False` at the benchmark size (reference src/phifem/mesh_scripts.py:635-645): the submesh of Omega_h
    (19.8 M cells), tags transferred onto it, and the operator assembled on it with the plain `ds` measure -- the
    submesh keeps the relative order of cells and vertices, so its CSR values are the box-mode values, entry by entry."""
    mesh, phi, ctags, ftags, ds = problem
    fn = fem.Function(fem.functionspace_p1_device(mesh), phi)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", RuntimeWarning)
        sct, sft, sub, ds_sub, maps = mesh_scripts.compute_tags_measures(mesh, fn, 1, box_mode=False)
    assert sub.num_cells == 19081700 + 725850
    cmap = torch.as_tensor(maps[0].astype(np.int64), device="cuda")
    vmap = torch.as_tensor(maps[1].astype(np.int64), device="cuda")
    assert torch.equal(sct.values_dev, ctags.values_dev[cmap])                    # _transfer_tags, cells (:244)
    assert torch.equal(sft.values_dev[sub.c2f.long()], ftags.values_dev[mesh.c2f[cmap].long()])   # ... facets (:244-260)
    assert torch.equal(sub.x, mesh.x[vmap])
    f = synthetic.ball_source(mesh.x)
    box = assemble.build_plan(mesh, ctags, ftags, ds(100))
    A, b = assemble.assemble_strong_dirichlet(box, phi, f, stab_coef=1.0)
    splan = assemble.build_plan(sub, sct, sft, ds_sub)
    As, bs = assemble.assemble_strong_dirichlet(splan, phi[vmap], f[vmap], stab_coef=1.0)
    assert splan.nnz == box.nnz and splan.n_rows == vmap.numel()
    rows_nnz = (box.indptr[1:] - box.indptr[:-1])
    assert torch.equal(rows_nnz[vmap], splan.indptr[1:] - splan.indptr[:-1])       # rows outside Omega_h are empty
    assert int(rows_nnz.sum()) == int(rows_nnz[vmap].sum())
    assert torch.equal(vmap[As.indices.long()].to(torch.int32), A.indices)
    scale = float(A.data.abs().max())
    assert float((As.data - A.data).abs().max()) <= 1e-12 * scale
    assert float((bs - b[vmap]).abs().max()) <= 1e-12 * float(b.abs().max())
