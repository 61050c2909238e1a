"""GPU parity: CUDA assembly of the strong-Dirichlet operator (through the C ABI) vs the CPU oracle.

Bar (BASELINE.json north_star): identical CSR sparsity, entries within 1e-12 relative in fp64.
"Relative" is measured against the scale of the row (max |entry| of the row): many entries are
exact cancellations, so a per-entry relative error is meaningless."""
import warnings

import numpy as np
import pytest
import torch

from oracle import assembly as OA
from oracle import tags as OT
from phifem_b200 import assemble, fem, mesh_scripts, synthetic

pytestmark = pytest.mark.gpu
RTOL = 1e-12


def _setup(kind, n):
    if kind.startswith("tri"):
        mesh = synthetic.rectangle_mesh(n, device="cuda")
        center, radius = (0.013, -0.021), 0.61
    else:
        mesh = synthetic.box_mesh(n, device="cuda")
        center, radius = synthetic.SPHERE_CENTER, 0.37
    if kind.endswith("unstructured"):
        mesh = synthetic.unstructured_variant(mesh, jitter=0.2, seed=7)
    phi = synthetic.sphere_levelset(mesh.x, center=center, radius=radius)
    f = torch.from_numpy(np.random.default_rng(1234).uniform(-1, 1, mesh.num_vertices)).cuda()
    return mesh, phi, f


def _row_scale(indptr, data):
    scale = np.zeros(len(indptr) - 1)
    rows = np.repeat(np.arange(len(indptr) - 1), np.diff(indptr))
    np.maximum.at(scale, rows, np.abs(data))
    return scale, rows


@pytest.mark.parametrize("method,capacity", [("rows", None), ("rows", "morton"), ("rows", "geometry"),
                                             ("rows", "tiles128"), ("rows", "tiles256"),
                                             ("rows", "push128"), ("rows", "push256"),
                                             ("blocked", None), ("blocked", 2500), ("atomic", None)])
@pytest.mark.parametrize("kind,n", [("tri", 40), ("tri-unstructured", 24), ("tet", 10),
                                    ("tet-unstructured", 8)])
def test_cuda_operator_matches_oracle(kind, n, method, capacity):
    mesh, phi, f = _setup(kind, n)
    fn = fem.Function(fem.functionspace(mesh, 1), phi.cpu().numpy())
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", RuntimeWarning)
        ctags, ftags, _, ds, _ = mesh_scripts.compute_tags_measures(mesh, fn, 1, box_mode=True)
    order, geometry, cell_pass, rpt = "natural", False, "rows", 256
    if method == "rows" and str(capacity).startswith("tiles"):   # cell-once cell pass (csrc/assemble_tiles.cu)
        order, cell_pass, rpt, capacity = "morton", "tiles", int(capacity[5:]), None
    elif method == "rows" and str(capacity).startswith("push"):   # cell-once pass, push form (shared-memory atomics)
        order, cell_pass, rpt, capacity = "morton", "push", int(capacity[4:]), None
    elif method == "rows" and capacity == "geometry":   # cell pass from the plan's geometry table instead of the coordinates
        geometry, capacity = True, None
    elif method == "rows" and capacity:
        order, capacity = capacity, None
    plan = assemble.build_plan(mesh, ctags, ftags, ds(100), method=method, capacity=capacity, order=order,
                               geometry=geometry, cell_pass=cell_pass, rows_per_tile=rpt)
    assert plan.method == method
    if method == "rows":
        assert (plan.rowsplan.cell_geom is not None) == geometry
        assert (plan.rowsplan.tiles is not None) == (cell_pass in ("tiles", "push"))
        assert (plan.rowsplan.tiles is not None and plan.rowsplan.tiles.push is not None) == (cell_pass == "push")
    A, b = assemble.assemble_strong_dirichlet(plan, phi, f, stab_coef=1.0)
    if method in ("blocked", "rows"):
        # owner-computes sums in a fixed order: bitwise reproducible, and independent of what the
        # output buffers held before (no zero-fill needed for the matrix)
        data2 = torch.full_like(A.data, float("nan"))
        b2 = torch.zeros_like(b)
        assemble.assemble_into(plan, phi, f, 1.0, data2, b2)
        if cell_pass == "push":   # unordered shared-memory atomics: equal to rounding; still no zero-fill needed
            assert not torch.isnan(data2).any()
            assert (data2 - A.data).abs().max() <= 1e-14 * A.data.abs().max()
            assert (b2 - b).abs().max() <= 1e-14 * b.abs().max()
        else:
            assert torch.equal(data2, A.data) and torch.equal(b2, b)
        if capacity:
            assert plan.blocked.n_blocks > 4
        if method == "rows":
            assert plan.rowsplan.order == order and plan.rowsplan.n_cell_records > 0

    x = mesh.x.cpu().numpy()
    cells = mesh.cells.cpu().numpy().astype(np.int64)
    ip, ix, data, bo = OA.assemble_strong_dirichlet(
        x, cells, cells, len(x), phi.cpu().numpy(), f.cpu().numpy(), ctags.values_dev.cpu().numpy(),
        ftags.values_dev.cpu().numpy(), mesh.c2f.cpu().numpy(), mesh.f2c.cpu().numpy(),
        ds(100).integration_entities, sigma=1.0)
    # identical sparsity, structural zeros and empty rows included
    assert np.array_equal(A.indptr.cpu().numpy(), ip)
    assert np.array_equal(A.indices.cpu().numpy(), ix)
    assert plan.ghost.numel() > 0 and plan.entities.shape[0] > 0
    got = A.data.cpu().numpy()
    scale, rows = _row_scale(ip, data)
    assert np.all(np.abs(got - data) <= RTOL * scale[rows])
    bg = b.cpu().numpy()
    assert np.all(np.abs(bg - bo) <= RTOL * max(np.abs(bo).max(), 1e-300))


def test_cuda_each_integral_separately():
    """Cell, one-sided boundary and ghost-penalty kernels checked one by one against the oracle's
    element tensors (catches compensating errors)."""
    mesh, phi, f = _setup("tet-unstructured", 6)
    fn = fem.Function(fem.functionspace(mesh, 1), phi.cpu().numpy())
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", RuntimeWarning)
        ctags, ftags, _, ds, _ = mesh_scripts.compute_tags_measures(mesh, fn, 1, box_mode=True)
    x = mesh.x.cpu().numpy()
    cells = mesh.cells.cpu().numpy().astype(np.int64)
    ph, fh = phi.cpu().numpy(), f.cpu().numpy()
    ct = ctags.values_dev.cpu().numpy()
    n = len(x)
    import scipy.sparse as sp

    def dense_from(dofs, tensors):
        k = dofs.shape[1]
        r = np.repeat(dofs, k, axis=1).ravel()
        c = np.tile(dofs, (1, k)).ravel()
        return sp.coo_matrix((tensors.ravel(), (r, c)), shape=(n, n)).tocsr()

    from phifem_b200 import _lib
    lib = _lib.load()
    plan = assemble.build_plan(mesh, ctags, ftags, ds(100), method="atomic")
    cm = _lib.c_mesh(mesh)
    st = _lib.stream()

    def run(kernel):
        data, b = plan.new_outputs()
        kernel(data, b)
        torch.cuda.synchronize()
        return sp.csr_matrix((data.cpu().numpy(), plan.indices.cpu().numpy(), plan.indptr.cpu().numpy()),
                             shape=(n, n)), b.cpu().numpy()

    active = plan.active.cpu().numpy()
    Ac, bc = OA.cell_tensors_closed_form(x, cells[active], ph, fh, ct[active] == 2, 0.9)
    got, gb = run(lambda d, b: _lib.check(lib.phifem_assemble_cells_p1(
        cm, _lib.ptr(phi), _lib.ptr(f), _lib.ptr(plan.cell_tags8), _lib.ptr(plan.active), len(active),
        _lib.ptr(plan.slots_cells), 0.9, _lib.ptr(d), _lib.ptr(b), st)))
    want = dense_from(cells[active], Ac)
    assert abs(got - want).max() <= RTOL * abs(want).max()
    wb = np.zeros(n)
    np.add.at(wb, cells[active].ravel(), bc.ravel())
    assert np.abs(gb - wb).max() <= RTOL * np.abs(wb).max()

    ents = plan.entities.cpu().numpy()
    Ab = OA.boundary_tensors_closed_form(x, cells, ph, ents)
    got, _ = run(lambda d, b: _lib.check(lib.phifem_assemble_boundary_p1(
        cm, _lib.ptr(phi), _lib.ptr(plan.entities), len(ents), _lib.ptr(plan.slots_boundary),
        _lib.ptr(d), st)))
    want = dense_from(cells[ents[:, 0]], Ab)
    assert abs(want).max() > 0 and abs(got - want).max() <= RTOL * abs(want).max()

    ghost = plan.ghost.cpu().numpy()
    Eg, macro = OA.ghost_tensors_closed_form(x, cells, ph, mesh.c2f.cpu().numpy(), mesh.f2c.cpu().numpy(),
                                             ghost, 0.9)
    got, _ = run(lambda d, b: _lib.check(lib.phifem_assemble_ghost_p1(
        cm, _lib.ptr(phi), _lib.ptr(plan.ghost), len(ghost), _lib.ptr(plan.slots_ghost), 0.9,
        _lib.ptr(d), st)))
    want = dense_from(macro, Eg)
    assert abs(want).max() > 0 and abs(got - want).max() <= RTOL * abs(want).max()


def test_cuda_assembly_properties_at_scale():
    """Size-independent properties on a mesh the oracle cannot finish quickly (1.3 M tetrahedra):
    constants are in the kernel of the stiffness part => with phi == 1 every row of A sums to ~0 and
    A is symmetric without the boundary term; sum(b) == int f phi over Omega_h."""
    mesh = synthetic.box_mesh(60, device="cuda")
    phi_tag = synthetic.sphere_levelset(mesh.x)
    fn = fem.Function(fem.functionspace(mesh, 1), phi_tag.cpu().numpy())
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", RuntimeWarning)
        ctags, ftags, _, ds, _ = mesh_scripts.compute_tags_measures(mesh, fn, 1, box_mode=True)
    plan = assemble.build_plan(mesh, ctags, ftags, None, method="rows")   # no one-sided term
    assert plan.method == "rows" and plan.rowsplan.cells.n_slices > 148
    one = torch.ones(mesh.num_vertices, dtype=torch.float64, device="cuda")
    A, b = assemble.assemble_strong_dirichlet(plan, one, one, stab_coef=1.0)
    M = A.to_scipy()
    # phi == 1: a(w,v) = int grad w . grad v + ghost jumps; constants are in its kernel
    rs = np.abs(M @ np.ones(M.shape[0]))
    assert rs.max() <= 1e-10 * abs(M).max()
    assert abs(M - M.T).max() <= 1e-12 * abs(M).max()
    # sum_i b_i = int_{Omega_h} f phi = |Omega_h| (stabilisation term vanishes: grad phi = 0)
    tags = ctags.values_dev
    vol_cell = (1.0 / 60) ** 3 / 6.0
    want = float(((tags == 1) | (tags == 2)).sum()) * vol_cell
    assert abs(float(b.sum()) - want) <= 1e-10 * want
    # sphere of radius 0.45: |Omega_h| slightly above the ball volume
    assert 0.38 < want < 0.42


@pytest.mark.parametrize("kind,n", [("tri-unstructured", 28), ("tet-unstructured", 9)])
def test_submesh_route_equals_box_mode(kind, n):
    """`python main.py sub` vs `python main.py bg` of the demo (main.py:58-70): the operator assembled on the
    submesh of Omega_h with the plain `ds` measure equals the box-mode operator restricted to the dofs of
    Omega_h (same sparsity, same entries), and the box-mode rows outside Omega_h are empty."""
    mesh, phi, f = _setup(kind, n)
    fn = fem.Function(fem.functionspace(mesh, 1), phi.cpu().numpy())
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", RuntimeWarning)
        ctags, ftags, _, ds_bdy, _ = mesh_scripts.compute_tags_measures(mesh, fn, 1, box_mode=True)
        sct, sft, sub, ds_sub, maps = mesh_scripts.compute_tags_measures(mesh, fn, 1, box_mode=False)
    A, b = assemble.assemble_strong_dirichlet(assemble.build_plan(mesh, ctags, ftags, ds_bdy(100)), phi, f)
    v_map = torch.as_tensor(maps[1].astype(np.int64), device="cuda")
    As, bs = assemble.assemble_strong_dirichlet(assemble.build_plan(sub, sct, sft, ds_sub), phi[v_map], f[v_map])
    M, Ms = A.to_scipy().tocsr(), As.to_scipy().tocsr()
    vm = maps[1].astype(np.int64)
    outside = np.setdiff1d(np.arange(mesh.num_vertices), vm)
    assert np.all(np.diff(M.indptr)[outside] == 0) and np.all(b.cpu().numpy()[outside] == 0.0)
    R = M[vm][:, vm].tocsr()
    R.sort_indices()
    Ms.sort_indices()
    assert np.array_equal(R.indptr, Ms.indptr) and np.array_equal(R.indices, Ms.indices)
    scale = np.abs(R.data).max()
    assert np.abs(R.data - Ms.data).max() <= 1e-12 * scale
    assert np.abs(b.cpu().numpy()[vm] - bs.cpu().numpy()).max() <= 1e-12 * float(b.abs().max())
