"""Edge cases of the path on the GPU against the oracle: no active cell at all, no exterior cell at all (the
`len(exterior_cells) == 0` branch of reference src/phifem/mesh_scripts.py:469-474, where the mesh-boundary facets
become Gamma_h), a level set that only grazes the mesh boundary, a one-cell mesh, single_layer_cut removing every
cut cell."""
import warnings

import numpy as np
import pytest
import torch

from oracle import assembly as OA
from oracle import tags as OT
from phifem_b200 import assemble, fem, mesh_scripts, solve, synthetic
from phifem_b200.mesh import Mesh

pytestmark = pytest.mark.gpu


def _run(mesh, phi, single=False):
    fn = fem.Function(fem.functionspace(mesh, 1), phi.cpu().numpy())
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", RuntimeWarning)
        return mesh_scripts.compute_tags_measures(mesh, fn, 1, box_mode=True, single_layer_cut=single)


def _oracle(mesh, phi, single=False):
    x = mesh.x.cpu().numpy()
    cells = mesh.cells.cpu().numpy().astype(np.int64)
    ph = phi.cpu().numpy()
    ct = mesh.cell_type
    pts = OT.cell_detection_points(ct, 1)
    ftab = np.asarray([OT.coordinate_basis(ct, p)[0] for p in OT.facet_points_in_cell(ct, 1)])
    return x, cells, ph, OT.compute_tags_measures(x, cells, ct, ph[cells], OT.point_values_function(ph, cells, ftab),
                                                  box_mode=True, single_layer_cut=single, detection_points=pts)


def _check_against_oracle(mesh, phi, single=False, f=None):
    ctags, ftags, _, ds, _ = _run(mesh, phi, single)
    x, cells, ph, out = _oracle(mesh, phi, single)
    assert np.array_equal(ctags.values_dev.cpu().numpy(), out["cell_tags"])
    assert np.array_equal(ftags.values_dev.cpu().numpy(), out["facet_tags"])
    assert np.array_equal(ds(100).integration_entities, np.asarray(out["ds100"]).ravel())
    assert np.array_equal(ds(101).integration_entities, np.asarray(out["ds101"]).ravel())
    f = torch.ones_like(phi) if f is None else f
    plan = assemble.build_plan(mesh, ctags, ftags, ds(100))
    A, b = assemble.assemble_strong_dirichlet(plan, phi, f, stab_coef=1.0)
    ip, ix, data, bo = OA.assemble_strong_dirichlet(x, cells, cells, len(x), ph, f.cpu().numpy(), out["cell_tags"],
                                                    out["facet_tags"], out["c2f"], out["f2c"], out["ds100"])
    assert np.array_equal(A.indptr.cpu().numpy(), ip) and np.array_equal(A.indices.cpu().numpy(), ix)
    if len(data):
        assert np.abs(A.data.cpu().numpy() - data).max() <= 1e-12 * np.abs(data).max()
        assert np.abs(b.cpu().numpy() - bo).max() <= 1e-12 * max(np.abs(bo).max(), 1e-300)
    return out, plan, A, b


@pytest.mark.parametrize("kind", ["tri", "tet"])
def test_no_active_cell(kind):
    mesh = synthetic.rectangle_mesh(6, device="cuda") if kind == "tri" else synthetic.box_mesh(3, device="cuda")
    phi = synthetic.sphere_levelset(mesh.x, center=(9.0, 9.0, 9.0), radius=0.5)      # positive everywhere
    out, plan, A, b = _check_against_oracle(mesh, phi)
    assert np.all(out["cell_tags"] == 3) and np.all(out["facet_tags"] == 5)
    assert plan.nnz == 0 and A.data.numel() == 0 and float(b.abs().max()) == 0.0
    x, info = solve.bicgstab(A, b)
    assert info.n_active == 0 and float(x.abs().max()) == 0.0


@pytest.mark.parametrize("kind", ["tri", "tet"])
def test_no_exterior_cell(kind):
    """phi < 0 on the whole box: every cell interior, the mesh boundary is Gamma_h (:469-470)."""
    mesh = synthetic.rectangle_mesh(6, device="cuda") if kind == "tri" else synthetic.box_mesh(3, device="cuda")
    phi = synthetic.sphere_levelset(mesh.x, center=(0.3, 0.4, 0.5), radius=9.0)
    out, plan, A, b = _check_against_oracle(mesh, phi)
    assert np.all(out["cell_tags"] == 1)
    bnd = out["f2c"][:, 1] < 0
    assert np.all(out["facet_tags"][bnd] == 4) and np.all(out["facet_tags"][~bnd] == 1)
    assert plan.entities.shape[0] == int(bnd.sum()) and plan.ghost.numel() == 0


@pytest.mark.parametrize("kind,single", [("tri", False), ("tri", True), ("tet", False), ("tet", True)])
def test_level_set_cutting_the_mesh_boundary(kind, single):
    """A ball centred on a corner: cut cells own mesh-boundary facets (the ds detection of :434-452 decides 2 vs 4)."""
    mesh = synthetic.rectangle_mesh(9, device="cuda") if kind == "tri" else synthetic.box_mesh(5, device="cuda")
    mesh = synthetic.unstructured_variant(mesh, jitter=0.15, seed=21)
    lo = mesh.x.min(dim=0).values
    phi = ((mesh.x - lo) ** 2).sum(dim=1) - (0.83 if kind == "tri" else 0.61) ** 2
    out, plan, A, b = _check_against_oracle(mesh, phi, single)
    bnd = out["f2c"][:, 1] < 0
    assert set(np.unique(out["facet_tags"][bnd])) >= {1, 2, 5}       # interior, cut and exterior boundary facets


def test_single_cell_meshes():
    tri = Mesh(np.array([[0.0, 0.0], [1.0, 0.0], [0.0, 1.0]]), np.array([[0, 1, 2]]), "triangle", device="cuda")
    tet = Mesh(np.array([[0.0, 0, 0], [1.0, 0, 0], [0, 1.0, 0], [0, 0, 1.0]]), np.array([[0, 1, 2, 3]]),
               "tetrahedron", device="cuda")
    for mesh in (tri, tet):
        for shift in (-0.4, 0.3, 5.0):      # cut, cut, exterior
            phi = mesh.x[:, 0] + 0.5 * mesh.x[:, 1] - 0.45 + shift
            _check_against_oracle(mesh, phi)


def test_single_layer_cut_can_remove_every_cut_cell():
    """A ball smaller than a cell: cut cells without interior neighbour are re-tagged exterior (:349-358)."""
    mesh = synthetic.box_mesh(4, device="cuda")
    phi = synthetic.sphere_levelset(mesh.x, center=(0.52, 0.47, 0.51), radius=0.09)
    out, _, _, _ = _check_against_oracle(mesh, phi, single=False)
    assert (out["cell_tags"] == 2).sum() > 0 and (out["cell_tags"] == 1).sum() == 0
    out, plan, A, _ = _check_against_oracle(mesh, phi, single=True)
    assert np.all(out["cell_tags"] == 3) and plan.nnz == 0


def test_no_active_cell_quadrature_and_mixed_operators():
    mesh = synthetic.rectangle_mesh(5, device="cuda")
    phi1 = synthetic.sphere_levelset(mesh.x, center=(9.0, 9.0), radius=0.5)
    ctags, ftags, _, ds, _ = _run(mesh, phi1)
    V = fem.functionspace(mesh, 2)
    phi = synthetic.sphere_levelset(V.dof_coordinates_dev(), center=(9.0, 9.0), radius=0.5)
    f = torch.ones(V.num_dofs, dtype=torch.float64, device="cuda")
    plan = assemble.build_plan(mesh, ctags, ftags, ds(100), V=V, V_phi=V)
    A, b = assemble.assemble_strong_dirichlet(plan, phi, f)
    assert plan.nnz == 0 and A.shape == (V.num_dofs, V.num_dofs) and float(b.abs().max()) == 0.0
    planw = assemble.build_plan_weak_dirichlet(mesh, ctags, ftags, ds(100), V=V)
    Aw, bw = assemble.assemble_weak_dirichlet(planw, phi, f)
    assert planw.nnz == 0 and Aw.shape == (2 * V.num_dofs, 2 * V.num_dofs) and float(bw.abs().max()) == 0.0


def test_overwrite_values_beyond_int8_match_nothing():
    """A facet tagged 260 must not be picked up as Gamma_h (4) by the ds(100) search, a cell tagged 257 / 258 must leave
    Omega_h (reference :606-615 allows any value but 1..6 / 100 / 101; `find(4)` / `find(1)` do not match them)."""
    from phifem_b200.mesh import MeshTags
    mesh = synthetic.unstructured_variant(synthetic.rectangle_mesh(14, device="cuda"), seed=5)
    phi = synthetic.sphere_levelset(mesh.x, center=(0.03, -0.02), radius=0.55)
    ctags, ftags, _, ds, _ = _run(mesh, phi)
    x, cells, ph, out = _oracle(mesh, phi)
    g_facets = np.nonzero(out["facet_tags"] == 4)[0][::2]
    in_cells, cut_cells = np.nonzero(out["cell_tags"] == 1)[0][::3], np.nonzero(out["cell_tags"] == 2)[0][::2]
    assert len(g_facets) and len(in_cells) and len(cut_cells)
    oc = MeshTags.from_lists(mesh, 2, np.concatenate([in_cells, cut_cells]),
                             np.concatenate([np.full(len(in_cells), 257), np.full(len(cut_cells), 258)]))
    of = MeshTags.from_lists(mesh, 1, g_facets, np.full(len(g_facets), 260))
    fn = fem.Function(fem.functionspace(mesh, 1), phi.cpu().numpy())
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", RuntimeWarning)
        c2, f2, _, ds2, _ = mesh_scripts.compute_tags_measures(mesh, fn, 1, box_mode=True,
                                                               overwrite_tags={"cells": oc, "facets": of})
        s2 = mesh_scripts.compute_tags_measures(mesh, fn, 1, box_mode=False, overwrite_tags={"cells": oc, "facets": of})
    ct_o, ft_o = out["cell_tags"].copy(), out["facet_tags"].copy()
    ct_o[in_cells], ct_o[cut_cells], ft_o[g_facets] = 257, 258, 260
    assert np.array_equal(c2.values_dev.cpu().numpy(), ct_o) and np.array_equal(f2.values_dev.cpu().numpy(), ft_o)
    want100 = OT.integration_entities(out["c2f"], out["f2c"], (ct_o == 1) | (ct_o == 2), ft_o == 4)
    want101 = OT.integration_entities(out["c2f"], out["f2c"], (ct_o == 2) | (ct_o == 3), ft_o == 3)
    assert np.array_equal(ds2(100).integration_entities, want100)
    assert np.array_equal(ds2(101).integration_entities, want101)
    assert np.array_equal(s2[4][0], np.nonzero((ct_o == 1) | (ct_o == 2))[0])      # submesh = cells still tagged 1 / 2
    plan = assemble.build_plan(mesh, c2, f2, ds2(100))
    assert np.array_equal(np.sort(plan.active.cpu().numpy()), np.nonzero((ct_o == 1) | (ct_o == 2))[0])


def test_counters_posted_to_the_host_equal_a_copy_and_ignore_other_streams():
    """`phifem_post_to_host` (the one host synchronisation of compute_tags_measures): the words a kernel stores into
    pinned host memory equal a device -> host copy of the counter block, from a side stream too; a destination the
    device cannot address is refused."""
    from phifem_b200 import _lib
    mesh = synthetic.box_mesh(12, device="cuda")
    phi = synthetic.sphere_levelset(mesh.x)
    dls = mesh_scripts._DeviceLevelset(mesh, fem.Function(fem.functionspace_p1_device(mesh), phi), 1)
    ws = mesh_scripts.classify(mesh, dls)
    want = ws.counters.cpu().numpy()
    assert np.array_equal(mesh_scripts._read_counters(ws), want) and want[:3].sum() == mesh.num_cells
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        ws2 = mesh_scripts.classify(mesh, dls)
        assert np.array_equal(mesh_scripts._read_counters(ws2), want)
    pageable = np.zeros(_lib.N_COUNTERS, dtype=np.int64)
    rc = _lib.load().phifem_post_to_host(_lib.ptr(ws.counters), pageable.ctypes.data, _lib.N_COUNTERS, _lib.stream())
    torch.cuda.synchronize()
    if rc == 0:      # a system that lets the device address pageable memory: the words arrive all the same
        assert np.array_equal(pageable, want)
    else:
        assert rc == -1 and b"page-locked" in _lib.load().phifem_last_error()


def test_plan_matches_on_the_device():
    """`AssemblyPlan.matches` through `phifem_tags_match`: same answers as the torch passes of the CPU path -- for the
    plan's own tags, for tags that differ in one cell / one facet class at the very end of the arrays, and for
    unaligned views (the byte path of the kernel)."""
    from phifem_b200 import _lib
    from phifem_b200.mesh import MeshTags
    mesh = synthetic.box_mesh(11, device="cuda")
    phi = synthetic.sphere_levelset(mesh.x, radius=0.37)
    ctags, ftags, _, ds, _ = _run(mesh, phi)
    plan = assemble.build_plan(mesh, ctags, ftags, ds(100))
    assert plan.matches(ctags, ftags)
    c8, f8 = ctags.tags8.clone(), ftags.tags8.clone()
    assert plan.matches(MeshTags(mesh, 3, None, tags8=c8), MeshTags(mesh, 2, None, tags8=f8))
    other = f8.clone()
    other[other == 5] = 6                                  # a tag no integral looks at
    assert plan.matches(ctags, MeshTags(mesh, 2, None, tags8=other))
    last_c = c8.clone()
    last_c[-1] = 1 if int(last_c[-1]) != 1 else 3
    assert not plan.matches(MeshTags(mesh, 3, None, tags8=last_c), ftags)
    last_f = f8.clone()
    last_f[-1] = 4 if int(last_f[-1]) != 4 else 5
    assert not plan.matches(ctags, MeshTags(mesh, 2, None, tags8=last_f))
    ctags2, ftags2, _, _, _ = _run(mesh, synthetic.sphere_levelset(mesh.x, radius=0.40))
    assert not plan.matches(ctags2, ftags2)
    # the kernel on unaligned arrays
    lib, bad = _lib.load(), torch.empty(1, dtype=torch.int64, device="cuda")
    sig_f = ((f8 == 2) | (f8 == 3)).to(torch.int8) + 2 * (f8 == 4).to(torch.int8)
    for shift, want in ((1, 0), (3, 0)):
        a, b, c, d = c8[shift:], c8.clone()[shift:], f8[shift:], sig_f[shift:]
        _lib.check(lib.phifem_tags_match(a.data_ptr(), b.data_ptr(), a.numel(), c.data_ptr(), d.data_ptr(), c.numel(),
                                         _lib.ptr(bad), _lib.stream()))
        assert int(bad.item()) == want
    d2 = sig_f.clone()
    d2[5] = (int(d2[5]) + 1) % 3
    _lib.check(lib.phifem_tags_match(c8[1:].data_ptr(), c8[1:].data_ptr(), c8.numel() - 1, f8[1:].data_ptr(),
                                     d2[1:].data_ptr(), f8.numel() - 1, _lib.ptr(bad), _lib.stream()))
    assert int(bad.item()) > 0
