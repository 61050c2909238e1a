"""GPU parity: CUDA classification (through the C ABI) vs the CPU oracle and the reference goldens.

Bit-exact bar: cell tags, facet tags, one-sided entity lists and submesh transfers must equal the
oracle's on every case of the reference's own test matrix (tests/test_compute_meshtags.py), and the
reference goldens wherever the oracle is pinned to them."""
import itertools
import warnings

import numpy as np
import pytest
import torch

import cases
import oracle_driver as od
from oracle import tags as OT
from phifem_b200 import fem, mesh_scripts, synthetic
from phifem_b200.mesh import Mesh

pytestmark = pytest.mark.gpu

PARAMS = [(d, N, disc, single, box)
          for d in cases.TAG_DATA
          for N, disc, single, box in itertools.product((1, 2, 3), (True, False), (True, False),
                                                        (True, False))]


def _id(p):
    d, N, disc, single, box = p
    return "%s-%d-%s-%s-%s" % (d[0], N, "disc" if disc else "expr", "single" if single else "multi",
                               "box" if box else "sub")


def _gpu_levelset(mesh, levelset, degree, discretize):
    if not discretize:
        return levelset
    V = fem.functionspace(mesh, ("Lagrange", degree))
    return fem.Function(V).interpolate(levelset)


@pytest.mark.parametrize("p", PARAMS, ids=_id)
def test_cuda_tags_match_oracle_and_goldens(p):
    (name, mesh_name, levelset), N, disc, single, box = p
    x, cells, ct = cases.load_mesh_arrays(mesh_name)
    mesh = Mesh(x, cells, ct, device="cuda")
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", RuntimeWarning)
        ctags, ftags, sub, ds, maps = mesh_scripts.compute_tags_measures(
            mesh, _gpu_levelset(mesh, levelset, N, disc), N, box_mode=box, single_layer_cut=single)
    out = od.run_oracle(x, cells, ct, levelset, N, disc, box, single)
    # topology: same numbering as the oracle's
    assert np.array_equal(mesh.c2f.cpu().numpy(), out["c2f"])
    assert np.array_equal(mesh.f2c.cpu().numpy(), out["f2c"])
    for mine, want in ((ctags, out["cell_tags"]), (ftags, out["facet_tags"])):
        idx = np.nonzero(want)[0]
        assert mine.indices.dtype == np.int32 and mine.values.dtype == np.int32
        assert np.array_equal(mine.indices, idx)
        assert np.array_equal(mine.values, want[idx])
    if box:
        assert sub is None and maps is None
        assert np.array_equal(ds(100).integration_entities, out["ds100"])
        assert np.array_equal(ds(101).integration_entities, out["ds101"])
    else:
        assert np.array_equal(maps[0], out["c_map"])
        assert np.array_equal(maps[1], out["v_map"])
        assert np.array_equal(sub.cells.cpu().numpy(), out["sub_cells"])
        assert np.array_equal(sub.x.cpu().numpy(), out["sub_x"])
    # and the reference's golden files where its numbering is reproducible
    if cases.case_class(name, mesh_name, N, disc) == "exact":
        cname, fname = cases.golden_names(name, N, disc, box, single)
        for mine, gold in ((ctags, cases.golden(cname)), (ftags, cases.golden(fname))):
            assert np.array_equal(mine.indices, gold[0])
            assert np.array_equal(mine.values, gold[1])


@pytest.mark.parametrize("discretize", [True, False])
@pytest.mark.parametrize("degree", [1, 2, 3])
@pytest.mark.parametrize("data", cases.ONE_SIDED, ids=lambda d: d[0])
def test_cuda_one_sided_integrals(data, degree, discretize):
    """reference tests/test_one_sided_integral.py: known answers of ds(100) / ds(101)."""
    name, mesh_name, levelset, expected, kind = data
    x, cells, ct = cases.load_mesh_arrays(mesh_name)
    mesh = Mesh(x, cells, ct, device="cuda")
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", RuntimeWarning)
        _, _, _, ds, _ = mesh_scripts.compute_tags_measures(
            mesh, _gpu_levelset(mesh, levelset, degree, discretize), degree, box_mode=True)
    for sid, want in ((100, expected[0]), (101, expected[1])):
        n, meas = OT.outward_normals(x, cells, ct, ds(sid).integration_entities)
        w = n[:, 0] + n[:, 1] if kind == "signed" else np.abs(n[:, 0]) + np.abs(n[:, 1])
        assert np.isclose((w * meas).sum(), want, atol=1e-20)


def _oracle_p1(mesh_np, phi, single=False):
    x, cells, ct = mesh_np
    pts = OT.cell_detection_points(ct, 1)
    fpts = OT.facet_points_in_cell(ct, 1)
    ftab = np.asarray([OT.coordinate_basis(ct, p)[0] for p in fpts])
    return OT.compute_tags_measures(x, cells, ct, phi[cells], OT.point_values_function(phi, cells, ftab),
                                    box_mode=True, single_layer_cut=single, detection_points=pts)


@pytest.mark.parametrize("kind", ["tri", "tet", "tet-unstructured"])
@pytest.mark.parametrize("single", [False, True])
def test_cuda_p1_fast_kernel_vs_oracle_synthetic(kind, single):
    """The headline kernel (P1, detection degree 1) on synthetic meshes incl. the 3D extension."""
    if kind == "tri":
        mesh = synthetic.rectangle_mesh(48, device="cuda")
        phi = synthetic.sphere_levelset(mesh.x, center=(0.013, -0.021), radius=0.61)
    else:
        mesh = synthetic.box_mesh(14, device="cuda")
        if kind == "tet-unstructured":
            mesh = synthetic.unstructured_variant(mesh, jitter=0.2, seed=4)
        phi = synthetic.sphere_levelset(mesh.x)
    V = fem.functionspace(mesh, ("Lagrange", 1))
    fn = fem.Function(V, phi.cpu().numpy())
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", RuntimeWarning)
        ctags, ftags, _, ds, _ = mesh_scripts.compute_tags_measures(mesh, fn, 1, box_mode=True,
                                                                    single_layer_cut=single)
    out = _oracle_p1((mesh.x.cpu().numpy(), mesh.cells.cpu().numpy().astype(np.int64), mesh.cell_type),
                     phi.cpu().numpy(), single)
    assert np.array_equal(ctags.values_dev.cpu().numpy(), out["cell_tags"])
    assert np.array_equal(ftags.values_dev.cpu().numpy(), out["facet_tags"])
    assert np.array_equal(ds(100).integration_entities, out["ds100"])
    assert np.array_equal(ds(101).integration_entities, out["ds101"])
    assert set(np.unique(out["cell_tags"])) == {1, 2, 3}


@pytest.mark.parametrize("kind", ["tet", "tet-tube", "tri"])
def test_cuda_staged_kernels_and_boundary_records_match_the_plain_kernels(kind, monkeypatch):
    """The TMA-staged classifiers (several tiles per CTA and stage: the pipeline wraps around) and the mesh-boundary
    pass from per-mesh records against the per-thread vector-load kernels and the on-the-fly boundary walk: tags and
    counters bit for bit.  "tet-tube" puts the interface through the mesh boundary (ds detection finds cut owners)."""
    if kind == "tri":
        mesh = synthetic.rectangle_mesh(1500, device="cuda")                # 4.5 M triangles, 6.8 M facets
        phi = synthetic.sphere_levelset(mesh.x, center=(0.013, -0.021), radius=0.61)
    else:
        mesh = synthetic.box_mesh(96, device="cuda")                        # 5.3 M tetrahedra, 10.7 M facets
        phi = synthetic.sphere_levelset(mesh.x)
        if kind == "tet-tube":
            lo, hi = mesh.x.min(dim=0).values, mesh.x.max(dim=0).values
            mid = 0.5 * (lo + hi)
            r2 = ((mesh.x[:, 1:] - mid[1:]) ** 2).sum(dim=1)
            phi = torch.minimum(phi, r2 - (0.21 * float(hi[1] - lo[1])) ** 2)   # a tube along x through both faces
    fn = fem.Function(fem.functionspace_p1_device(mesh), phi)
    dls = mesh_scripts._DeviceLevelset(mesh, fn, 1)
    got = {}
    for variant, env in (("plain", {"PHIFEM_CELLS_KERNEL": "ldg", "PHIFEM_FACETS_KERNEL": "ldg",
                                    "PHIFEM_BOUNDARY_KERNEL": "walk"}),
                         ("default", {})):
        for k in ("PHIFEM_CELLS_KERNEL", "PHIFEM_FACETS_KERNEL", "PHIFEM_BOUNDARY_KERNEL"):
            monkeypatch.delenv(k, raising=False)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        for rep in range(3):   # the race this test pins down (stage refilled under pending reads) was rare
            ws = mesh_scripts.classify(mesh, dls, ws=None)
            torch.cuda.synchronize()
            cur = (ws.cell_tags8.clone(), ws.facet_tags8.clone(), ws.counters.clone())
            if variant in got:
                for a, b in zip(got[variant], cur):
                    assert torch.equal(a, b)
            got[variant] = cur
    for a, b in zip(got["plain"], got["default"]):
        assert torch.equal(a, b)
    if kind == "tet-tube":
        bf = mesh.boundary_facets.long()
        assert int((got["default"][1][bf] == 2).sum()) > 0     # cut facets on the mesh boundary exist


def test_cuda_degenerate_values_follow_the_exact_equality_rule():
    """zeros do not make a cell cut; all-zero / NaN cells are cut; tiny negative terms
    absorbed by rounding keep the cell exterior -- bit-identical to the oracle."""
    mesh = synthetic.rectangle_mesh(8, device="cuda")
    rng = np.random.default_rng(0)
    phi = rng.choice([0.0, 0.0, 1.0, -1.0, 0.17, -1.1e-16, 1e-300, -1e-300, np.nan, 1e200, -0.025],
                     size=mesh.num_vertices)
    fn = fem.Function(fem.functionspace(mesh, 1), phi)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", RuntimeWarning)
        ctags, ftags, _, _, _ = mesh_scripts.compute_tags_measures(mesh, fn, 1, box_mode=True)
    out = _oracle_p1((mesh.x.cpu().numpy(), mesh.cells.cpu().numpy().astype(np.int64), "triangle"), phi)
    assert np.array_equal(ctags.values_dev.cpu().numpy(), out["cell_tags"])
    assert np.array_equal(ftags.values_dev.cpu().numpy(), out["facet_tags"])


def test_cuda_api_errors_and_warnings():
    x, cells, ct = cases.load_mesh_arrays("coarse_square")
    mesh = Mesh(x, cells, ct, device="cuda")
    ls = cases.TAG_DATA[5][2]
    with pytest.warns(RuntimeWarning):
        ctags, ftags, _, _, _ = mesh_scripts.compute_tags_measures(mesh, ls, 1, box_mode=True)
    from phifem_b200.mesh import MeshTags
    bad = MeshTags.from_lists(mesh, 2, [0, 1], [2, 7])
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", RuntimeWarning)
        with pytest.raises(ValueError, match="Cannot overwrite cells tags"):
            mesh_scripts.compute_tags_measures(mesh, ls, 1, box_mode=True, overwrite_tags={"cells": bad})
        ok = MeshTags.from_lists(mesh, 2, [0, 5], [7, 9])
        c2, _, _, _, _ = mesh_scripts.compute_tags_measures(mesh, ls, 1, box_mode=True,
                                                         overwrite_tags={"cells": ok})
    want = ctags.values.copy()
    want[[0, 5]] = [7, 9]
    assert np.array_equal(c2.values, want)
    assert np.array_equal(c2.find(9), [5])
    with pytest.raises(ValueError, match="source_mesh"):
        mesh_scripts._transfer_tags(ftags, mesh, np.arange(3))
