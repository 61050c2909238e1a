"""GPU parity on an UNSTRUCTURED mesh of more than a million tetrahedra (SURVEY.md 8d's variant: jittered vertices,
permuted cells, relabelled vertices), renumbered along the Morton curve by `Mesh.reordered()` as the mesh-level
symbolic phase does (the reference's tests run on arbitrary XDMF meshes: tests/test_compute_meshtags.py:136-137).
The oracle (C restatement, oracle/csrc) works on the USER's arrays; the product works on the renumbered mesh; tags and
the CSR operator are compared through `original_cell_index` / `input_global_indices`, facets through their vertices."""
import warnings

import numpy as np
import pytest
import scipy.sparse as sp
import torch

from oracle import native as ON
from oracle import tags as OT
from phifem_b200 import assemble, fem, mesh_scripts, synthetic
from phifem_b200.mesh import Mesh, MeshTags

pytestmark = pytest.mark.gpu
N = 56            # 6 * 56^3 = 1 053 696 tetrahedra


@pytest.fixture(scope="module")
def problem():
    base = synthetic.unstructured_variant_device(synthetic.box_mesh(N, device="cuda"), jitter=0.2, seed=0)
    mesh = base.reordered()
    assert mesh.num_cells == 6 * N ** 3 > 10 ** 6 and mesh.sfc_ordered
    phi = synthetic.sphere_levelset(mesh.x)
    f = torch.from_numpy(np.random.default_rng(1234).uniform(-1, 1, mesh.num_vertices)).cuda()
    fn = fem.Function(fem.functionspace_p1_device(mesh), phi)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", RuntimeWarning)
        ctags, ftags, _, ds, _ = mesh_scripts.compute_tags_measures(mesh, fn, 1, box_mode=True)
    # ---- oracle on the user's numbering ----------------------------------------------------------------
    igi = mesh.input_global_indices.cpu().numpy()
    oci = mesh.original_cell_index.cpu().numpy()
    host = Mesh(base.x.cpu(), base.cells.cpu(), "tetrahedron", device="cpu")
    x, cells = host.x.numpy(), np.ascontiguousarray(host.cells.numpy())
    c2f, f2c = np.ascontiguousarray(host.c2f.numpy()), np.ascontiguousarray(host.f2c.numpy())
    phi_o = np.empty(len(x))
    phi_o[igi] = phi.cpu().numpy()
    # (the level set is data of the mesh numbering: the oracle gets the SAME values, moved to the user's numbering;
    # re-evaluating it on the host would differ in the last bit, the device contracts a*a + b into an FMA)
    assert np.abs(phi_o - synthetic.sphere_levelset(host.x).numpy()).max() < 1e-15
    f_o = np.empty(len(x))
    f_o[igi] = f.cpu().numpy()
    ct_o = ON.tag_cells_p1(x, cells, phi_o)
    ft_o = ON.tag_facets_p1(x, cells, c2f, f2c, phi_o, ct_o)
    return dict(base=base, mesh=mesh, phi=phi, f=f, ctags=ctags, ftags=ftags, ds=ds, igi=igi, oci=oci, host=host,
                x=x, cells=cells, c2f=c2f, f2c=f2c, phi_o=phi_o, f_o=f_o, ct_o=ct_o, ft_o=ft_o)


def _facet_map(p):
    """For every facet of the renumbered mesh, the index of the same facet (same vertices) of the user's mesh."""
    fv_new = p["igi"][p["mesh"].facet_vertices.cpu().numpy().astype(np.int64)]
    fv_new.sort(axis=1)
    fv_old = p["host"].facet_vertices.numpy().astype(np.int64)               # rows sorted, lexicographic order
    nv = len(p["x"])
    key_old = (fv_old[:, 0] * nv + fv_old[:, 1]) * nv + fv_old[:, 2]
    key_new = (fv_new[:, 0] * nv + fv_new[:, 1]) * nv + fv_new[:, 2]
    idx = np.searchsorted(key_old, key_new)
    assert np.array_equal(key_old[idx], key_new)
    return idx


def test_unstructured_tags_equal_the_oracle(problem):
    p = problem
    got_c = p["ctags"].values_dev.cpu().numpy()
    assert np.array_equal(got_c, p["ct_o"][p["oci"]])
    hist = np.bincount(got_c, minlength=4)
    assert hist[1] > 3e5 and hist[2] > 3e4 and hist[3] > 5e5
    fmap = _facet_map(p)
    assert np.array_equal(p["ftags"].values_dev.cpu().numpy(), p["ft_o"][fmap])
    # ds(100): the same (cell, local facet) SET (the order follows the facet numbering, which the renumbering changes)
    ents_o = OT.integration_entities(p["c2f"], p["f2c"], (p["ct_o"] == 1) | (p["ct_o"] == 2), p["ft_o"] == 4)
    got = p["ds"](100).integration_entities.reshape(-1, 2)
    got_old = np.stack([p["oci"][got[:, 0]], got[:, 1]], axis=1)
    assert sorted(map(tuple, got_old)) == sorted(map(tuple, ents_o.reshape(-1, 2)))


@pytest.mark.parametrize("cell_pass", ["rows", "tiles", "push"])
def test_unstructured_operator_equals_the_oracle(problem, cell_pass):
    p = problem
    mesh = p["mesh"]
    plan = assemble.build_plan(mesh, p["ctags"], p["ftags"], p["ds"](100), cell_pass=cell_pass)
    assert plan.method == "rows" and plan.rowsplan.order == "natural"          # the mesh is already SFC-numbered
    A, b = assemble.assemble_strong_dirichlet(plan, p["phi"], p["f"], stab_coef=1.0)
    A2, b2 = assemble.assemble_strong_dirichlet(plan, p["phi"], p["f"], stab_coef=1.0)
    if cell_pass == "push":                                                     # unordered shared-memory atomics
        assert (A.data - A2.data).abs().max() <= 1e-14 * A.data.abs().max()
    else:
        assert torch.equal(A.data, A2.data) and torch.equal(b, b2)              # fixed summation order
    # oracle operator on the user's numbering (slot maps from the product's host plumbing on CPU tensors, which
    # tests/test_host_logic.py holds to the oracle's pattern)
    host = p["host"]
    ents_o = OT.integration_entities(p["c2f"], p["f2c"], (p["ct_o"] == 1) | (p["ct_o"] == 2), p["ft_o"] == 4)
    po = assemble.build_plan(host, MeshTags(host, 3, torch.from_numpy(p["ct_o"])),
                             MeshTags(host, 2, torch.from_numpy(p["ft_o"])), ents_o, method="atomic")
    arr = {k: np.ascontiguousarray(getattr(po, k).numpy())
           for k in ("active", "slots_cells", "entities", "slots_boundary", "ghost", "slots_ghost")}
    data_o, b_o = ON.assemble_p1(p["x"], p["cells"], p["c2f"], p["f2c"], p["phi_o"], p["f_o"], p["ct_o"], arr["active"],
                                 arr["slots_cells"], arr["entities"], arr["slots_boundary"], arr["ghost"],
                                 arr["slots_ghost"], 1.0, po.nnz)
    n = len(p["x"])
    want = sp.csr_matrix((data_o, po.indices.numpy(), po.indptr.numpy()), shape=(n, n))
    # product operator moved to the user's numbering: row / column v_new -> igi[v_new]
    igi = p["igi"]
    indptr = A.indptr.cpu().numpy()
    rows_new = np.repeat(np.arange(n), np.diff(indptr))
    got = sp.coo_matrix((A.data.cpu().numpy(), (igi[rows_new], igi[A.indices.cpu().numpy().astype(np.int64)])),
                        shape=(n, n)).tocsr()
    got.sort_indices()
    want.sort_indices()
    assert np.array_equal(got.indptr, want.indptr) and np.array_equal(got.indices, want.indices)   # identical sparsity
    scale = np.zeros(n)
    r = np.repeat(np.arange(n), np.diff(want.indptr))
    np.maximum.at(scale, r, np.abs(want.data))
    assert np.all(np.abs(got.data - want.data) <= 1e-12 * scale[r])
    b_new = b.cpu().numpy()
    b_got = np.empty(n)
    b_got[igi] = b_new
    assert np.abs(b_got - b_o).max() <= 1e-12 * np.abs(b_o).max()
