"""Consumes dolfinx-produced fixtures tests/golden/dolfinx_csr_<name>.npz (written by baseline/dolfinx_reference.py in a
dolfinx 0.9.0 environment: the reference's `compute_tags_measures` and the demos' literal UFL forms assembled by
dolfinx/PETSc) and holds the oracle -- and, on a GPU, the CUDA path -- to them: tags bit-exact, CSR sparsity identical,
values within 1e-12 of the row scale.

No such environment exists in this image (no dolfinx, no network), so no fixture is committed yet and these tests
SKIP, saying so: until then the assembled operator is pinned by tests/test_oracle_sympy.py only (element tensors
from the integrands) and the dolfinx conventions (dS '+' side, ds(100) owner, pattern) stay [dep-knowledge]."""
import glob
import os
import warnings

import numpy as np
import pytest

from oracle import assembly as OA
from oracle import tags as OT

HERE = os.path.dirname(os.path.abspath(__file__))
FIXTURES = sorted(glob.glob(os.path.join(HERE, "golden", "dolfinx_csr_*.npz")))
needs_fixture = pytest.mark.skipif(not FIXTURES, reason="no dolfinx-produced fixture committed (run "
                                   "baseline/dolfinx_reference.py fixtures where dolfinx 0.9.0 is installed): CSR parity "
                                   "is pinned by the sympy derivation only")


def _load(path):
    """The fixture in the oracle's terms: vertices numbered as dolfinx numbered the P1 dofs, so that cells == dofmap."""
    d = np.load(path)
    gd, vd = d["geometry_dofmap"], d["V_dofmap"]
    assert gd.shape == vd.shape, "P1 fixture expected"
    n = int(vd.max()) + 1
    dof_of_node = np.full(int(gd.max()) + 1, -1, dtype=np.int64)
    dof_of_node[gd] = vd
    x = np.empty((n, 2))
    x[dof_of_node] = d["geometry_x"]
    assert np.abs(x - d["V_dof_coordinates"]).max() < 1e-14
    cells = vd.astype(np.int64)
    ctags = np.zeros(len(cells), dtype=np.int32)
    ctags[d["cell_tag_indices"]] = d["cell_tag_values"]
    fnodes = np.sort(dof_of_node[d["facet_geometry_nodes"]], axis=1)
    ftags_dfx = np.zeros(len(fnodes), dtype=np.int32)
    ftags_dfx[d["facet_tag_indices"]] = d["facet_tag_values"]
    return d, x, cells, ctags, fnodes, ftags_dfx


def _oracle_tags(x, cells, phi):
    pts = OT.cell_detection_points("triangle", 1)
    ftab = np.asarray([OT.coordinate_basis("triangle", p)[0] for p in OT.facet_points_in_cell("triangle", 1)])
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", RuntimeWarning)
        return OT.compute_tags_measures(x, cells, "triangle", phi[cells], OT.point_values_function(phi, cells, ftab),
                                        box_mode=True, detection_points=pts)


def _facet_tags_in_oracle_numbering(out, fnodes, ftags_dfx, n):
    fv = out["facet_vertices"].astype(np.int64)
    key_o, key_d = fv[:, 0] * n + fv[:, 1], fnodes[:, 0] * n + fnodes[:, 1]
    idx = np.searchsorted(key_o, key_d)
    assert np.array_equal(key_o[idx], key_d)
    tags = np.zeros(len(fv), dtype=np.int32)
    tags[idx] = ftags_dfx
    return tags


def _compare_csr(ip, ix, data, b, d, prefix):
    want_ip, want_ix, want = d[prefix + "_indptr"], d[prefix + "_indices"], d[prefix + "_data"]
    assert np.array_equal(ip, want_ip) and np.array_equal(ix, want_ix), "sparsity differs from dolfinx's"
    rows = np.repeat(np.arange(len(ip) - 1), np.diff(ip))
    scale = np.zeros(len(ip) - 1)
    np.maximum.at(scale, rows, np.abs(want))
    assert np.all(np.abs(data - want) <= 1e-12 * scale[rows])
    assert np.abs(b - d[prefix + "_b"]).max() <= 1e-12 * np.abs(d[prefix + "_b"]).max()


@needs_fixture
@pytest.mark.parametrize("path", FIXTURES or [None])
def test_oracle_equals_dolfinx(path):
    _run_fixture_check(path)


def _run_fixture_check(path):
    d, x, cells, ctags, fnodes, ftags_dfx = _load(path)
    out = _oracle_tags(x, cells, d["phi"])
    assert np.array_equal(out["cell_tags"], ctags)
    ftags = _facet_tags_in_oracle_numbering(out, fnodes, ftags_dfx, len(x))
    assert np.array_equal(out["facet_tags"], ftags)
    ents = d["ds100"].reshape(-1, 2)
    assert sorted(map(tuple, ents)) == sorted(map(tuple, np.asarray(out["ds100"]).reshape(-1, 2)))
    ip, ix, data, b = OA.assemble_strong_dirichlet(x, cells, cells, len(x), d["phi"], d["f"], ctags, ftags, out["c2f"],
                                                   out["f2c"], out["ds100"], sigma=float(d["stab_coef"]))
    _compare_csr(ip, ix, data, b, d, "strong")
    n = len(x)
    md = d["M_dofmap"]
    # mixed numbering of the fixture: u at M_sub0_dofs[s], p at M_sub1_dofs[s]; the oracle numbers 2 s / 2 s + 1
    ip, ix, data, b = OA.assemble_weak_dirichlet(x, cells, cells, n, d["phi"], d["f"], d["u_D"], ctags, ftags,
                                                 out["c2f"], out["f2c"], out["ds100"], gamma=float(d["pen_coef"]),
                                                 sigma=float(d["stab_coef"]))
    import scipy.sparse as sp
    perm = np.empty(2 * n, dtype=np.int64)
    perm[0::2], perm[1::2] = d["M_sub0_dofs"], d["M_sub1_dofs"]
    A = sp.csr_matrix((data, ix, ip), shape=(2 * n, 2 * n)).tocoo()
    A = sp.coo_matrix((A.data, (perm[A.row], perm[A.col])), shape=A.shape).tocsr()
    A.sort_indices()
    bb = np.empty(2 * n)
    bb[perm] = b
    assert md.shape[1] == 2 * cells.shape[1]
    _compare_csr(A.indptr, A.indices, A.data, bb, d, "weak")


@needs_fixture
@pytest.mark.gpu
@pytest.mark.parametrize("path", FIXTURES or [None])
def test_cuda_equals_dolfinx(path):
    import torch
    from phifem_b200 import assemble, fem, mesh_scripts
    from phifem_b200.mesh import Mesh
    d, x, cells, ctags, fnodes, ftags_dfx = _load(path)
    mesh = Mesh(x, cells, "triangle", device="cuda")
    fn = fem.Function(fem.functionspace(mesh, 1), d["phi"])
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", RuntimeWarning)
        ct, ft, _, ds, _ = mesh_scripts.compute_tags_measures(mesh, fn, 1, box_mode=True)
    assert np.array_equal(ct.values_dev.cpu().numpy(), ctags)
    out = {"facet_vertices": mesh.facet_vertices.cpu().numpy()}
    assert np.array_equal(ft.values_dev.cpu().numpy(), _facet_tags_in_oracle_numbering(out, fnodes, ftags_dfx, len(x)))
    plan = assemble.build_plan(mesh, ct, ft, ds(100))
    A, b = assemble.assemble_strong_dirichlet(plan, torch.from_numpy(d["phi"]).cuda(), torch.from_numpy(d["f"]).cuda(),
                                              stab_coef=float(d["stab_coef"]))
    _compare_csr(A.indptr.cpu().numpy(), A.indices.cpu().numpy(), A.data.cpu().numpy(), b.cpu().numpy(), d, "strong")


def test_fixture_status_is_reported():
    """Always runs: states in the test log whether the dolfinx pin exists."""
    print("dolfinx fixtures present: %d (%s)" % (len(FIXTURES), ", ".join(os.path.basename(p) for p in FIXTURES) or
                                                 "none: CSR parity pinned by tests/test_oracle_sympy.py only"))


def test_fixture_plumbing_on_a_synthetic_fixture(tmp_path, monkeypatch):
    """The loader / comparison code above, exercised on a fixture fabricated FROM THE ORACLE in the layout
    baseline/dolfinx_reference.py writes (geometry nodes, P1 dofs and facets each in their own random numbering, mixed
    dofs blocked instead of interleaved): when a real dolfinx fixture arrives, a failure is about conventions, not about
    this file's index juggling."""
    import cases
    rng = np.random.default_rng(3)
    x0, cells0, _ = cases.load_mesh_arrays("coarse_square")
    nn = len(x0)
    node_perm = rng.permutation(nn)              # geometry node numbering
    dof_of_node = rng.permutation(nn)            # P1 dof numbering
    gx = np.empty_like(x0[:, :2])
    gx[node_perm] = x0[:, :2]
    gd = node_perm[cells0]
    node_of_old = node_perm
    vd = dof_of_node[gd]
    x = np.empty((nn, 2))
    x[dof_of_node] = gx
    phi = (x[:, 0] - 0.5) ** 2 + (x[:, 1] - 0.5) ** 2 - 0.2
    f = np.sin(3 * x[:, 0]) + 1.0
    ud = 0.25 * x[:, 0] - 0.5 * x[:, 1]
    out = _oracle_tags(x, vd.astype(np.int64), phi)
    nf = len(out["facet_tags"])
    fperm = rng.permutation(nf)                  # dolfinx's own facet numbering: fixture facet k = oracle facet fperm[k]
    node_of_dof = np.empty(nn, dtype=np.int64)
    node_of_dof[dof_of_node] = np.arange(nn)
    fx = dict(geometry_x=gx, geometry_dofmap=gd, V_dofmap=vd, V_dof_coordinates=x, phi=phi, f=f, u_D=ud,
              cell_tag_indices=np.arange(len(gd)), cell_tag_values=out["cell_tags"],
              facet_geometry_nodes=node_of_dof[out["facet_vertices"][fperm]][:, ::-1],
              facet_tag_indices=np.arange(nf), facet_tag_values=out["facet_tags"][fperm],
              ds100=np.asarray(out["ds100"]).reshape(-1, 2)[::-1].ravel(), stab_coef=1.5, pen_coef=2.0)
    ip, ix, data, b = OA.assemble_strong_dirichlet(x, vd, vd, nn, phi, f, out["cell_tags"], out["facet_tags"],
                                                   out["c2f"], out["f2c"], out["ds100"], sigma=1.5)
    fx.update(strong_indptr=ip, strong_indices=ix, strong_data=data, strong_b=b)
    ip, ix, data, b = OA.assemble_weak_dirichlet(x, vd, vd, nn, phi, f, ud, out["cell_tags"], out["facet_tags"],
                                                 out["c2f"], out["f2c"], out["ds100"], gamma=2.0, sigma=1.5)
    import scipy.sparse as sp
    perm = np.empty(2 * nn, dtype=np.int64)      # blocked mixed numbering: u dofs first, then p dofs
    perm[0::2], perm[1::2] = np.arange(nn), nn + np.arange(nn)
    A = sp.csr_matrix((data, ix, ip), shape=(2 * nn, 2 * nn)).tocoo()
    A = sp.coo_matrix((A.data, (perm[A.row], perm[A.col])), shape=A.shape).tocsr()
    A.sort_indices()
    bb = np.empty(2 * nn)
    bb[perm] = b
    fx.update(weak_indptr=A.indptr, weak_indices=A.indices, weak_data=A.data, weak_b=bb,
              M_dofmap=np.concatenate([vd, nn + vd], axis=1), M_sub0_dofs=np.arange(nn), M_sub1_dofs=nn + np.arange(nn))
    del node_of_old
    path = tmp_path / "dolfinx_csr_synthetic.npz"
    np.savez(path, **fx)
    _run_fixture_check(str(path))
