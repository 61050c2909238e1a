"""The two independent restatements of the strong-Dirichlet operator (closed forms vs brute-force
quadrature, oracle/assembly.py) must agree; plus structural properties of the assembled CSR."""
import numpy as np
import pytest

from oracle import assembly as OA
from oracle import tags as OT
from phifem_b200 import synthetic
from phifem_b200.mesh import Mesh


def _random_simplices(d, n, seed):
    rng = np.random.default_rng(seed)
    ref = np.concatenate([np.zeros((1, d)), np.eye(d)], axis=0)
    x = []
    for _ in range(n):
        while True:
            A = rng.normal(size=(d, d))
            if abs(np.linalg.det(A)) > 0.2:
                break
        x.append(ref @ A.T + rng.normal(size=(1, d)))
    x = np.concatenate(x, axis=0)
    cells = np.arange(n * (d + 1)).reshape(n, d + 1)
    return x, cells, rng


@pytest.mark.parametrize("d", [2, 3])
def test_cell_tensors_closed_form_vs_quadrature(d):
    x, cells, rng = _random_simplices(d, 12, 10 + d)
    phi = rng.normal(size=len(x))
    f = rng.normal(size=len(x))
    cut = np.arange(len(cells)) % 2 == 0
    A1, b1 = OA.cell_tensors_closed_form(x, cells, phi, f, cut, 0.7)
    A2, b2 = OA.cell_tensors_quadrature(x, cells, phi[cells], f[cells], cut, 0.7)
    assert np.allclose(A1, A2, rtol=1e-12, atol=1e-12 * np.abs(A2).max())
    assert np.allclose(b1, b2, rtol=1e-12, atol=1e-12 * np.abs(b2).max())
    assert np.allclose(A1, np.swapaxes(A1, 1, 2), rtol=1e-13, atol=1e-13 * np.abs(A1).max())


@pytest.mark.parametrize("d", [2, 3])
def test_boundary_tensors_closed_form_vs_quadrature(d):
    x, cells, rng = _random_simplices(d, 8, 20 + d)
    phi = rng.normal(size=len(x))
    ents = np.array([(c, c % (d + 1)) for c in range(len(cells))])
    A1 = OA.boundary_tensors_closed_form(x, cells, phi, ents)
    A2 = OA.boundary_tensors_quadrature(x, cells, phi[cells[ents[:, 0]]], ents)
    assert np.allclose(A1, A2, rtol=1e-12, atol=1e-12 * np.abs(A2).max())


def _small_mesh(d, n=3, jitter=0.15, seed=3):
    m = synthetic.rectangle_mesh(n, device="cpu") if d == 2 else synthetic.box_mesh(n, device="cpu")
    m = synthetic.unstructured_variant(m, jitter=jitter, seed=seed)
    return m.x.numpy(), m.cells.numpy().astype(np.int64), m.cell_type


@pytest.mark.parametrize("d", [2, 3])
def test_ghost_tensors_closed_form_vs_quadrature(d):
    x, cells, ct = _small_mesh(d)
    c2f, f2c, _ = OT.build_topology(cells, ct)
    rng = np.random.default_rng(5)
    phi = rng.normal(size=len(x))
    facets = np.nonzero(f2c[:, 1] >= 0)[0][:10]
    E1, macro = OA.ghost_tensors_closed_form(x, cells, phi, c2f, f2c, facets, 1.3)
    E2 = OA.ghost_tensors_quadrature(x, cells, phi[cells[f2c[facets, 0]]], phi[cells[f2c[facets, 1]]],
                                     c2f, f2c, facets, 1.3)
    assert np.allclose(E1, E2, rtol=1e-11, atol=1e-12 * np.abs(E2).max())
    assert np.array_equal(macro[:, :d + 1], cells[f2c[facets, 0]])


def _assembled(d, method, n=6):
    m = synthetic.rectangle_mesh(n, device="cpu") if d == 2 else synthetic.box_mesh(n, device="cpu")
    m = synthetic.unstructured_variant(m, jitter=0.1, seed=1)
    x, cells, ct = m.x.numpy(), m.cells.numpy().astype(np.int64), m.cell_type
    center = np.array([0.1, 0.05]) if d == 2 else np.array([0.52, 0.49, 0.51])
    r = 0.6 if d == 2 else 0.33
    phi = ((x - center) ** 2).sum(axis=1) - r * r
    f = np.random.default_rng(1234).uniform(-1, 1, size=len(x))
    pts = OT.cell_detection_points(ct, 1)
    fpts = OT.facet_points_in_cell(ct, 1)
    out = OT.compute_tags_measures(x, cells, ct, phi[cells], OT.point_values_function(
        phi, cells, np.asarray([OT.coordinate_basis(ct, p)[0] for p in fpts])), box_mode=True,
        detection_points=pts)
    ip, ix, data, b = OA.assemble_strong_dirichlet(
        x, cells, cells, len(x), phi, f, out["cell_tags"], out["facet_tags"], out["c2f"], out["f2c"],
        out["ds100"], sigma=1.0, method=method)
    return x, cells, out, ip, ix, data, b


@pytest.mark.parametrize("d", [2, 3])
def test_assembled_operator_methods_agree_and_structure(d):
    x, cells, out, ip, ix, data, b = _assembled(d, "closed_form")
    _, _, _, ip2, ix2, data2, b2 = _assembled(d, "quadrature")
    assert np.array_equal(ip, ip2) and np.array_equal(ix, ix2)
    assert np.allclose(data, data2, rtol=0, atol=1e-12 * np.abs(data2).max())
    assert np.allclose(b, b2, rtol=0, atol=1e-12 * np.abs(b2).max())
    ct = out["cell_tags"]
    assert (ct == 2).any() and (ct == 1).any() and (ct == 3).any()
    # rows of dofs only touched by exterior cells are empty
    touched = np.zeros(len(x), dtype=bool)
    touched[np.unique(cells[ct != 3])] = True
    assert np.all((np.diff(ip) > 0) == touched)
    # columns sorted and unique per row
    for r in range(len(x)):
        row = ix[ip[r]:ip[r + 1]]
        assert np.all(np.diff(row) > 0)
    # without the one-sided boundary term the operator is symmetric
    ip3, ix3, data3, _ = OA.assemble_strong_dirichlet(
        x, cells, cells, len(x), ((x - x.mean(0)) ** 2).sum(1) - 0.1, np.ones(len(x)), ct,
        out["facet_tags"], out["c2f"], out["f2c"], np.zeros(0, dtype=np.int32))
    A = OA.to_scipy(ip3, ix3, data3, len(x))
    assert abs(A - A.T).max() <= 1e-12 * abs(A).max()
