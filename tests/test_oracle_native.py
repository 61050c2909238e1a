"""The C restatement (oracle/csrc/phifem_oracle.c) must equal the numpy oracle: bit-exact tags,
assembled values within 1e-13 of the matrix scale."""
import numpy as np
import pytest
import torch

from oracle import assembly as OA
from oracle import native as ON
from oracle import tags as OT
from phifem_b200 import assemble, synthetic
from phifem_b200.mesh import MeshTags


@pytest.mark.parametrize("d", [2, 3])
def test_native_oracle_equals_numpy_oracle(d):
    m = synthetic.rectangle_mesh(12, device="cpu") if d == 2 else synthetic.box_mesh(6, device="cpu")
    m = synthetic.unstructured_variant(m, jitter=0.15, seed=9)
    x, cells = m.x.numpy(), np.ascontiguousarray(m.cells.numpy())
    center = np.array([0.1, 0.05]) if d == 2 else np.array([0.52, 0.49, 0.51])
    r = 0.6 if d == 2 else 0.33
    phi = ((x - center) ** 2).sum(axis=1) - r * r
    rng = np.random.default_rng(3)
    phi[rng.integers(0, len(phi), 5)] = 0.0              # exercise the exact-equality rule
    f = rng.uniform(-1, 1, len(x))
    ct = m.cell_type
    pts = OT.cell_detection_points(ct, 1)
    ftab = np.asarray([OT.coordinate_basis(ct, p)[0] for p in OT.facet_points_in_cell(ct, 1)])
    c64 = cells.astype(np.int64)
    out = OT.compute_tags_measures(x, c64, ct, phi[c64], OT.point_values_function(phi, c64, ftab),
                                   box_mode=True, detection_points=pts)
    ctags = ON.tag_cells_p1(x, cells, phi)
    assert np.array_equal(ctags, out["cell_tags"])
    ftags = ON.tag_facets_p1(x, cells, out["c2f"], out["f2c"], phi, ctags)
    assert np.array_equal(ftags, out["facet_tags"])

    plan = assemble.build_plan(m, MeshTags(m, d, torch.from_numpy(ctags)),
                               MeshTags(m, d - 1, torch.from_numpy(ftags)), out["ds100"])
    data, b = ON.assemble_p1(x, cells, out["c2f"], out["f2c"], phi, f, ctags, plan.active.numpy(),
                             plan.slots_cells.numpy(), plan.entities.numpy(), plan.slots_boundary.numpy(),
                             plan.ghost.numpy(), plan.slots_ghost.numpy(), 1.0, plan.nnz)
    ip, ix, want, wb = OA.assemble_strong_dirichlet(x, c64, c64, len(x), phi, f, ctags, ftags,
                                                    out["c2f"], out["f2c"], out["ds100"], sigma=1.0)
    assert np.array_equal(plan.indptr.numpy(), ip)
    assert np.abs(data - want).max() <= 1e-13 * np.abs(want).max()
    assert np.abs(b - wb).max() <= 1e-13 * np.abs(wb).max()
    assert ON.num_threads() >= 1
