"""Rows a3 / a4 against the REFERENCE'S OWN `_tag_cells` / `_tag_facets` code run on fixture and synthetic meshes,
tetrahedra included (tests/golden/reference_tagging.npz; see tests/golden/make_reference_tagging_fixture.py for how
the unmodified reference functions were run with only the dolfinx detection assembly stubbed by the oracle)."""
import os
import warnings

import numpy as np
import pytest

from oracle import tags as OT
from phifem_b200 import fem, mesh_scripts
from phifem_b200.mesh import Mesh

GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_tagging.npz"))
NAMES = [str(n) for n in GOLD["names"]]


def _case(key):
    return (GOLD["x_" + key], GOLD["cells_" + key].astype(np.int64), str(GOLD["type_" + key]), GOLD["phi_" + key],
            key.endswith("_single"))


@pytest.mark.parametrize("key", NAMES)
def test_oracle_tags_equal_the_reference_code(key):
    x, cells, ct, phi, single = _case(key)
    pts = OT.cell_detection_points(ct, 1)
    ftab = np.asarray([OT.coordinate_basis(ct, p)[0] for p in OT.facet_points_in_cell(ct, 1)])
    out = OT.compute_tags_measures(x, cells, ct, phi[cells], OT.point_values_function(phi, cells, ftab),
                                   box_mode=True, single_layer_cut=single, detection_points=pts)
    assert np.array_equal(out["cell_tags"], GOLD["ctags_" + key])
    assert np.array_equal(out["facet_tags"], GOLD["ftags_" + key])
    assert len(out["duplicates"]) == 0


@pytest.mark.gpu
@pytest.mark.parametrize("key", NAMES)
def test_cuda_tags_equal_the_reference_code(key):
    x, cells, ct, phi, single = _case(key)
    mesh = Mesh(x, cells, ct, device="cuda")
    fn = fem.Function(fem.functionspace(mesh, 1), phi)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", RuntimeWarning)
        ctags, ftags, _, _, _ = mesh_scripts.compute_tags_measures(mesh, fn, 1, box_mode=True,
                                                                  single_layer_cut=single)
    assert np.array_equal(ctags.values_dev.cpu().numpy(), GOLD["ctags_" + key])
    assert np.array_equal(ftags.values_dev.cpu().numpy(), GOLD["ftags_" + key])
    # the host views the reference's callers read (MeshTags.indices / .values / .find)
    assert np.array_equal(ctags.values, GOLD["ctags_" + key][ctags.indices])
    assert np.array_equal(ftags.find(4), np.nonzero(GOLD["ftags_" + key] == 4)[0])
