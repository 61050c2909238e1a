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


@pytest.mark.parametrize("key", NAMES)
def test_oracle_submesh_tags_equal_the_reference_transfer(key):
    """`_transfer_tags` (:217-281) run by the reference onto the submesh of Omega_h vs the oracle's submesh route."""
    x, cells, ct, phi, single = _case(key)
    pts = OT.cell_detection_points(ct, 1)
    ftab = np.asarray([OT.coordinate_basis(ct, p)[0] for p in OT.facet_points_in_cell(ct, 1)])
    out = OT.compute_tags_measures(x, cells, ct, phi[cells], OT.point_values_function(phi, cells, ftab),
                                   box_mode=False, single_layer_cut=single, detection_points=pts)
    assert np.array_equal(out["cell_tags"], GOLD["sub_ctags_" + key])
    assert np.array_equal(out["facet_tags"], GOLD["sub_ftags_" + key])


@pytest.mark.gpu
@pytest.mark.parametrize("key", NAMES)
def test_cuda_submesh_route_and_overwrite_equal_the_reference_code(key):
    from phifem_b200.mesh import MeshTags
    x, cells, ct, phi, single = _case(key)
    mesh = Mesh(x, cells, ct, device="cuda")
    fn = fem.Function(fem.functionspace(mesh, 1), phi)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", RuntimeWarning)
        sct, sft, sub, _, maps = mesh_scripts.compute_tags_measures(mesh, fn, 1, box_mode=False,
                                                                   single_layer_cut=single)
    assert np.array_equal(sct.values_dev.cpu().numpy(), GOLD["sub_ctags_" + key])
    assert np.array_equal(sft.values_dev.cpu().numpy(), GOLD["sub_ftags_" + key])
    assert np.array_equal(maps[0], np.nonzero((GOLD["ctags_" + key] == 1) | (GOLD["ctags_" + key] == 2))[0])
    # the reference's explicit helper on the product's objects (:217-281)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", RuntimeWarning)
        pct, pft, _, _, _ = mesh_scripts.compute_tags_measures(mesh, fn, 1, box_mode=True,
                                                               single_layer_cut=single)
    tf = mesh_scripts._transfer_tags(pft, sub, maps[0], source_mesh=mesh)
    assert np.array_equal(tf.values_dev.cpu().numpy(), GOLD["sub_ftags_" + key])
    with pytest.raises(ValueError, match="source_mesh"):
        mesh_scripts._transfer_tags(pft, sub, maps[0])
    # user overlay through the public argument (:606-615, `_overwrite_tags` :561-568)
    oc = MeshTags.from_lists(mesh, mesh.topology.dim, np.arange(0, mesh.num_cells, 5), np.full(
        len(range(0, mesh.num_cells, 5)), 7))
    of = MeshTags.from_lists(mesh, mesh.topology.dim - 1, np.arange(0, mesh.num_facets, 7), np.full(
        len(range(0, mesh.num_facets, 7)), 9))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", RuntimeWarning)
        c2, f2, _, _, _ = mesh_scripts.compute_tags_measures(mesh, fn, 1, box_mode=True, single_layer_cut=single,
                                                             overwrite_tags={"cells": oc, "facets": of})
    assert np.array_equal(c2.indices, GOLD["ow_c_idx_" + key]) and np.array_equal(c2.values, GOLD["ow_c_val_" + key])
    assert np.array_equal(f2.indices, GOLD["ow_f_idx_" + key]) and np.array_equal(f2.values, GOLD["ow_f_val_" + key])
