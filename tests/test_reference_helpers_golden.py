"""Rows a1 / a5 / a6 of SURVEY.md section 8 against outputs of the REFERENCE'S OWN CODE
(tests/golden/reference_helpers.npz, made by tests/golden/make_reference_helpers_fixture.py, which runs the
unmodified functions of reference src/phifem/mesh_scripts.py with stubbed dolfinx imports)."""
import os

import numpy as np
import pytest
import torch

import cases
from oracle import tags as OT
from phifem_b200 import _geometry as G
from phifem_b200 import mesh_scripts
from phifem_b200.mesh import Mesh

GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_helpers.npz"))
ENT_KEYS = sorted(k[len("ents100_"):] for k in GOLD.files if k.startswith("ents100_"))


@pytest.mark.parametrize("N", range(5))
def test_detection_points_equal_the_reference_generators(N):
    """mesh_scripts.py:28-92 -- oracle and product (host) generators, bit for bit."""
    assert np.array_equal(OT.segment_points(N), GOLD["segment_%d" % N])
    assert np.array_equal(OT.triangle_boundary_points(N), GOLD["triangle_%d" % N])
    assert np.array_equal(OT.square_boundary_points(N), GOLD["square_%d" % N])
    assert np.array_equal(G.cell_detection_points("triangle", N), GOLD["triangle_%d" % N])
    assert np.array_equal(G.cell_detection_points("quadrilateral", N), GOLD["square_%d" % N])


@pytest.mark.parametrize("name", ["coarse_square", "square_tri", "square_quad", "disk"])
def test_reshape_map_equals_the_reference(name):
    """mesh_scripts.py:195-214: padded map, reverse link order."""
    x, cells, ct = cases.load_mesh_arrays(name)
    mesh = Mesh(x, cells, ct, device="cpu")
    tdim = mesh.topology.dim
    mesh.topology.create_connectivity(tdim - 1, tdim)
    emap, width = mesh_scripts._reshape_map(mesh.topology.connectivity(tdim - 1, tdim))
    assert width == 2 and np.array_equal(emap, GOLD["reshape_f2c_" + name])


@pytest.mark.parametrize("key", ENT_KEYS)
def test_oracle_integration_entities_equal_the_reference(key):
    """mesh_scripts.py:137-192 run by the reference itself on the golden tags vs oracle/tags.py."""
    x, cells, ct = cases.load_mesh_arrays(str(GOLD["mesh_" + key]))
    c2f, f2c, _ = OT.build_topology(cells.astype(np.int64), ct)
    ctags, ftags = GOLD["ctags_" + key], GOLD["ftags_" + key]
    e100 = OT.integration_entities(c2f, f2c, (ctags == 1) | (ctags == 2), ftags == 4)
    e101 = OT.integration_entities(c2f, f2c, (ctags == 2) | (ctags == 3), ftags == 3)
    assert np.array_equal(np.asarray(e100).ravel(), GOLD["ents100_" + key])
    assert np.array_equal(np.asarray(e101).ravel(), GOLD["ents101_" + key])
    assert len(GOLD["ents100_" + key]) > 0


@pytest.mark.gpu
@pytest.mark.parametrize("key", ENT_KEYS)
def test_cuda_integration_entities_equal_the_reference(key):
    """The device entity search (phifem_entity_records + ordering) vs the reference's own output."""
    x, cells, ct = cases.load_mesh_arrays(str(GOLD["mesh_" + key]))
    mesh = Mesh(x, cells, ct, device="cuda")
    ct8 = torch.from_numpy(GOLD["ctags_" + key]).cuda()
    ft8 = torch.from_numpy(GOLD["ftags_" + key]).cuda()
    e100 = mesh_scripts._integration_entities_dev(mesh, ct8, ft8, 4, (1, 2)).cpu().numpy()
    e101 = mesh_scripts._integration_entities_dev(mesh, ct8, ft8, 3, (2, 3)).cpu().numpy()
    assert np.array_equal(e100, GOLD["ents100_" + key]) and np.array_equal(e101, GOLD["ents101_" + key])
