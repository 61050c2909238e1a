"""Multi-process path on CPU: world_size 2 over gloo.  The ranks' owned CSR rows, stacked, must equal the
single-domain operator of the whole box ("8-GPU CSR (gathered) == 1-GPU CSR")."""
import os
import socket
import subprocess
import sys

import numpy as np
import scipy.sparse as sp

from oracle import assembly as OA
from oracle import tags as OT
from phifem_b200 import dist as pdist

HERE = os.path.dirname(os.path.abspath(__file__))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


import pytest


@pytest.mark.parametrize("mode", ["exchange", "rows"])
def test_two_ranks_reproduce_the_single_domain_operator(tmp_path, mode):
    n, world = 4, 2
    port = _free_port()
    procs = []
    for rank in range(world):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1",
                   MASTER_PORT=str(port), OMP_NUM_THREADS="2")
        procs.append(subprocess.Popen([sys.executable, os.path.join(HERE, "dist_worker.py"), str(n),
                                       str(tmp_path), mode], env=env))
    for p in procs:
        assert p.wait(timeout=300) == 0
    mesh, phi, f = pdist.SlabProblem.global_reference(n, world)
    x, cells = mesh.x.numpy(), mesh.cells.numpy().astype(np.int64)
    ph, fh = phi.numpy(), f.numpy()
    pts = OT.cell_detection_points("tetrahedron", 1)
    ftab = np.asarray([OT.coordinate_basis("tetrahedron", p)[0]
                       for p in OT.facet_points_in_cell("tetrahedron", 1)])
    out = OT.compute_tags_measures(x, cells, "tetrahedron", ph[cells], OT.point_values_function(ph, cells, ftab),
                                   box_mode=True, detection_points=pts)
    ip, ix, data, b = OA.assemble_strong_dirichlet(x, cells, cells, len(x), ph, fh, out["cell_tags"],
                                                   out["facet_tags"], out["c2f"], out["f2c"], out["ds100"])
    want = sp.csr_matrix((data, ix, ip), shape=(len(x), len(x)))
    row, tags, halo = 0, [], 0
    for rank in range(world):
        r = np.load(os.path.join(tmp_path, "rank%d.npz" % rank))
        assert int(r["row_lo"]) == row
        row = int(r["row_hi"])
        got = sp.csr_matrix((r["data"], r["indices"], r["indptr"]), shape=(row - int(r["row_lo"]), len(x)))
        ref = want[int(r["row_lo"]):row]
        assert np.array_equal(got.indptr, ref.indptr) and np.array_equal(got.indices, ref.indices)
        assert np.abs(got.data - ref.data).max() <= 1e-12 * np.abs(ref.data).max()
        assert np.abs(r["b"] - b[int(r["row_lo"]):row]).max() <= 1e-12 * np.abs(b).max()
        tags.append(r["cell_tags"])
        halo += int(r["n_send"]) + int(r["n_halo_b"])
    assert row == len(x)
    assert np.array_equal(np.concatenate(tags), out["cell_tags"])
    if mode == "exchange":   # the partition boundary cuts through active cells: a real exchange happened
        assert halo > 0
    else:                    # owner computes: nothing exchanged, rank 0 carries two ghost layers instead
        assert halo == 0
