"""General (Morton-curve) sharding on CPU over gloo: the ranks' owned rows, merged by global row id, must equal
the single-domain operator on the unstructured (jittered, permuted, relabelled) mesh -- triangles and
tetrahedra, 2 and 3 ranks."""
import os
import subprocess
import sys

import numpy as np
import pytest
import scipy.sparse as sp

from oracle import assembly as OA
from oracle import tags as OT
from test_dist_gloo import _free_port

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)


@pytest.mark.parametrize("kind,n,world,single,scatter", [
    ("tri", 12, 2, False, False), ("tet", 5, 2, False, False), ("tet", 5, 3, False, True), ("tri", 12, 3, True, True),
    ("tet", 8, 2, True, False)])
def test_partitioned_ranks_reproduce_the_single_domain_operator(tmp_path, kind, n, world, single, scatter):
    """scatter=True: rank 0 alone holds the global mesh, computes the partition once and sends the shares
    (`PartitionedProblem.scatter`); otherwise every rank cuts its share from the global arrays."""
    port = _free_port()
    procs = []
    for rank in range(world):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1",
                   MASTER_PORT=str(port), OMP_NUM_THREADS="2", PHIFEM_SCATTER="1" if scatter else "0")
        procs.append(subprocess.Popen([sys.executable, os.path.join(HERE, "partition_worker.py"), kind, str(n),
                                       str(tmp_path)] + (["single"] if single else []), env=env))
    for p in procs:
        assert p.wait(timeout=300) == 0
    from partition_worker import global_problem
    mesh, phi, f = global_problem(kind, n, single)
    x, cells = mesh.x.numpy(), mesh.cells.numpy().astype(np.int64)
    ph, fh = phi.numpy(), f.numpy()
    ct = mesh.cell_type
    pts = OT.cell_detection_points(ct, 1)
    ftab = np.asarray([OT.coordinate_basis(ct, p)[0] for p in OT.facet_points_in_cell(ct, 1)])
    out = OT.compute_tags_measures(x, cells, ct, ph[cells], OT.point_values_function(ph, cells, ftab),
                                   box_mode=True, single_layer_cut=single, detection_points=pts)
    if single:       # the small ball's cut cells were re-tagged: the case exercises :349-358
        plain = OT.compute_tags_measures(x, cells, ct, ph[cells], OT.point_values_function(ph, cells, ftab),
                                         box_mode=True, detection_points=pts)
        assert (plain["cell_tags"] != out["cell_tags"]).sum() > 0
    ip, ix, data, b = OA.assemble_strong_dirichlet(x, cells, cells, len(x), ph, fh, out["cell_tags"],
                                                   out["facet_tags"], out["c2f"], out["f2c"], out["ds100"])
    want = sp.csr_matrix((data, ix, ip), shape=(len(x), len(x)))
    seen_rows = np.zeros(len(x), dtype=int)
    seen_cells = np.zeros(len(cells), dtype=int)
    sizes = []
    for rank in range(world):
        r = np.load(os.path.join(tmp_path, "rank%d.npz" % rank))
        rows = r["rows"]
        seen_rows[rows] += 1
        seen_cells[r["owned_cells"]] += 1
        assert np.array_equal(r["cell_tags"], out["cell_tags"][r["owned_cells"]])
        got = sp.csr_matrix((r["data"], r["cols"], r["indptr"]), shape=(len(rows), len(x)))
        ref = want[rows]
        assert np.array_equal(got.indptr, ref.indptr) and np.array_equal(got.indices, ref.indices)
        assert np.abs(got.data - ref.data).max() <= 1e-12 * np.abs(ref.data).max()
        assert np.abs(r["b"] - b[rows]).max() <= 1e-12 * np.abs(b).max()
        sizes.append((len(r["owned_cells"]), int(r["n_local_cells"])))
    assert np.all(seen_rows == 1) and np.all(seen_cells == 1)        # a partition of rows and of cells
    counts = [s[0] for s in sizes]
    assert max(counts) - min(counts) <= 1                             # equal ranges of the Morton order
    assert all(s[1] < len(cells) for s in sizes)                      # nobody holds the whole mesh
