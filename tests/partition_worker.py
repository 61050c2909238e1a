"""One rank of the CPU (gloo) test of phifem_b200.partition: the CUDA kernels are replaced by the oracle's C
functions (test-only); the Morton partition, ownership, local mesh, row-gather plan restricted to the owned
rows and the owned-row extraction are the product code."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import native as ON  # noqa: E402
from phifem_b200 import dist as pdist  # noqa: E402
from phifem_b200 import partition, synthetic  # noqa: E402
from dist_worker import oracle_kernels  # noqa: E402


def global_problem(kind, n, single=False):
    mesh = synthetic.rectangle_mesh(n, device="cpu") if kind == "tri" else synthetic.box_mesh(n, device="cpu")
    mesh = synthetic.unstructured_variant(mesh, jitter=0.2, seed=11)
    center = (0.013, -0.021) if kind == "tri" else synthetic.SPHERE_CENTER
    phi = synthetic.sphere_levelset(mesh.x, center=center, radius=0.61 if kind == "tri" else 0.37)
    if single:   # add a ball around ONE vertex far from the main body: the cells touching that vertex are cut and
        # have no interior neighbour, so `single_layer_cut` must re-tag them
        h = (2.0 if kind == "tri" else 1.0) / n
        far = torch.argmax(torch.where(phi > 4 * h, phi, torch.full_like(phi, -1.0)))
        c2 = tuple(float(v) for v in mesh.x[far])
        phi = torch.minimum(phi, synthetic.sphere_levelset(mesh.x, center=c2, radius=0.3 * h))
    f = torch.from_numpy(np.random.default_rng(99).uniform(-1, 1, mesh.num_vertices))
    return mesh, phi, f


def main():
    kind, n, out_dir = sys.argv[1], int(sys.argv[2]), sys.argv[3]
    single = len(sys.argv) > 4 and sys.argv[4] == "single"
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    if os.environ.get("PHIFEM_SCATTER") == "1":
        # only rank 0 ever builds the global problem; the others receive their share
        gmesh, phi, f = global_problem(kind, n, single) if rank == 0 else (None, None, None)
        prob = partition.PartitionedProblem.scatter(gmesh, phi, f, rank, world, single_layer_cut=single, device="cpu")
    else:
        gmesh, phi, f = global_problem(kind, n, single)
        prob = partition.PartitionedProblem(gmesh, phi, f, rank, world, single_layer_cut=single)
    m = prob.mesh
    x, cells = m.x.numpy(), np.ascontiguousarray(m.cells.numpy())
    c2f, f2c = np.ascontiguousarray(m.c2f.numpy()), np.ascontiguousarray(m.f2c.numpy())
    ct = ON.tag_cells_p1(x, cells, prob.phi.numpy())
    if single:   # reference :349-358 on the local mesh (numpy oracle; the C port has no single-layer pass)
        from oracle import tags as OT
        pts = OT.cell_detection_points(m.cell_type, 1)
        ct = OT.tag_cells(prob.phi.numpy()[cells], OT.cell_scale(x, cells.astype(np.int64), m.cell_type, pts),
                          cells.astype(np.int64), single_layer_cut=True, warn=False)
    # "any exterior cell" is global (mesh_scripts.py:469-474): the oracle's facet pass takes it from the local
    # cell tags, so make sure every rank agrees (true here: the disc / sphere leaves exterior cells everywhere)
    flag = torch.tensor([int((ct == 3).any())])
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    assert int(flag) == 1
    ft = ON.tag_facets_p1(x, cells, c2f, f2c, prob.phi.numpy(), ct)
    ct8, ft8 = torch.from_numpy(ct.astype(np.int8)), torch.from_numpy(ft.astype(np.int8))
    prob.build_plan(ct8, ft8, entities=pdist.entities_host(m, ct8, ft8))
    prob.assemble(1.0, local_kernels=oracle_kernels)
    rows, indptr, cols, data, b = prob.owned_csr()
    rp = prob.plan.rowsplan
    owned = prob.row_mask.numpy()
    for rl in (rp.cells, rp.surface):          # every record belongs to an owned row
        assert owned[rl.rows.numpy()].all()
    np.savez(os.path.join(out_dir, "rank%d.npz" % rank), rows=rows.numpy(), indptr=indptr.numpy(),
             cols=cols.numpy(), data=data.numpy(), b=b.numpy(),
             owned_cells=prob.global_cell[prob.cell_owned].numpy(), cell_tags=ct[prob.cell_owned.numpy()],
             n_local_cells=m.num_cells)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
