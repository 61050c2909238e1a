"""One rank of the CPU (gloo) multi-process test of phifem_b200.dist: the CUDA kernels are replaced by
the oracle's C functions (test-only), everything else -- partition, symbolic exchange, slot maps into the
send segments, halo exchange, owner-side add -- is the product code."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import native as ON  # noqa: E402
from phifem_b200 import dist as pdist  # noqa: E402


def oracle_kernels(plan, phi, f, sigma, data, b):
    m = plan.mesh
    x, cells = m.x.numpy(), np.ascontiguousarray(m.cells.numpy())
    c2f, f2c = np.ascontiguousarray(m.c2f.numpy()), np.ascontiguousarray(m.f2c.numpy())
    ct = plan.cell_tags8.numpy().astype(np.int32)
    d, bb = ON.assemble_p1(x, cells, c2f, f2c, phi.numpy(), f.numpy(), ct, plan.active.numpy(),
                           plan.slots_cells.numpy(), plan.entities.numpy(), plan.slots_boundary.numpy(),
                           plan.ghost.numpy(), plan.slots_ghost.numpy(), sigma, data.numel())
    data.copy_(torch.from_numpy(d))
    b.copy_(torch.from_numpy(bb))


def main():
    n, out_dir = int(sys.argv[1]), sys.argv[2]
    mode = sys.argv[3] if len(sys.argv) > 3 else "exchange"
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    prob = pdist.SlabProblem(n, rank, world, "cpu", mode=mode)
    m = prob.mesh
    x, cells = m.x.numpy(), np.ascontiguousarray(m.cells.numpy())
    ct = ON.tag_cells_p1(x, cells, prob.phi.numpy())
    ft = ON.tag_facets_p1(x, cells, np.ascontiguousarray(m.c2f.numpy()), np.ascontiguousarray(m.f2c.numpy()),
                          prob.phi.numpy(), ct)
    plan = prob.build_plan(torch.from_numpy(ct.astype(np.int8)), torch.from_numpy(ft.astype(np.int8)))
    data, b = prob.assemble(1.0, local_kernels=oracle_kernels)
    if mode == "rows":     # owner computes: nothing was exchanged; the owned rows are a slice of the local CSR
        indptr, indices, _, _ = prob.owned_csr()
        n_send = n_halo_b = 0
        # every record of the row-gather plan belongs to an owned row
        rp = plan.rowsplan
        owned = (prob.vertex_owner == rank).numpy()
        for rl in (rp.cells, rp.surface):
            assert owned[rl.rows.numpy()].all()
    else:
        indptr, indices = plan.indptr, plan.indices
        n_send = sum(hi - lo for lo, hi in plan.send_ranges)
        n_halo_b = sum(int(v.numel()) for v in plan.b_send_local)
    np.savez(os.path.join(out_dir, "rank%d.npz" % rank), row_lo=prob.row_lo, row_hi=prob.row_hi,
             indptr=indptr.numpy(), indices=indices.numpy(), data=data.numpy(), b=b.numpy(),
             cell_tags=ct[prob.cell_owned.numpy()], n_send=n_send, n_halo_b=n_halo_b,
             n_local_cells=m.num_cells)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
