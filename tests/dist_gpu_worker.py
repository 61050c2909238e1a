"""torchrun worker of tests/test_gpu_dist.py: N ranks assemble their slabs with the CUDA kernels and the
NCCL halo exchange; rank 0 checks the stacked result against the single-GPU product path on the whole box."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from phifem_b200 import assemble, fem, mesh_scripts  # noqa: E402
from phifem_b200 import dist as pdist  # noqa: E402
from phifem_b200.mesh import MeshTags  # noqa: E402


def main():
    n = int(sys.argv[1])
    mode = sys.argv[2] if len(sys.argv) > 2 else "exchange"
    peer_halo = mode == "exchange-peer"      # the same numeric path, halo over NVLink peer memory (csrc/peer.cu)
    mode = "exchange" if peer_halo else mode
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=dev)
    prob = pdist.SlabProblem(n, rank, world, dev, mode=mode)
    dls = mesh_scripts._DeviceLevelset(prob.mesh, fem.Function(fem.functionspace_p1_device(prob.mesh), prob.phi), 1)
    ws = prob.classify(dls, mesh_scripts.TagWorkspace(prob.mesh))
    plan = prob.build_plan(ws.cell_tags8, ws.facet_tags8)
    if peer_halo:
        assert prob.enable_peer_halo(), "peer mapping of the halo buffers failed"
    data, b = prob.assemble(1.0)
    torch.cuda.synchronize()
    if peer_halo:   # two more exchanges: both parities of the receive buffers, a reused epoch slot
        for _ in range(2):
            data, b = prob.assemble(1.0)
        torch.cuda.synchronize()
        assert not prob._halo.a.timed_out() and not prob._halo.b.timed_out()
    if mode == "rows":
        indptr, indices, _, _ = prob.owned_csr()
    else:
        indptr, indices = plan.indptr, plan.indices
    mine = dict(row_lo=prob.row_lo, row_hi=prob.row_hi, indptr=indptr.cpu().numpy(),
                indices=indices.cpu().numpy(), data=data.cpu().numpy(), b=b.cpu().numpy(),
                tags=ws.cell_tags[prob.cell_owned].cpu().numpy(),
                sent=sum(hi - lo for lo, hi in plan.send_ranges))
    parts = [None] * world
    dist.gather_object(mine, parts if rank == 0 else None, dst=0)
    if rank == 0:
        mesh, phi, f = pdist.SlabProblem.global_reference(n, world, dev)
        fn = fem.Function(fem.functionspace_p1_device(mesh), phi)
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore", RuntimeWarning)
            ct, ft, _, ds, _ = mesh_scripts.compute_tags_measures(mesh, fn, 1, box_mode=True)
        A, bb = assemble.assemble_strong_dirichlet(assemble.build_plan(mesh, ct, ft, ds(100)), phi, f)
        ip, ix, dd, bb = (A.indptr.cpu().numpy(), A.indices.cpu().numpy(), A.data.cpu().numpy(),
                          bb.cpu().numpy())
        assert np.array_equal(np.concatenate([p["tags"] for p in parts]), ct.values_dev.cpu().numpy())
        row = 0
        for p in parts:
            assert p["row_lo"] == row
            row = p["row_hi"]
            lo, hi = ip[p["row_lo"]], ip[row]
            assert np.array_equal(p["indptr"], ip[p["row_lo"]:row + 1] - lo)
            assert np.array_equal(p["indices"], ix[lo:hi])
            assert np.abs(p["data"] - dd[lo:hi]).max() <= 1e-12 * np.abs(dd).max()
            assert np.abs(p["b"] - bb[p["row_lo"]:row]).max() <= 1e-12 * np.abs(bb).max()
        assert row == mesh.num_vertices
        assert (sum(p["sent"] for p in parts) > 0) == (mode == "exchange")
        print("DIST-OK world=%d mode=%s cells=%d halo_entries=%d"
              % (world, mode + ("-peer" if peer_halo else ""), mesh.num_cells, sum(p["sent"] for p in parts)))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
