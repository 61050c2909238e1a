"""torchrun worker of tests/test_gpu_dist.py::test_partitioned_gpus_*: N ranks shard an unstructured mesh along
the Morton curve (phifem_b200/partition.py), classify and assemble their rows with the CUDA kernels; rank 0
checks that the merged rows are BITWISE identical to the single-GPU product path."""
import os
import sys
import warnings

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from phifem_b200 import assemble, fem, mesh_scripts, partition, synthetic  # noqa: E402


def main():
    kind, n = sys.argv[1], int(sys.argv[2])
    single = len(sys.argv) > 3 and sys.argv[3] == "single"
    halfspace = len(sys.argv) > 3 and sys.argv[3] == "halfspace"
    use_peer = os.environ.get("PHIFEM_PEER") == "1"
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    gmesh = synthetic.rectangle_mesh(n, device=dev) if kind == "tri" else synthetic.box_mesh(n, device=dev)
    gmesh = synthetic.unstructured_variant(gmesh, jitter=0.2, seed=11)
    center = (0.013, -0.021) if kind == "tri" else synthetic.SPHERE_CENTER
    phi = synthetic.sphere_levelset(gmesh.x, center=center, radius=0.61 if kind == "tri" else 0.37)
    if single:   # a ball around one far vertex: cut cells without interior neighbour (reference :349-358)
        h = (2.0 if kind == "tri" else 1.0) / n
        far = torch.argmax(torch.where(phi > 4 * h, phi, torch.full_like(phi, -1.0)))
        phi = torch.minimum(phi, synthetic.sphere_levelset(gmesh.x, center=tuple(float(v) for v in gmesh.x[far]),
                                                           radius=0.3 * h))
    if halfspace:
        # Omega = {x_last < 0.6}: the first half of the Morton curve (x_last < 0.5) holds interior cells only, so that
        # rank has NO exterior cell of its own while the mesh has: its mesh-boundary facets are tagged 1 only if the
        # global flag of reference :469-474 reaches it (tag 4 = Gamma_h otherwise, which changes ds(100) and the operator)
        last = gmesh.x[:, -1]
        lo, hi = float(last.min()), float(last.max())
        phi = ((last - lo) / (hi - lo) - 0.6 - 1e-3 * torch.sin(7.0 * gmesh.x[:, 0])).contiguous()
    f = torch.from_numpy(np.random.default_rng(99).uniform(-1, 1, gmesh.num_vertices)).to(dev)
    prob = partition.PartitionedProblem(gmesh, phi, f, rank, world, single_layer_cut=single)
    if use_peer:
        assert prob.enable_peer_flags(), "peer mapping unavailable"
    dls = mesh_scripts._DeviceLevelset(prob.mesh, fem.Function(fem.functionspace_p1_device(prob.mesh), prob.phi), 1)
    ws = mesh_scripts.TagWorkspace(prob.mesh)
    for _ in range(3 if use_peer else 1):        # several epochs: the slots alternate
        prob.classify(dls, ws)
    if halfspace:
        n_ext = int((ws.cell_tags8[prob.cell_owned] == 3).sum())
        flags = [None] * world
        dist.all_gather_object(flags, n_ext)
        assert min(flags) == 0 and max(flags) > 0, "the case needs a rank without exterior cells: %s" % flags
    prob.build_plan(ws.cell_tags8, ws.facet_tags8)
    prob.assemble(1.0)
    torch.cuda.synchronize()
    if use_peer:
        assert not prob.peer.timed_out()
    rows, indptr, cols, data, b = prob.owned_csr()
    mine = dict(rows=rows.cpu().numpy(), indptr=indptr.cpu().numpy(), cols=cols.cpu().numpy(),
                data=data.cpu().numpy(), b=b.cpu().numpy(),
                owned_cells=prob.global_cell[prob.cell_owned].cpu().numpy(),
                tags=ws.cell_tags[prob.cell_owned].cpu().numpy(), n_local=prob.mesh.num_cells)
    parts = [None] * world
    dist.gather_object(mine, parts if rank == 0 else None, dst=0)
    if rank == 0:
        fn = fem.Function(fem.functionspace_p1_device(gmesh), phi)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore", RuntimeWarning)
            ct, ft, _, ds, _ = mesh_scripts.compute_tags_measures(gmesh, fn, 1, box_mode=True,
                                                                  single_layer_cut=single)
            if single:
                ct0 = mesh_scripts.compute_tags_measures(gmesh, fn, 1, box_mode=True)[0]
                assert int((ct0.values_dev != ct.values_dev).sum()) > 0, "single_layer_cut changed nothing"
        A, bb = assemble.assemble_strong_dirichlet(assemble.build_plan(gmesh, ct, ft, ds(100)), phi, f)
        ip, ix, dd, bb = (A.indptr.cpu().numpy(), A.indices.cpu().numpy(), A.data.cpu().numpy(), bb.cpu().numpy())
        tags = ct.values_dev.cpu().numpy()
        seen = np.zeros(gmesh.num_vertices, dtype=int)
        for p in parts:
            seen[p["rows"]] += 1
            assert np.array_equal(p["tags"], tags[p["owned_cells"]])
            cnt = np.diff(p["indptr"])
            assert np.array_equal(cnt, ip[p["rows"] + 1] - ip[p["rows"]])
            slots = np.repeat(ip[p["rows"]] - p["indptr"][:-1], cnt) + np.arange(len(p["cols"]))
            assert np.array_equal(p["cols"], ix[slots])
            assert np.array_equal(p["data"], dd[slots])          # same contributions, same order: bitwise
            assert np.array_equal(p["b"], bb[p["rows"]])
            assert p["n_local"] < gmesh.num_cells
        assert np.all(seen == 1)
        print("PARTITION-OK world=%d kind=%s single=%s cells=%d local=%s peer=%s halfspace=%s"
              % (world, kind, single, gmesh.num_cells, [p["n_local"] for p in parts], use_peer, halfspace))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
