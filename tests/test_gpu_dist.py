"""Multi-GPU parity (needs >= 2 GPUs: run with `gpurun --gpus 2 -- python -m pytest tests -m gpu`):
the gathered N-GPU tags and CSR operator equal the single-GPU ones."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
@pytest.mark.parametrize("mode", ["rows", "exchange", "exchange-peer"])
def test_two_gpus_match_single_gpu(mode):
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29611",
                          os.path.join(HERE, "dist_gpu_worker.py"), "12", mode], capture_output=True, text=True,
                         timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-4000:]
    assert "DIST-OK world=2 mode=%s" % mode in out.stdout


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
@pytest.mark.parametrize("kind,n,single", [("tri", 48, False), ("tet", 12, False), ("tet", 12, True)])
def test_partitioned_gpus_are_bitwise_identical_to_single_gpu(kind, n, single):
    """General Morton-curve sharding of an unstructured mesh (phifem_b200/partition.py), owner computes."""
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29613",
                          os.path.join(HERE, "partition_gpu_worker.py"), kind, str(n)]
                         + (["single"] if single else []), capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-4000:]
    assert "PARTITION-OK world=2 kind=%s single=%s" % (kind, single) in out.stdout


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
@pytest.mark.parametrize("peer", [False, True])
def test_global_exterior_flag_reaches_a_rank_without_exterior_cells(peer):
    """One rank holds interior cells only while the mesh has exterior cells: its mesh-boundary facets (and with them
    ds(100) and the operator) are right only if the "any exterior cell" flag of reference :469-474 is exchanged --
    through the all-reduce (peer=False) or through NVLink peer memory (csrc/peer.cu, three epochs)."""
    env = dict(os.environ, PHIFEM_PEER="1" if peer else "0")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29615",
                          os.path.join(HERE, "partition_gpu_worker.py"), "tet", "10", "halfspace"],
                         capture_output=True, text=True, timeout=600, env=env)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-4000:]
    assert "PARTITION-OK world=2 kind=tet single=False" in out.stdout and "peer=%s halfspace=True" % peer in out.stdout
