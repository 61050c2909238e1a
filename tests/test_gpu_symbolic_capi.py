"""The symbolic phase behind the C ABI (csrc/symbolic.cu, phifem_pattern_create_p1) produces bit for bit the arrays of the
torch-based plan (phifem_b200/assemble.py) -- CSR pattern, active cells, ghost-penalty facets, the three slot maps -- and
the operator assembled through it matches the oracle; a host program without Python (examples/capi_host.cu: tags ->
pattern -> assembly through the shared library alone) reproduces the checksums of the Python path."""
import os
import subprocess
import warnings

import numpy as np
import pytest
import torch

from oracle import assembly as OA
from phifem_b200 import assemble, fem, mesh_scripts, symbolic, synthetic

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _problem(kind, n):
    if kind.startswith("tri"):
        mesh = synthetic.rectangle_mesh(n, device="cuda")
        center, radius = (0.013, -0.021), 0.61
    else:
        mesh = synthetic.box_mesh(n, device="cuda")
        center, radius = synthetic.SPHERE_CENTER, 0.37
    if kind.endswith("unstructured"):
        mesh = synthetic.unstructured_variant(mesh, jitter=0.2, seed=7)
    phi = synthetic.sphere_levelset(mesh.x, center=center, radius=radius)
    fn = fem.Function(fem.functionspace(mesh, 1), phi.cpu().numpy())
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", RuntimeWarning)
        ctags, ftags, _, ds, _ = mesh_scripts.compute_tags_measures(mesh, fn, 1, box_mode=True)
    return mesh, phi, ctags, ftags, ds


@pytest.mark.parametrize("kind,n", [("tri", 40), ("tri-unstructured", 24), ("tet", 10), ("tet-unstructured", 8)])
def test_device_pattern_equals_the_torch_plan_and_the_oracle(kind, n):
    mesh, phi, ctags, ftags, ds = _problem(kind, n)
    plan = assemble.build_plan(mesh, ctags, ftags, ds(100), method="atomic")
    pat = symbolic.DevicePattern(mesh, ctags, ftags, ds(100))
    assert pat.nnz == plan.nnz and pat.n_rows == plan.n_rows
    for name in ("indptr", "indices", "active", "ghost", "slots_cells", "slots_ghost", "slots_boundary"):
        assert torch.equal(getattr(pat, name), getattr(plan, name)), name
    f = torch.from_numpy(np.random.default_rng(3).uniform(-1, 1, mesh.num_vertices)).cuda()
    A, b = pat.assemble(phi, f, stab_coef=1.0)
    x = mesh.x.cpu().numpy()
    cells = mesh.cells.cpu().numpy().astype(np.int64)
    ip, ix, data, bo = OA.assemble_strong_dirichlet(
        x, cells, cells, len(x), phi.cpu().numpy(), f.cpu().numpy(), ctags.values_dev.cpu().numpy(),
        ftags.values_dev.cpu().numpy(), mesh.c2f.cpu().numpy(), mesh.f2c.cpu().numpy(),
        ds(100).integration_entities, sigma=1.0)
    assert np.array_equal(A.indptr.cpu().numpy(), ip) and np.array_equal(A.indices.cpu().numpy(), ix)
    scale = np.zeros(len(ip) - 1)
    rows = np.repeat(np.arange(len(ip) - 1), np.diff(ip))
    np.maximum.at(scale, rows, np.abs(data))
    assert np.all(np.abs(A.data.cpu().numpy() - data) <= 1e-12 * scale[rows])
    assert np.all(np.abs(b.cpu().numpy() - bo) <= 1e-12 * np.abs(bo).max())
    del pat


def test_device_pattern_without_active_cells_or_entities():
    mesh = synthetic.rectangle_mesh(6, device="cuda")
    phi = synthetic.sphere_levelset(mesh.x, center=(5.0, 5.0), radius=0.5)      # the disc lies outside the mesh
    fn = fem.Function(fem.functionspace(mesh, 1), phi.cpu().numpy())
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", RuntimeWarning)
        ctags, ftags, _, ds, _ = mesh_scripts.compute_tags_measures(mesh, fn, 1, box_mode=True)
    pat = symbolic.DevicePattern(mesh, ctags, ftags, ds(100))
    assert pat.nnz == 0 and pat.active.numel() == 0 and pat.ghost.numel() == 0
    assert torch.equal(pat.indptr, torch.zeros(mesh.num_vertices + 1, dtype=torch.int32, device="cuda"))
    A, b = pat.assemble(phi, phi)
    assert A.data.numel() == 0 and float(b.abs().max()) == 0.0


def test_c_host_without_python_reproduces_the_python_path(tmp_path):
    """examples/capi_host.cu: a unit-square triangle mesh built in C, tags -> pattern -> assembly through the C ABI."""
    exe = tmp_path / "capi_host"
    lib_dir = os.path.join(ROOT, "phifem_b200")
    subprocess.run(["nvcc", "-O2", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a",
                    os.path.join(ROOT, "examples", "capi_host.cu"), "-I", os.path.join(ROOT, "include"),
                    "-L", lib_dir, "-lphifem_b200", "-Xlinker", "-rpath=" + lib_dir, "-o", str(exe)], check=True)
    n = 48
    out = subprocess.run([str(exe), str(n)], check=True, capture_output=True, text=True).stdout
    got = dict(line.split("=") for line in out.split())
    # the same problem through the Python package
    mesh = synthetic.rectangle_mesh(n, lo=(0.0, 0.0), hi=(1.0, 1.0), device="cuda")
    X = mesh.x
    phi = (X[:, 0] - 0.503) ** 2 + (X[:, 1] - 0.497) ** 2 - 0.3 ** 2
    f = 1.0 + X[:, 0] + 2.0 * X[:, 1]
    fn = fem.Function(fem.functionspace(mesh, 1), phi)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", RuntimeWarning)
        ctags, ftags, _, ds, _ = mesh_scripts.compute_tags_measures(mesh, fn, 1, box_mode=True)
    c8 = ctags.values_dev
    assert [int(got["interior"]), int(got["cut"]), int(got["exterior"])] == \
        [int((c8 == 1).sum()), int((c8 == 2).sum()), int((c8 == 3).sum())]
    plan = assemble.build_plan(mesh, ctags, ftags, ds(100), method="atomic")
    A, b = assemble.assemble_strong_dirichlet(plan, phi, f, stab_coef=1.0)
    assert int(got["nnz"]) == plan.nnz and int(got["n_active"]) == plan.active.numel()
    assert int(got["n_ghost"]) == plan.ghost.numel() and int(got["n_entities"]) == plan.entities.shape[0] > 0
    assert int(got["indices_checksum"]) == int((plan.indices.long() * (torch.arange(plan.nnz, device="cuda") % 7 + 1)).sum())
    assert abs(float(got["data_abs_sum"]) - float(A.data.abs().sum())) <= 1e-11 * float(A.data.abs().sum())
    assert abs(float(got["b_sum"]) - float(b.sum())) <= 1e-11 * float(b.abs().sum())
    # the benchmarked row-gather path from the same C++ host: same pattern, same operator
    assert int(got["rows_same_pattern"]) == 1 and float(got["rows_max_rel_diff"]) <= 1e-12
    rows_plan = assemble.build_plan(mesh, ctags, ftags, ds(100))
    A2, b2 = assemble.assemble_strong_dirichlet(rows_plan, phi, f, stab_coef=1.0)
    assert abs(float(got["rows_data_abs_sum"]) - float(A2.data.abs().sum())) <= 1e-13 * float(A2.data.abs().sum())
    assert abs(float(got["rows_b_sum"]) - float(b2.sum())) <= 1e-12 * float(b2.abs().sum())
