#!/usr/bin/env python
"""The LITERAL reference path, for whoever has a dolfinx 0.9.0 environment (none exists in this image: dolfinx,
basix, ufl, ffcx, petsc4py, mpi4py are not installed and there is no network -- BASELINE.md section 3).

    # e.g. inside ghcr.io/fenics/dolfinx/dolfinx:v0.9.0 with the reference checked out beside this repository
    pip install /path/to/phiFEM            # or: export PYTHONPATH=/path/to/phiFEM/src
    python baseline/dolfinx_reference.py fixtures          # writes tests/golden/dolfinx_csr_*.npz
    python baseline/dolfinx_reference.py time --n 1414     # times tags + assembly on the synthetic meshes

`fixtures` runs, on the meshes of tests/golden/meshes.npz (the reference's own test meshes), exactly what the demos
run -- `compute_tags_measures` (reference src/phifem/mesh_scripts.py:571) and the UFL forms of
demo/strong-dirichlet/flower/main.py:92-131 resp. demo/weak-dirichlet/flower/main.py:102-154, assembled with
`dolfinx.fem.petsc.assemble_matrix / assemble_vector` -- and dumps every array a parity check needs: the mesh as dolfinx
numbered it, the dofmaps, the coefficient vectors, the tags, the ds(100) entities, `A.getValuesCSR()` and b.
tests/test_dolfinx_fixture.py consumes these files when present (oracle on CPU, CUDA path on the GPU); until
somebody commits them, parity of the assembled operator rests on the sympy derivation (tests/test_oracle_sympy.py).

`time` is the CPU baseline BASELINE.md asks for: the same calls on the synthetic configurations of bench.py
(dolfinx `create_rectangle` / `create_box` meshes), JIT excluded, printed as one JSON line per configuration.

This file is NOT imported by anything in the repository and has never been executed here.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")

# (fixture name, mesh of tests/golden/meshes.npz, level set) -- the reference's own test data
# (tests/test_compute_meshtags.py:21-104)
CASES = [
    ("circle_in_square", "square_tri", lambda x: x[0] ** 2 + x[1] ** 2 - 0.125),
    ("circle_near_boundary", "coarse_square", lambda x: (x[0] - 0.5) ** 2 + (x[1] - 0.5) ** 2 - 0.2),
    ("circle_in_circle", "disk", lambda x: x[0] ** 2 + x[1] ** 2 - 0.125),
]


def source_term(x):
    return np.sin(3.0 * x[0]) * np.cos(2.0 * x[1]) + 1.0


def dirichlet_data(x):
    return 0.25 * x[0] - 0.5 * x[1]


def _imports():
    from mpi4py import MPI
    import basix.ufl
    import dolfinx as dfx
    import ufl
    from dolfinx.fem.petsc import assemble_matrix, assemble_vector
    from phifem.mesh_scripts import compute_tags_measures
    return MPI, basix, dfx, ufl, assemble_matrix, assemble_vector, compute_tags_measures


def make_mesh(x, cells, cell_name):
    MPI, basix, dfx, ufl, *_ = _imports()
    gdim = x.shape[1]
    domain = ufl.Mesh(basix.ufl.element("Lagrange", cell_name, 1, shape=(gdim,)))
    return dfx.mesh.create_mesh(MPI.COMM_SELF, np.ascontiguousarray(cells, dtype=np.int64),
                                np.ascontiguousarray(x, dtype=np.float64), domain)


def strong_forms(mesh, cells_tags, facets_tags, ds, phi_h, f_h, V, stab_coef):
    """demo/strong-dirichlet/flower/main.py:92-128, verbatim."""
    _, _, dfx, ufl, *_ = _imports()
    w = ufl.TrialFunction(V)
    phiw = phi_h * w
    v = ufl.TestFunction(V)
    phiv = phi_h * v
    dx = ufl.Measure("dx", domain=mesh, subdomain_data=cells_tags)
    dS = ufl.Measure("dS", domain=mesh, subdomain_data=facets_tags)
    h_T = ufl.CellDiameter(mesh)
    n = ufl.FacetNormal(mesh)
    a = (ufl.inner(ufl.grad(phiw), ufl.grad(phiv)) * dx((1, 2))
         - ufl.inner(ufl.inner(ufl.grad(phiw), n), phiv) * ds
         + stab_coef * h_T ** 2 * ufl.inner(ufl.div(ufl.grad(phiw)), ufl.div(ufl.grad(phiv))) * dx(2)
         + stab_coef * ufl.avg(h_T) * ufl.inner(ufl.jump(ufl.grad(phiw), n), ufl.jump(ufl.grad(phiv), n)) * dS((2, 3)))
    L = ufl.inner(f_h, phiv) * dx((1, 2)) - stab_coef * h_T ** 2 * ufl.inner(f_h, ufl.div(ufl.grad(phiv))) * dx(2)
    return dfx.fem.form(a), dfx.fem.form(L)


def weak_forms(mesh, cells_tags, facets_tags, ds, phi_h, f_h, u_D, M, pen_coef, stab_coef):
    """demo/weak-dirichlet/flower/main.py:102-151, verbatim."""
    _, _, dfx, ufl, *_ = _imports()
    u, p = ufl.TrialFunctions(M)
    v, q = ufl.TestFunctions(M)
    dx = ufl.Measure("dx", domain=mesh, subdomain_data=cells_tags)
    dS = ufl.Measure("dS", domain=mesh, subdomain_data=facets_tags)
    h_T = ufl.CellDiameter(mesh)
    n = ufl.FacetNormal(mesh)
    a = (ufl.inner(ufl.grad(u), ufl.grad(v)) * dx((1, 2))
         - ufl.inner(ufl.inner(ufl.grad(u), n), v) * ds
         + pen_coef * h_T ** (-2) * ufl.inner(u - h_T ** (-1) * ufl.inner(phi_h, p),
                                              v - h_T ** (-1) * ufl.inner(phi_h, q)) * dx(2)
         + stab_coef * h_T ** 2 * ufl.inner(ufl.div(ufl.grad(u)), ufl.div(ufl.grad(v))) * dx(2)
         + stab_coef * ufl.avg(h_T) * ufl.inner(ufl.jump(ufl.grad(u), n), ufl.jump(ufl.grad(v), n)) * dS((2, 3)))
    L = (ufl.inner(f_h, v) * dx((1, 2))
         + pen_coef * h_T ** (-2) * ufl.inner(u_D, v - h_T ** (-1) * ufl.inner(phi_h, q)) * dx(2)
         - stab_coef * h_T ** 2 * ufl.inner(f_h, ufl.div(ufl.grad(v))) * dx(2))
    return dfx.fem.form(a), dfx.fem.form(L)


def _csr(A):
    indptr, indices, data = A.getValuesCSR()
    return np.asarray(indptr, dtype=np.int64), np.asarray(indices, dtype=np.int64), np.asarray(data, dtype=np.float64)


def run_case(name, mesh_name, levelset, degree=1):
    MPI, basix, dfx, ufl, assemble_matrix, assemble_vector, compute_tags_measures = _imports()
    meshes = np.load(os.path.join(GOLDEN, "meshes.npz"))
    x, cells = meshes[mesh_name + "_x"], meshes[mesh_name + "_cells"]
    mesh = make_mesh(x[:, :2], cells, "triangle")
    tdim = mesh.topology.dim
    cell_name = mesh.topology.cell_name()
    V = dfx.fem.functionspace(mesh, basix.ufl.element("Lagrange", cell_name, degree))
    phi_h = dfx.fem.Function(V)
    phi_h.interpolate(levelset)
    f_h = dfx.fem.Function(V)
    f_h.interpolate(source_term)
    cells_tags, facets_tags, _, ds_bdy, _ = compute_tags_measures(mesh, phi_h, 1, box_mode=True)
    ds = ds_bdy(100)
    out = {}
    # the mesh as dolfinx numbered it
    mesh.topology.create_connectivity(tdim - 1, 0)
    mesh.topology.create_connectivity(tdim, tdim - 1)
    nf = mesh.topology.index_map(tdim - 1).size_local
    out["geometry_x"] = mesh.geometry.x[:, :2].copy()
    out["geometry_dofmap"] = np.asarray(mesh.geometry.dofmap, dtype=np.int64)
    out["original_cell_index"] = np.asarray(mesh.topology.original_cell_index, dtype=np.int64)
    out["input_global_indices"] = np.asarray(mesh.geometry.input_global_indices, dtype=np.int64)
    out["facet_geometry_nodes"] = np.asarray(
        dfx.mesh.entities_to_geometry(mesh, tdim - 1, np.arange(nf, dtype=np.int32)), dtype=np.int64)
    out["cell_facets"] = mesh.topology.connectivity(tdim, tdim - 1).array.reshape(-1, tdim + 1).astype(np.int64)
    out["V_dofmap"] = np.asarray(V.dofmap.list, dtype=np.int64).reshape(out["geometry_dofmap"].shape[0], -1)
    out["V_dof_coordinates"] = V.tabulate_dof_coordinates()[:, :2].copy()
    out["phi"], out["f"] = phi_h.x.array.copy(), f_h.x.array.copy()
    out["cell_tag_indices"], out["cell_tag_values"] = cells_tags.indices.copy(), cells_tags.values.copy()
    out["facet_tag_indices"], out["facet_tag_values"] = facets_tags.indices.copy(), facets_tags.values.copy()
    sd = ds_bdy.subdomain_data()
    for sid, ents in sd:
        out["ds%d" % sid] = np.asarray(ents, dtype=np.int64)
    # strong-Dirichlet operator
    stab = 1.0
    a, L = strong_forms(mesh, cells_tags, facets_tags, ds, phi_h, f_h, V, stab)
    A = assemble_matrix(a)
    A.assemble()
    b = assemble_vector(L)
    out["strong_indptr"], out["strong_indices"], out["strong_data"] = _csr(A)
    out["strong_b"] = b.array.copy()
    out["stab_coef"] = stab
    # weak-Dirichlet operator
    el = basix.ufl.element("Lagrange", cell_name, degree)
    M = dfx.fem.functionspace(mesh, basix.ufl.mixed_element([el, el]))
    u_D = dfx.fem.Function(V)
    u_D.interpolate(dirichlet_data)
    pen = 1.0
    a, L = weak_forms(mesh, cells_tags, facets_tags, ds, phi_h, f_h, u_D, M, pen, stab)
    A = assemble_matrix(a)
    A.assemble()
    b = assemble_vector(L)
    out["weak_indptr"], out["weak_indices"], out["weak_data"] = _csr(A)
    out["weak_b"] = b.array.copy()
    out["u_D"] = u_D.x.array.copy()
    out["pen_coef"] = pen
    out["M_dofmap"] = np.asarray(M.dofmap.list, dtype=np.int64).reshape(out["geometry_dofmap"].shape[0], -1)
    out["M_sub0_dofs"] = np.asarray(M.sub(0).collapse()[1], dtype=np.int64)     # mixed dof of each V dof (u)
    out["M_sub1_dofs"] = np.asarray(M.sub(1).collapse()[1], dtype=np.int64)     # ... (p)
    out["versions"] = np.array([dfx.__version__, ufl.__version__, basix.__version__])
    path = os.path.join(GOLDEN, "dolfinx_csr_%s.npz" % name)
    np.savez_compressed(path, **out)
    print("wrote", path, "cells", out["geometry_dofmap"].shape[0], "nnz", len(out["strong_data"]))


def run_time(n, dim):
    """Tags + assembly of the strong-Dirichlet operator on a synthetic mesh of bench.py's shape, JIT excluded."""
    MPI, basix, dfx, ufl, assemble_matrix, assemble_vector, compute_tags_measures = _imports()
    if dim == 2:
        mesh = dfx.mesh.create_rectangle(MPI.COMM_WORLD, [np.array([-1.0, -1.0]), np.array([1.0, 1.0])], [n, n])
        c, r = (np.pi / 1000.0, np.e / 1000.0), 0.6
    else:
        mesh = dfx.mesh.create_box(MPI.COMM_WORLD, [np.zeros(3), np.ones(3)], [n, n, n])
        c, r = (0.5 + np.pi / 1000.0, 0.5 + np.e / 1000.0, 0.5 + np.sqrt(2.0) / 1000.0), 0.45
    V = dfx.fem.functionspace(mesh, ("Lagrange", 1))
    phi_h = dfx.fem.Function(V)
    phi_h.interpolate(lambda x: sum((x[k] - c[k]) ** 2 for k in range(dim)) - r * r)
    f_h = dfx.fem.Function(V)
    f_h.interpolate(lambda x: np.where(sum((x[k] - c[k]) ** 2 for k in range(dim)) <= 0.04, 10.0, 0.0))
    times = {}
    for rep in range(3):            # rep 0 pays the FFCx JIT
        t0 = time.perf_counter()
        cells_tags, facets_tags, _, ds_bdy, _ = compute_tags_measures(mesh, phi_h, 1, box_mode=True)
        t1 = time.perf_counter()
        a, L = strong_forms(mesh, cells_tags, facets_tags, ds_bdy(100), phi_h, f_h, V, 1.0)
        t2 = time.perf_counter()
        A = assemble_matrix(a)
        A.assemble()
        b = assemble_vector(L)
        t3 = time.perf_counter()
        times = {"tags_s": t1 - t0, "form_s": t2 - t1, "assembly_s": t3 - t2}
    nc = mesh.topology.index_map(mesh.topology.dim).size_global
    print(json.dumps({"impl": "dolfinx", "dim": dim, "n": n, "cells": int(nc), "ranks": MPI.COMM_WORLD.size, **times,
                      "cells_per_s": nc / (times["tags_s"] + times["assembly_s"])}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("mode", choices=["fixtures", "time"])
    ap.add_argument("--n", type=int, default=200)
    ap.add_argument("--dim", type=int, default=2, choices=[2, 3])
    args = ap.parse_args()
    try:
        _imports()
    except ImportError as exc:
        sys.exit("baseline/dolfinx_reference.py needs dolfinx 0.9.0 + the reference's phifem package: %s" % exc)
    if args.mode == "fixtures":
        for name, mesh_name, ls in CASES:
            run_case(name, mesh_name, ls)
    else:
        run_time(args.n, args.dim)


if __name__ == "__main__":
    main()
