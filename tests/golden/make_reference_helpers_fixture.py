"""Golden vectors produced by the REFERENCE'S OWN CODE for the pure-numpy parts of the path.

reference src/phifem/mesh_scripts.py imports dolfinx / ufl / basix at module level (not installed here), but the
functions below only touch numpy and duck-typed topology objects.  This script imports the reference module with
stub modules standing in for those packages and runs, unmodified:

  * `_reference_segment_points`, `_reference_triangle_boundary_points`, `_reference_square_boundary_points`
    (:28-92) for N = 0..4;
  * `_reshape_map` (:195-214, row a5) on the facet -> cell connectivity of the four fixture meshes;
  * `_compute_integration_entities` (:137-192, row a6) for ds(100) (facets tagged 4 seen from cells tagged 1/2)
    and ds(101) (facets tagged 3 from cells tagged 2/3), with the reference's golden tags of each fixture mesh
    (only the three dolfinx-written meshes, whose numbering we reproduce).

The duck-typed mesh is phifem_b200.mesh.Mesh on CPU tensors (its `.topology` mirrors the dolfinx calls the reference
makes).  Run HERE; the output tests/golden/reference_helpers.npz is committed.

    python tests/golden/make_reference_helpers_fixture.py
"""
import importlib.util
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(HERE))


def import_reference():
    def stub(name, **attrs):
        m = types.ModuleType(name)
        for k, v in attrs.items():
            setattr(m, k, v)
        sys.modules[name] = m
        return m

    class _Any:
        def __init__(self, *a, **k):
            pass

    stub("dolfinx", mesh=None, fem=None)
    stub("ufl", inner=None, Measure=_Any)
    stub("basix")
    stub("basix.ufl", element=None)
    stub("dolfinx.cpp")
    stub("dolfinx.cpp.graph", AdjacencyList_int32=_Any)
    stub("dolfinx.fem", Function=_Any)
    stub("dolfinx.fem.petsc", assemble_vector=None)
    stub("dolfinx.mesh", Mesh=_Any, MeshTags=_Any)
    spec = importlib.util.spec_from_file_location("ref_mesh_scripts",
                                                  "/root/reference/src/phifem/mesh_scripts.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main():
    ref = import_reference()
    import cases
    from phifem_b200.mesh import Mesh
    out = {}
    for N in range(5):
        out["segment_%d" % N] = ref._reference_segment_points(N)
        out["triangle_%d" % N] = ref._reference_triangle_boundary_points(N)
        out["square_%d" % N] = ref._reference_square_boundary_points(N)
    for name in ("coarse_square", "square_tri", "square_quad", "disk"):
        x, cells, ct = cases.load_mesh_arrays(name)
        mesh = Mesh(x, cells, ct, device="cpu")
        tdim = mesh.topology.dim
        mesh.topology.create_connectivity(tdim - 1, tdim)
        mesh.topology.create_connectivity(tdim, tdim - 1)
        emap, width = ref._reshape_map(mesh.topology.connectivity(tdim - 1, tdim))
        out["reshape_f2c_" + name] = emap
        assert width == emap.shape[1]
    # integration entities with the reference's golden tags (box mode, degree 1)
    for data_name, mesh_name, _ in cases.TAG_DATA:
        if mesh_name == "disk":
            continue
        x, cells, ct = cases.load_mesh_arrays(mesh_name)
        mesh = Mesh(x, cells, ct, device="cpu")
        # compute_tags_measures has created the cell -> facet connectivity by the time it builds the measures (:419)
        mesh.topology.create_connectivity(mesh.topology.dim, mesh.topology.dim - 1)
        for single in (False, True):
            cname, fname = cases.golden_names(data_name, 1, True, True, single)
            gc, gf = cases.golden(cname), cases.golden(fname)
            if gc is None or gf is None:
                continue
            ctags = np.zeros(mesh.num_cells, dtype=np.int64)
            ctags[gc[0]] = gc[1]
            ftags = np.zeros(mesh.num_facets, dtype=np.int64)
            ftags[gf[0]] = gf[1]
            key = "%s%s" % (data_name, "_single" if single else "")
            e100 = ref._compute_integration_entities(mesh, np.nonzero((ctags == 1) | (ctags == 2))[0],
                                                     np.nonzero(ftags == 4)[0], 100)
            e101 = ref._compute_integration_entities(mesh, np.nonzero((ctags == 2) | (ctags == 3))[0],
                                                     np.nonzero(ftags == 3)[0], 101)
            assert e100[0][0] == 100 and e101[0][0] == 101
            out["ents100_" + key] = e100[0][1]
            out["ents101_" + key] = e101[0][1]
            out["ctags_" + key] = ctags.astype(np.int8)
            out["ftags_" + key] = ftags.astype(np.int8)
            out["mesh_" + key] = np.array(mesh_name)
    path = os.path.join(HERE, "reference_helpers.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), len(out), "arrays")


if __name__ == "__main__":
    main()
