"""Golden tags produced by the REFERENCE'S OWN `_tag_cells` / `_tag_facets` code (rows a3 / a4 of SURVEY.md 8a).

Everything in those two functions downstream of the detection vector is numpy: the exact comparisons with +-1
(:343-347), `single_layer_cut` (:349-358) and the ~30 set operations of the facet algebra (:454-496).  The only
dolfinx-dependent piece is `_compute_detection_vector` (:95-134, two DG0 forms assembled by dolfinx).  This script
imports the unmodified reference module with stubbed dolfinx / ufl / basix imports, answers the two
`assemble_vector` calls inside `_compute_detection_vector` with the oracle's sums (oracle/tags.py; that piece is
pinned separately by the reference's golden CSVs) -- the ratio, its 0.5 fallback and the warning (:124-133) run for
real -- and runs the reference's code on

  * the reference's triangle / quadrilateral fixture meshes (both functions), and
  * jittered, permuted, relabelled triangle and TETRAHEDRON meshes -- the reference's `_tag_cells` refuses
    tetrahedra before it ever looks at the detection vector (:320-329), so for them the stand-in mesh reports its
    cell type as "triangle" to get past that guard; nothing after the guard depends on the cell type.  This is
    what pins the 3D extension to the reference's own post-detection logic.

Output: tests/golden/reference_tagging.npz (committed; meshes included, they are small).

    python tests/golden/make_reference_tagging_fixture.py
"""
import os
import sys
import types
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)

from make_reference_helpers_fixture import import_reference  # noqa: E402


class Tags:
    def __init__(self, mesh, dim, indices, values):
        self.dim, self.indices, self.values = dim, np.asarray(indices), np.asarray(values)

    def find(self, v):
        return self.indices[self.values == v]


def main():
    ref = import_reference()
    import cases
    from oracle import tags as OT
    from phifem_b200 import synthetic
    from phifem_b200.mesh import Mesh

    # the dolfinx calls the two functions make besides the detection vector
    ref.dfx.mesh = types.SimpleNamespace(
        meshtags=Tags,
        locate_entities_boundary=lambda mesh, fdim, marker: np.nonzero(mesh._f2c[:, 1] < 0)[0].astype(np.int32))

    def run(x, cells, cell_type, phi_v, single):
        """phi_v: vertex values of a P1 level set; detection degree 1."""
        mesh = Mesh(x, cells, cell_type, device="cpu")
        c2f, f2c, _ = OT.build_topology(cells.astype(np.int64), cell_type)
        assert np.array_equal(c2f, mesh.c2f.numpy())
        mesh._f2c = f2c
        pts = OT.cell_detection_points(cell_type, 1)
        fpts = OT.facet_points_in_cell(cell_type, 1)
        ftab = np.asarray([OT.coordinate_basis(cell_type, p)[0] for p in fpts])
        phi_cell = phi_v[cells]
        phi_facet = OT.point_values_function(phi_v, cells, ftab)
        num_c, den_c = OT.detection_sums_cells(phi_cell, OT.cell_scale(x, cells, cell_type, pts))
        num_f, den_f = OT.detection_sums_facets(phi_facet, OT.facet_scale(x, cells, cell_type), c2f, f2c)
        # `_compute_detection_vector` (:95-134) runs for real; only the two dolfinx `assemble_vector` calls inside
        # it are answered with the oracle's sums (the UFL expression it builds is reduced to a (which, measure) key)
        sums = {("num", "dx"): num_c, ("den", "dx"): den_c, ("num", "ds"): num_f, ("den", "ds"): den_f}

        class Integrand:
            def __init__(self, which):
                self.which = which

            def __mul__(self, measure):
                return (self.which, measure.kind)

        class LevelSet:
            def __abs__(self):
                return "abs"

        ref.inner = lambda a, v0: Integrand("den" if a == "abs" else "num")
        ref.element = lambda *a, **k: None
        ref.dfx.fem = types.SimpleNamespace(functionspace=lambda m, e: None, form=lambda f: f)
        ref.ufl.TestFunction = lambda space: "v0"
        ref.assemble_vector = lambda form: types.SimpleNamespace(array=sums[form].copy())
        ref.ufl.Measure = lambda kind, **kw: types.SimpleNamespace(kind=kind)
        levelset = LevelSet()
        if cell_type == "tetrahedron":       # get past the guard of :326-329 (see the module docstring)
            mesh.topology.cell_type.name = "triangle"
        with warnings.catch_warnings():
            warnings.simplefilter("ignore", RuntimeWarning)    # the ds detection warns on every mesh (:129-133)
            ct = ref._tag_cells(mesh, levelset, 1, single_layer_cut=single)
            ft = ref._tag_facets(mesh, ct, levelset, 1)
        cd = np.zeros(mesh.num_cells, dtype=np.int8)
        cd[ct.indices] = ct.values
        fd = np.zeros(len(f2c), dtype=np.int8)
        fd[ft.indices] = ft.values
        assert len(np.unique(ft.indices)) == len(ft.indices), "the reference emitted a facet twice"
        # submesh route (:635-645): the reference's `_transfer_tags` (:217-281) onto the submesh of Omega_h.
        # dolfinx's create_submesh is not available; the submesh stand-in keeps the cells tagged 1/2 in
        # ascending order and renumbers their vertices ascending (oracle/tags.py `submesh`)
        keep = np.nonzero((cd == 1) | (cd == 2))[0]
        sx, scells, _ = OT.submesh(x, cells.astype(np.int64), keep)
        sub = Mesh(sx, scells, cell_type, device="cpu")
        sct = ref._transfer_tags(ct, sub, keep)
        sft = ref._transfer_tags(ft, sub, keep, source_mesh=mesh)
        assert np.array_equal(sct.indices, np.arange(len(keep))) and np.array_equal(sft.indices,
                                                                                    np.arange(sub.num_facets))
        # user overlay (:561-568): tag 7 on every fifth cell, 9 on every seventh facet
        oc = ref._overwrite_tags(mesh, ct, Tags(mesh, ct.dim, np.arange(0, mesh.num_cells, 5, dtype=np.int32),
                                                np.full(len(range(0, mesh.num_cells, 5)), 7, dtype=np.int32)))
        of = ref._overwrite_tags(mesh, ft, Tags(mesh, ft.dim, np.arange(0, len(f2c), 7, dtype=np.int32),
                                                np.full(len(range(0, len(f2c), 7)), 9, dtype=np.int32)))
        extra = {"sub_ctags": sct.values.astype(np.int8), "sub_ftags": sft.values.astype(np.int8),
                 "ow_c_idx": oc.indices.astype(np.int32), "ow_c_val": oc.values.astype(np.int32),
                 "ow_f_idx": of.indices.astype(np.int32), "ow_f_val": of.values.astype(np.int32)}
        return cd, fd, extra

    out, names = {}, []

    def add(name, x, cells, cell_type, phi_v):
        for single in (False, True):
            key = name + ("_single" if single else "")
            ct, ft, extra = run(x, cells, cell_type, phi_v, single)
            for k, v in extra.items():
                out[k + "_" + key] = v
            out["x_" + key], out["cells_" + key] = x, cells.astype(np.int32)
            out["type_" + key], out["phi_" + key] = np.array(cell_type), phi_v
            out["ctags_" + key], out["ftags_" + key] = ct, ft
            names.append(key)
            print("%-40s cells %6d  tags %s  facet tags %s" % (key, len(cells), np.bincount(ct, minlength=4)[1:],
                                                              np.bincount(ft, minlength=7)[1:]))

    for data_name, mesh_name, func in cases.TAG_DATA:
        if data_name in ("square_in_square", "nasty_levelset"):
            continue                          # degenerate at the vertices
        x, cells, ct = cases.load_mesh_arrays(mesh_name)
        x3 = np.zeros((3, len(x)))
        x3[:2] = x.T
        add("fixture_" + data_name, x, cells, ct, np.asarray(func(x3), dtype=np.float64))
    for kind, n, seed in (("tri", 24, 3), ("tet", 7, 5), ("tet", 9, 8)):
        base = synthetic.rectangle_mesh(n, device="cpu") if kind == "tri" else synthetic.box_mesh(n, device="cpu")
        m = synthetic.unstructured_variant(base, jitter=0.2, seed=seed)
        if kind == "tri":
            phi = synthetic.sphere_levelset(m.x, center=(0.013, -0.021), radius=0.61)
        elif seed == 5:
            phi = synthetic.sphere_levelset(m.x, radius=0.37)
        else:                                  # sphere leaving the box: cut cells on the mesh boundary
            phi = synthetic.sphere_levelset(m.x, center=(0.31, 0.52, 0.48), radius=0.45)
        add("synthetic_%s_%d" % (kind, n), m.x.numpy(), m.cells.numpy(), m.cell_type, phi.numpy())
    out["names"] = np.array(names)
    path = os.path.join(HERE, "reference_tagging.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path))


if __name__ == "__main__":
    main()
