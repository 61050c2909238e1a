"""CPU-only tests of the host-side plumbing of the product: facet numbering / connectivities (torch
sort-unique builder) and the symbolic phase of the assembly, against the oracle's numpy versions."""
import numpy as np
import pytest
import torch

import cases
from oracle import assembly as OA
from oracle import tags as OT
from phifem_b200 import assemble, fem, mesh_scripts, synthetic
from phifem_b200.mesh import Mesh, MeshTags


@pytest.mark.parametrize("name", ["coarse_square", "disk", "square_tri", "square_quad"])
def test_topology_builder_matches_oracle_on_fixtures(name):
    x, cells, ct = cases.load_mesh_arrays(name)
    m = Mesh(x, cells, ct, device="cpu")
    c2f, f2c, fv = OT.build_topology(cells, ct)
    assert np.array_equal(m.c2f.numpy(), c2f)
    assert np.array_equal(m.f2c.numpy(), f2c)
    assert np.array_equal(m.facet_vertices.numpy(), fv)


def test_topology_builder_tetrahedra_and_counts():
    n = 5
    m = synthetic.unstructured_variant(synthetic.box_mesh(n, device="cpu"), seed=3)
    c2f, f2c, fv = OT.build_topology(m.cells.numpy(), "tetrahedron")
    assert np.array_equal(m.c2f.numpy(), c2f) and np.array_equal(m.f2c.numpy(), f2c)
    assert np.array_equal(m.facet_vertices.numpy(), fv)
    # Nc = 6 n^3, Nv = (n+1)^3, Nf = 12 n^3 + 6 n^2
    assert m.num_cells == 6 * n ** 3 and m.num_vertices == (n + 1) ** 3
    assert m.num_facets == 12 * n ** 3 + 6 * n ** 2
    m2 = synthetic.rectangle_mesh(7, device="cpu")
    assert m2.num_cells == 2 * 49 and m2.num_facets == 3 * 49 + 2 * 7


def test_dolfinx_style_connectivity_accessors():
    x, cells, ct = cases.load_mesh_arrays("coarse_square")
    m = Mesh(x, cells, ct, device="cpu")
    tdim = m.topology.dim
    m.topology.create_connectivity(tdim - 1, tdim)
    f2c = m.topology.connectivity(tdim - 1, tdim)
    emap, width = mesh_scripts._reshape_map(f2c)
    assert width == 2 and emap.shape == (m.num_facets, 2)
    ref = m.f2c.numpy()
    interior = ref[:, 1] >= 0
    assert np.array_equal(emap[interior], ref[interior][:, ::-1])     # reverse link order (:195-214)
    assert np.array_equal(emap[~interior, 0], ref[~interior, 0]) and np.all(emap[~interior, 1] == -1)
    m.topology.create_connectivity(0, tdim)
    v2c = m.topology.connectivity(0, tdim)
    for v in range(m.num_vertices):
        assert set(v2c.links(v)) == set(np.nonzero((cells == v).any(axis=1))[0])
    assert m.topology.cell_name() == "triangle" and m.geometry.x.shape == (16, 3)


@pytest.mark.parametrize("d", [2, 3])
def test_symbolic_phase_matches_oracle_pattern(d):
    m = synthetic.rectangle_mesh(10, device="cpu") if d == 2 else synthetic.box_mesh(5, device="cpu")
    m = synthetic.unstructured_variant(m, jitter=0.1, seed=5)
    x, cells = m.x.numpy(), m.cells.numpy().astype(np.int64)
    center = np.array([0.1, 0.05]) if d == 2 else np.array([0.52, 0.49, 0.51])
    r = 0.6 if d == 2 else 0.33
    phi = ((x - center) ** 2).sum(axis=1) - r * r
    ct = m.cell_type
    pts = OT.cell_detection_points(ct, 1)
    ftab = np.asarray([OT.coordinate_basis(ct, p)[0] for p in OT.facet_points_in_cell(ct, 1)])
    out = OT.compute_tags_measures(x, cells, ct, phi[cells], OT.point_values_function(phi, cells, ftab),
                                   box_mode=True, detection_points=pts)
    ctags = MeshTags(m, d, torch.from_numpy(out["cell_tags"]))
    ftags = MeshTags(m, d - 1, torch.from_numpy(out["facet_tags"]))
    plan = assemble.build_plan(m, ctags, ftags, out["ds100"])
    active = np.nonzero(out["cell_tags"] != 3)[0]
    ghost = np.nonzero(np.isin(out["facet_tags"], (2, 3)) & (out["f2c"][:, 1] >= 0))[0]
    ip, ix = OA.sparsity_pattern(len(x), cells, active, ghost, out["f2c"])
    assert np.array_equal(plan.indptr.numpy(), ip) and np.array_equal(plan.indices.numpy(), ix)
    assert np.array_equal(plan.active.numpy(), active) and np.array_equal(plan.ghost.numpy(), ghost)
    # every slot points at the (row, col) pair it stands for
    rows = np.repeat(np.arange(len(x)), np.diff(ip))
    nv = d + 1
    sl = plan.slots_cells.numpy()
    dm = cells[active]
    assert np.array_equal(rows[sl], np.repeat(dm, nv, axis=1))
    assert np.array_equal(ix[sl], np.tile(dm, (1, nv)))
    mac = assemble.ghost_macro_vertices(m, plan.ghost).numpy()
    for e, fct in enumerate(ghost):                      # distinct vertices of the two cells, cell + first
        cp, cm = out["f2c"][fct]
        assert set(mac[e]) == set(cells[cp]) | set(cells[cm]) and len(set(mac[e])) == nv + 1
        assert list(mac[e][:nv - 1]) == [v for v in cells[cp] if v in cells[cm]]
    sg = plan.slots_ghost.numpy()
    assert np.array_equal(rows[sg], np.repeat(mac, nv + 1, axis=1))
    assert np.array_equal(ix[sg], np.tile(mac, (1, nv + 1)))
    ents = out["ds100"].reshape(-1, 2)
    sb = plan.slots_boundary.numpy()
    assert np.array_equal(ix[sb], np.tile(cells[ents[:, 0]], (1, nv)))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        assemble.assemble_strong_dirichlet(plan, phi, phi)


def test_lagrange_spaces_partition_of_unity_and_interpolation():
    for name, degs in (("square_tri", (1, 2, 3)), ("square_quad", (1, 2, 3))):
        x, cells, ct = cases.load_mesh_arrays(name)
        m = Mesh(x, cells, ct, device="cpu")
        for k in degs:
            V = fem.functionspace(m, ("Lagrange", k))
            pts = OT.cell_detection_points(ct, 3)
            tab = V.element.tabulate(pts)
            assert np.allclose(tab.sum(axis=1), 1.0, atol=1e-12)
            # nodal basis
            assert np.allclose(V.element.tabulate(V.element.nodes, snap=False), np.eye(V.element.ndofs),
                               atol=1e-12)
            # interpolating a degree-k polynomial reproduces it at the detection points
            poly = lambda X: (0.3 * X[0] - 0.2 * X[1] + 0.1) ** k
            fn = fem.Function(V).interpolate(poly)
            vals = OT.point_values_function(fn.x.array, V.dofmap, tab)
            xq = OT.physical_points(x, cells, ct, pts)
            assert np.allclose(vals, poly(np.moveaxis(xq, 2, 0)), atol=1e-12)
    m3 = synthetic.box_mesh(2, device="cpu")
    for k in (1, 2, 3):
        V = fem.functionspace(m3, k)
        assert V.element.ndofs == {1: 4, 2: 10, 3: 20}[k]


def test_user_tags_outside_int8_do_not_wrap_into_computed_values():
    """`overwrite_tags` values are free apart from 1..6 / 100 / 101 (reference src/phifem/mesh_scripts.py:606-615);
    the one-byte arrays the kernels read must not alias 257 -> 1, 260 -> 4, 200 -> -56."""
    v = torch.tensor([0, 1, 2, 3, 4, 5, 6, 7, 100, 127, 128, 200, 255, 256, 257, 258, 260, 262, -1, 65540],
                     dtype=torch.int32)
    t8 = mesh_scripts._narrow_tags(v)
    assert t8.dtype == torch.int8
    assert t8.tolist() == [0, 1, 2, 3, 4, 5, 6] + [0] * 13


@pytest.mark.parametrize("kind,curve", [("tet", "pencil"), ("tri", "pencil"), ("tet", "morton"), ("tri", "morton")])
def test_mesh_renumbering_maps_and_pencil_order(kind, curve):
    """Mesh.reordered: the two maps translate cells and vertices, local vertex order is kept; curve="pencil" on the
    scrambled variant of SURVEY.md 8(d) (jitter below half a spacing) is the lexicographic numbering of the grid."""
    base = synthetic.box_mesh(7, device="cpu") if kind == "tet" else synthetic.rectangle_mesh(11, device="cpu")
    scr = synthetic.unstructured_variant(base, jitter=0.2, seed=5)
    new = scr.reordered(curve)
    assert torch.equal(new.x, scr.x[new.input_global_indices])
    old_cells = scr.cells[new.original_cell_index].long()
    assert torch.equal(new.input_global_indices[new.cells.long()], old_cells)       # same cells, same local order
    assert torch.equal(torch.sort(new.input_global_indices).values, torch.arange(scr.num_vertices))
    assert torch.equal(torch.sort(new.original_cell_index).values, torch.arange(scr.num_cells))
    if curve == "pencil":
        n = 7 if kind == "tet" else 11
        h = float((base.x.max() - base.x.min()) / n)
        assert float((new.x - base.x).abs().max()) <= 0.2 * h + 1e-12             # vertex i of the grid is vertex i again
        assert torch.equal(new.cells.min(dim=1).values, base.cells.min(dim=1).values)   # cells grouped by their cube


def test_plan_matches_says_when_the_tags_need_a_new_plan():
    """`AssemblyPlan.matches`: True for the tags the plan was built from (also after the caller rewrote its arrays in place
    with the same values, and whatever happens to facet tags the plan never looks at), False as soon as a cell enters or
    leaves dx((1,2)) / dx(2) or a facet enters or leaves dS((2,3)) / Gamma_h."""
    m = synthetic.unstructured_variant(synthetic.rectangle_mesh(12, device="cpu"), jitter=0.1, seed=5)
    x, cells = m.x.numpy(), m.cells.numpy().astype(np.int64)

    def tags_of(radius):
        phi = ((x - np.array([0.1, 0.05])) ** 2).sum(axis=1) - radius * radius
        pts = OT.cell_detection_points("triangle", 1)
        ftab = np.asarray([OT.coordinate_basis("triangle", p)[0] for p in OT.facet_points_in_cell("triangle", 1)])
        return OT.compute_tags_measures(x, cells, "triangle", phi[cells], OT.point_values_function(phi, cells, ftab),
                                        box_mode=True, detection_points=pts)

    out = tags_of(0.6)
    c8, f8 = torch.from_numpy(out["cell_tags"]).to(torch.int8), torch.from_numpy(out["facet_tags"]).to(torch.int8)
    ctags, ftags = MeshTags(m, 2, None, tags8=c8), MeshTags(m, 1, None, tags8=f8)
    plan = assemble.build_plan(m, ctags, ftags, out["ds100"])
    assert plan.matches(ctags, ftags)
    assert plan.matches(MeshTags(m, 2, torch.from_numpy(out["cell_tags"])), MeshTags(m, 1, torch.from_numpy(out["facet_tags"])))
    # facet tags 1 / 5 / 6 are not part of any integral of the operator
    f_other = f8.clone()
    f_other[f_other == 5] = 6
    assert plan.matches(ctags, MeshTags(m, 1, None, tags8=f_other))
    # the interface moved: other cut cells, other ghost facets
    out2 = tags_of(0.63)
    assert not np.array_equal(out2["cell_tags"], out["cell_tags"])
    moved_c = MeshTags(m, 2, torch.from_numpy(out2["cell_tags"]))
    moved_f = MeshTags(m, 1, torch.from_numpy(out2["facet_tags"]))
    assert not plan.matches(moved_c, moved_f) and not plan.matches(moved_c, ftags) and not plan.matches(ctags, moved_f)
    # the caller's arrays rewritten in place: the plan kept its own copy of what it depends on
    c8.copy_(torch.from_numpy(out2["cell_tags"]).to(torch.int8))
    assert not plan.matches(ctags, ftags)
    c8.copy_(torch.from_numpy(out["cell_tags"]).to(torch.int8))
    assert plan.matches(ctags, ftags)
    # one facet leaving Gamma_h is enough
    f_one = f8.clone()
    f_one[torch.nonzero(f_one == 4)[0]] = 5
    assert not plan.matches(ctags, MeshTags(m, 1, None, tags8=f_one))
