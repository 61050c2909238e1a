"""The symbolic phase of the row-gather assembly behind the C ABI (csrc/rows_plan.cu, phifem_rows_plan_create) against
the torch-built plan (phifem_b200/assemble.py + rows.py, itself held to the oracle on the CPU by
tests/test_rows_symbolic.py): CSR pattern, active cells, ghost facets, both row lists -- rows, diagonal positions,
slice pointers, records, pads -- and the macro vertex lists must be equal BIT FOR BIT; the operator assembled from
either plan is then bitwise the same."""
import warnings

import numpy as np
import pytest
import torch

from phifem_b200 import assemble, fem, mesh_scripts, synthetic
from phifem_b200.assemble import AssemblyPlan, _plan_inputs

pytestmark = pytest.mark.gpu


def _problem(kind, n):
    if kind.startswith("tri"):
        mesh = synthetic.rectangle_mesh(n, device="cuda")
        center, radius = (0.013, -0.021), 0.61
    else:
        mesh = synthetic.box_mesh(n, device="cuda")
        center, radius = synthetic.SPHERE_CENTER, 0.37
    if kind.endswith("unstructured"):
        mesh = synthetic.unstructured_variant(mesh, jitter=0.2, seed=7)
    if kind.endswith("reordered"):
        mesh = synthetic.unstructured_variant(mesh, jitter=0.2, seed=7).reordered()
    phi = synthetic.sphere_levelset(mesh.x, center=center, radius=radius)
    f = torch.from_numpy(np.random.default_rng(1234).uniform(-1, 1, mesh.num_vertices)).cuda()
    fn = fem.Function(fem.functionspace(mesh, 1), phi.cpu().numpy())
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", RuntimeWarning)
        ctags, ftags, _, ds, _ = mesh_scripts.compute_tags_measures(mesh, fn, 1, box_mode=True)
    return mesh, phi, f, ctags, ftags, ds


def _same_list(a, b, name):
    assert a.n_listed == b.n_listed, name
    for attr in ("rows", "diag_pos", "ptr"):
        assert torch.equal(getattr(a, attr), getattr(b, attr)), "%s.%s" % (name, attr)
    assert a.n_records == b.n_records, name
    assert a.rec.numel() == b.rec.numel() and torch.equal(a.rec, b.rec), name + ".rec"


@pytest.mark.parametrize("order", ["natural", "morton"])
@pytest.mark.parametrize("kind,n", [("tri", 40), ("tri-unstructured", 24), ("tet", 10), ("tet-unstructured", 8),
                                    ("tet-reordered", 9)])
def test_native_plan_equals_the_torch_plan(kind, n, order):
    mesh, phi, f, ctags, ftags, ds = _problem(kind, n)
    pt = assemble.build_plan(mesh, ctags, ftags, ds(100), order=order, symbolic="torch")
    pn = assemble.build_plan(mesh, ctags, ftags, ds(100), order=order, symbolic="native")
    assert pt.symbolic == "torch" and pn.symbolic == "native" and pn.method == "rows"
    assert pn.nnz == pt.nnz and pn.n_rows == pt.n_rows
    for attr in ("indptr", "indices", "active", "ghost"):
        assert torch.equal(getattr(pn, attr), getattr(pt, attr)), attr
    rt, rn = pt.rowsplan, pn.rowsplan
    assert rn.max_row_nnz == rt.max_row_nnz and rn.order == rt.order == order
    _same_list(rn.cells, rt.cells, "cells")
    _same_list(rn.surface, rt.surface, "surface")
    assert torch.equal(rn.ghost_macro, rt.ghost_macro) and torch.equal(rn.entity_macro, rt.entity_macro)
    assert (rn.n_cell_records, rn.n_ghost_records, rn.n_entity_records) == \
        (rt.n_cell_records, rt.n_ghost_records, rt.n_entity_records)
    assert rt.surface.n_listed > 0 and rt.n_entity_records > 0
    # lazily built slot maps of the native plan = the torch plan's
    for attr in ("slots_cells", "slots_ghost", "slots_boundary"):
        assert torch.equal(getattr(pn, attr), getattr(pt, attr)), attr
    At, bt = assemble.assemble_strong_dirichlet(pt, phi, f, stab_coef=1.0)
    An, bn = assemble.assemble_strong_dirichlet(pn, phi, f, stab_coef=1.0)
    assert torch.equal(An.data, At.data) and torch.equal(bn, bt)


@pytest.mark.parametrize("kind,n", [("tri-unstructured", 24), ("tet", 9)])
def test_native_plan_with_a_row_mask(kind, n):
    """The rows a rank owns (multi-GPU): only they are listed, records of other rows are dropped."""
    mesh, phi, f, ctags, ftags, ds = _problem(kind, n)
    mask = torch.zeros(mesh.num_vertices, dtype=torch.bool, device="cuda")
    mask[torch.randperm(mesh.num_vertices, generator=torch.Generator().manual_seed(3))[: mesh.num_vertices // 2].cuda()] = True
    c8, f8, ents = _plan_inputs(mesh, ctags, ftags, ds(100))
    pt = AssemblyPlan(mesh, c8, f8, ents, row_mask=mask, symbolic="torch")
    pn = AssemblyPlan(mesh, c8, f8, ents, row_mask=mask, symbolic="native")
    for attr in ("indptr", "indices", "active", "ghost"):
        assert torch.equal(getattr(pn, attr), getattr(pt, attr)), attr
    _same_list(pn.rowsplan.cells, pt.rowsplan.cells, "cells")
    _same_list(pn.rowsplan.surface, pt.rowsplan.surface, "surface")
    dt, bt = pt.new_outputs()
    dn, bn = pn.new_outputs()
    assemble.assemble_into(pt, phi, f, 1.0, dt, bt)
    assemble.assemble_into(pn, phi, f, 1.0, dn, bn)
    assert torch.equal(dn, dt) and torch.equal(bn, bt)


def test_native_plan_edge_cases():
    """No active cell at all; no surface entity (level set negative everywhere: every cell interior)."""
    mesh = synthetic.box_mesh(4, device="cuda")
    for center, radius in (((9.0, 9.0, 9.0), 0.5), (synthetic.SPHERE_CENTER, 5.0)):
        phi = synthetic.sphere_levelset(mesh.x, center=center, radius=radius)
        fn = fem.Function(fem.functionspace(mesh, 1), phi.cpu().numpy())
        with warnings.catch_warnings():
            warnings.simplefilter("ignore", RuntimeWarning)
            ctags, ftags, _, ds, _ = mesh_scripts.compute_tags_measures(mesh, fn, 1, box_mode=True)
        pt = assemble.build_plan(mesh, ctags, ftags, ds(100), symbolic="torch")
        pn = assemble.build_plan(mesh, ctags, ftags, ds(100), symbolic="native")
        assert pn.nnz == pt.nnz
        assert torch.equal(pn.indptr, pt.indptr) and torch.equal(pn.indices, pt.indices)
        if pt.nnz:
            _same_list(pn.rowsplan.cells, pt.rowsplan.cells, "cells")
            _same_list(pn.rowsplan.surface, pt.rowsplan.surface, "surface")
            f = torch.ones_like(phi)
            At, bt = assemble.assemble_strong_dirichlet(pt, phi, f)
            An, bn = assemble.assemble_strong_dirichlet(pn, phi, f)
            assert torch.equal(An.data, At.data) and torch.equal(bn, bt)


def test_csr_matrix_outlives_its_plan():
    """`indptr` / `indices` of a natively built plan are zero-copy views of memory the C handle owns: the tensors keep
    the handle alive after the plan object is gone (the demos keep A, not the plan)."""
    import gc
    mesh, phi, f, ctags, ftags, ds = _problem("tri-unstructured", 20)
    A, b = assemble.assemble_strong_dirichlet(assemble.build_plan(mesh, ctags, ftags, ds(100), symbolic="native"), phi, f)
    ref = assemble.build_plan(mesh, ctags, ftags, ds(100), symbolic="torch")
    gc.collect()
    torch.cuda.synchronize()
    junk = [torch.full((1 << 20,), 7, dtype=torch.int32, device="cuda") for _ in range(8)]   # reuse freed memory, if any
    assert torch.equal(A.indptr, ref.indptr) and torch.equal(A.indices, ref.indices)
    assert A.to_scipy().nnz == ref.nnz
    del junk
