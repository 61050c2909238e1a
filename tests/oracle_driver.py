"""Run the CPU oracle on a (mesh, level set) case -- shared by the golden tests and the GPU
parity tests."""
import numpy as np

from oracle import tags as OT
from phifem_b200 import fem
from phifem_b200.mesh import Mesh


def levelset_point_values(x, cells, cell_type, levelset, degree, discretize, mesh=None):
    """phi at the cell / facet detection points, through a P_degree interpolant or the expression."""
    pts = OT.cell_detection_points(cell_type, degree)
    fpts = OT.facet_points_in_cell(cell_type, degree)
    if discretize:
        mesh = mesh or Mesh(x, cells, cell_type, device="cpu")
        V = fem.functionspace(mesh, ("Lagrange", degree))
        fn = fem.Function(V).interpolate(levelset)
        phi_cell = OT.point_values_function(fn.x.array, V.dofmap, V.element.tabulate(pts))
        phi_facet = OT.point_values_function(fn.x.array, V.dofmap, V.element.tabulate(fpts))
        return pts, phi_cell, phi_facet, fn
    phi_cell = OT.point_values_expression(levelset, x, cells, cell_type, pts)
    phi_facet = OT.point_values_expression(levelset, x, cells, cell_type, fpts)
    return pts, phi_cell, phi_facet, None


def run_oracle(x, cells, cell_type, levelset, degree, discretize, box_mode, single_layer):
    pts, phi_cell, phi_facet, _ = levelset_point_values(x, cells, cell_type, levelset, degree,
                                                        discretize)
    return OT.compute_tags_measures(x, cells, cell_type, phi_cell, phi_facet, box_mode=box_mode,
                                    single_layer_cut=single_layer, detection_points=pts)
