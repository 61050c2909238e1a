"""Drop-in module name of the reference (src/phifem/mesh_scripts.py): everything is implemented in
phifem_b200.mesh_scripts; `from phifem.mesh_scripts import compute_tags_measures` keeps working."""
from phifem_b200.mesh_scripts import *  # noqa: F401,F403
from phifem_b200.mesh_scripts import (_compute_integration_entities, _overwrite_tags,  # noqa: F401
                                      _reference_segment_points, _reference_square_boundary_points,
                                      _reference_triangle_boundary_points, _reshape_map, _tag_cells,
                                      _tag_facets, _transfer_tags, compute_meshtags,
                                      compute_tags_measures, debug_mode)
