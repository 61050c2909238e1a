"""phifem_b200: B200-native level-set cut-cell classification + phi-FEM CSR assembly."""


def release_scratch():
    """Hand the cached scratch of the symbolic entry points (sort buffers of `phifem_rows_plan_create`,
    `phifem_pattern_create_p1`, `phifem_integration_entities`: a private stream-ordered CUDA memory pool that stays
    cached between calls up to 6 GiB) back to the driver, e.g. before other allocators need the memory."""
    from . import _lib
    _lib.load().phifem_pattern_release_scratch()
