"""CPU oracle for the phiFEM hot path -- TEST INFRASTRUCTURE ONLY.

A restatement (numpy + a small C file) of the algorithm of the reference
`phifem.mesh_scripts.compute_tags_measures` (reference src/phifem/mesh_scripts.py)
and of the strong-Dirichlet phi-FEM operator assembled by the reference demo
(reference demo/strong-dirichlet/flower/main.py:92-131).

Who may import this package: `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py`.  The product
(`phifem_b200/`) never imports it; there is no CPU fallback in the product.

Parity status
  * tags (cells, facets, one-sided entities, submesh transfer): PINNED against the
    reference's golden CSVs (tests/golden/golden_tags.npz, generated from
    reference tests/tests_data/*.csv) and the 18 known answers of
    reference tests/test_one_sided_integral.py -- see tests/test_oracle_golden.py.
  * assembled CSR operator / load vector: PARITY UNPINNED -- the reference has no
    test that pins an assembled matrix and dolfinx/FFCx/PETSc cannot be run in this
    image.  The closed forms are cross-checked against an independent brute-force
    quadrature restatement (oracle/assembly.py) and property tests instead.
"""
