"""Golden inputs of BASELINE.json configs[0] (reference demo/weak-dirichlet/flower): the reference's own
`detection_levelset`, `levelset` and `source_term` (demo/weak-dirichlet/flower/data.py) evaluated at the vertices
of the demo's 200 x 200 background mesh on [-4.5, 4.5]^2 (main.py:45-46).  Run HERE (the reference is not on the
GPU box); the output tests/golden/flower_demo.npz is committed.

    python tests/golden/make_flower_fixture.py
"""
import os
import sys

import numpy as np

sys.path.insert(0, "/root/reference/demo/weak-dirichlet/flower")
import data  # noqa: E402  (the reference's module)

n = 200
t = np.linspace(-4.5, 4.5, n + 1)
X, Y = np.meshgrid(t, t, indexing="ij")
x = np.stack([X.ravel(), Y.ravel()])
out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "flower_demo.npz")
np.savez_compressed(out, n=n, detection=data.detection_levelset(x), levelset=data.levelset(x),
                    source=data.source_term(x), dirichlet=data.dirichlet_data(x))
print(out, os.path.getsize(out))
