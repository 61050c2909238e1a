"""Where do the staged and the vector-load facet kernels differ at full size?"""
import os, sys, warnings
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import torch
from phifem_b200 import synthetic, fem, mesh_scripts

N = int(os.environ.get("N", "204"))
mesh = synthetic.box_mesh(N, device="cuda")
phi = synthetic.sphere_levelset(mesh.x)
fn = fem.Function(fem.functionspace_p1_device(mesh), phi)
out = {}
for kind in ("ldg", "staged"):
    os.environ["PHIFEM_FACETS_KERNEL"] = kind
    os.environ["PHIFEM_CELLS_KERNEL"] = kind
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", RuntimeWarning)
        ct, ft, _, ds, _ = mesh_scripts.compute_tags_measures(mesh, fn, 1, box_mode=True)
    out[kind] = (ct.tags8.clone(), ft.tags8.clone())
torch.cuda.synchronize()
for i, what in enumerate(("cells", "facets")):
    a, b = out["ldg"][i], out["staged"][i]
    bad = torch.nonzero(a != b).reshape(-1)
    print(what, "mismatches:", bad.numel(), "of", a.numel())
    if bad.numel():
        print(" first", bad[:16].tolist(), "last", bad[-4:].tolist())
        tiles = torch.unique(bad // 1024)
        print(" tiles with mismatches:", tiles.numel(), tiles[:20].tolist())
        print(" positions in tile (first 32):", (bad[:32] % 1024).tolist())
        print(" ldg   :", a[bad[:16]].tolist())
        print(" staged:", b[bad[:16]].tolist())
