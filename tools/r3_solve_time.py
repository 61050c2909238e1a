"""Where does the time of solve.bicgstab go at config E?  (setup on torch ops / iterations on the fused kernels)"""
import os, sys, time, warnings
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from phifem_b200 import assemble, fem, mesh_scripts, solve, synthetic

N = int(os.environ.get("N", "204"))
mesh = synthetic.box_mesh(N, device="cuda")
phi = synthetic.sphere_levelset(mesh.x)
f = synthetic.ball_source(mesh.x)
fn = fem.Function(fem.functionspace_p1_device(mesh), phi)
with warnings.catch_warnings():
    warnings.simplefilter("ignore", RuntimeWarning)
    ct, ft, _, ds, _ = mesh_scripts.compute_tags_measures(mesh, fn, 1, box_mode=True)
A, b = assemble.assemble_strong_dirichlet(assemble.build_plan(mesh, ct, ft, ds(100)), phi, f)
for maxiter in (0, 10, 110, 4000):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    x, info = solve.bicgstab(A, b, rtol=1e-8, maxiter=maxiter)
    torch.cuda.synchronize()
    print("maxiter", maxiter, "-> %.1f ms" % ((time.perf_counter() - t0) * 1e3), info)
