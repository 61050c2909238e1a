"""Experiment (round 2, session 4): device-resident steps of config E on TWO streams -- step i on stream i % 2 with its
own tag workspace, CSR values, load vector and facet-once scratch -- against the same K steps on one stream.  The tag
kernels are HBM-bound, the cell pass is fp64-issue-bound: do they share the SMs?"""
import argparse
import ctypes
import sys
import time

import torch

sys.path.insert(0, ".")
import bench  # noqa: E402
from phifem_b200 import _lib, mesh_scripts, synthetic  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=204)
ap.add_argument("--steps", type=int, default=40)
a = ap.parse_args()
args = bench.parser().parse_args(["--no-replan"])
args.n = a.n
dev = torch.device("cuda", 0)
mesh = synthetic.box_mesh(a.n, device=dev)
phi = synthetic.sphere_levelset(mesh.x)
f = synthetic.ball_source(mesh.x)
w = bench.Workload(mesh, phi, f, args)
plan, rp = w.plan, w.plan.rowsplan
lib = _lib.load()


class Lane:
    def __init__(self, first):
        self.stream = torch.cuda.Stream()
        self.ws = w.ws if first else mesh_scripts.TagWorkspace(mesh)
        self.data, self.b = (w.data, w.b) if first else plan.new_outputs()
        self.c = _lib.CRowsPlan()
        ctypes.pointer(self.c)[0] = rp.c_struct()
        if not first:
            self.work = torch.empty_like(rp.surface_work)
            self.c.surface_work = self.work.data_ptr()

    def step(self):
        mesh_scripts.classify_cells(mesh, w.dls, self.ws)
        mesh_scripts.classify_facets(mesh, w.dls, self.ws)
        _lib.check(lib.phifem_assemble_rows_p1(_lib.c_mesh(mesh), _lib.ptr(phi), _lib.ptr(f), 1.0, ctypes.byref(self.c),
                                               _lib.ptr(self.data), _lib.ptr(self.b), _lib.stream()))


lanes = [Lane(True), Lane(False)]


def run(n_lanes, steps):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(steps):
        ln = lanes[i % n_lanes]
        with torch.cuda.stream(ln.stream):
            ln.step()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / steps * 1e3


for _ in range(2):
    run(1, 6)
    run(2, 6)
for rep in range(3):
    print("one stream  %.4f ms / step" % run(1, a.steps))
    print("two streams %.4f ms / step" % run(2, a.steps))
ref = lanes[0].data.clone()
run(2, 4)
print("lanes agree bitwise:", bool(torch.equal(lanes[0].data, lanes[1].data) and torch.equal(lanes[0].data, ref)
                                   and torch.equal(lanes[0].ws.facet_tags8, lanes[1].ws.facet_tags8)))
