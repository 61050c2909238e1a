#!/usr/bin/env python
"""Benchmark of the phi-FEM hot path: cut-cell tags + CSR assembly, cells/s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--n 204] [--impl ours|reference]

One "step" = one pass of the hot path over the synthetic 3D P1 configuration of BASELINE.json
(config E): classify all cells and facets of a 6 n^3 Kuhn-tetrahedra unit cube
against a sphere level set, then assemble the strong-Dirichlet phi-FEM operator and load vector into
CSR (pattern / slot maps are the precomputed symbolic phase, timed separately as `symbolic_ms`).
Prints ONE JSON line (rank 0).  `value` times the device-resident pass with CUDA events; `e2e` times
the public API with pinned host buffers for the level set, the source term and all results.
"""
import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "phi-FEM cells assembled/s (tags+CSR)"
UNIT = "cells/s"
# 2D configurations (BASELINE.json configs[1..2]): disc in [-1, 1]^2, offsets keep |phi_v| >> ulp
DISC_CENTER = (3.141592653589793 / 1000.0, 2.718281828459045 / 1000.0)
DISC_RADIUS = 0.6
CONFIGS = {   # name -> (default n, description)
    "3d-p1": (204, "synthetic 3D P1 phi-FEM Poisson (BASELINE.json configs[4]): %d Kuhn tetrahedra per GPU "
                   "(n=%d), sphere level set"),
    "2d-p1": (1414, "synthetic 2D P1 phi-FEM Poisson (BASELINE.json configs[1] scale, 4 M triangles): %d "
                    "triangles (n=%d), disc level set"),
    "2d-p2": (2828, "synthetic 2D P2 phi-FEM Poisson (BASELINE.json configs[2], 16 M triangles): %d triangles "
                    "(n=%d), disc level set, P2 trial/test space and P2 level set, P1 detection"),
    "3d-p2": (64, "synthetic 3D P2 phi-FEM Poisson: %d Kuhn tetrahedra (n=%d), sphere level set, P2 trial/test "
                  "space and P2 level set, P1 detection"),
}


def _peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons of one GPU through NVML while the timed region runs."""

    def __init__(self, index, period=0.004):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop_evt = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {}
        for key in dir(nv):
            if key.startswith("nvmlClocksEventReason") or key.startswith("nvmlClocksThrottleReason"):
                val = getattr(nv, key)
                if isinstance(val, int) and val and "All" not in key and "None" not in key:
                    names.setdefault(val, key.replace("nvmlClocksEventReason", "")
                                     .replace("nvmlClocksThrottleReason", ""))
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    bits = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    bits = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if bits & bit and name not in ("GpuIdle", "ApplicationsClocksSetting"):
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(self.period)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=1.0)
        return {"sm_mhz": statistics.median(self.samples) if self.samples else None,
                "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


def algorithmic_bytes(counts):
    """compulsory traffic, each input read once / each output written once."""
    nc, nv, nf, gdim, nvpc = counts["Nc"], counts["Nv"], counts["Nf"], counts["gdim"], counts["nvpc"]
    na, nva, ng, nnz = counts["Na"], counts["Nv_active"], counts["Ng"], counts["nnz"]
    nd, nrow = counts.get("nd", nvpc), counts.get("Nrow", nv)
    ndof_a = counts.get("Ndof_active", nva)          # active dofs of the trial/test space (= of phi and f)
    # SURVEY.md section 8(d) as written: tags counted as the int32 arrays of the reference's MeshTags (4 Nc + 4 Nf).
    # The kernels write them as one byte per entity and widen on demand, so their DRAM traffic (`roofline.traffic`,
    # profiles/) is below this yardstick for the tag half.
    b_tags_cells = 4 * nvpc * nc + 8 * nv + 4 * nc
    b_tags_facets = 4 * nvpc * nc + 4 * nf            # c2f (== f2c in size) + facet tags out
    geo = 4 * nvpc * na if nd != nvpc else 0          # P2: cell -> vertex for the geometry besides the dofmap
    b_asm = (4 * nd * na + geo + 8 * gdim * nva + 8 * ndof_a + 8 * ndof_a + 4 * na + 8 * ng + 12 * nnz
             + 4 * (nrow + 1) + 8 * nrow)
    # the numeric cell kernel alone: no column indices / indptr (they belong to the symbolic phase)
    b_cells_kernel = 4 * nd * na + geo + 8 * gdim * nva + 16 * ndof_a + 4 * na + 8 * nnz + 8 * ndof_a
    return {"tags_cells": b_tags_cells, "tags_facets": b_tags_facets, "assembly": b_asm,
            "cells_kernel": b_cells_kernel, "total": b_tags_cells + b_tags_facets + b_asm}


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle port on the host cores (cpu_baseline of our line, and --impl reference)
# ------------------------------------------------------------------------------------------------
class CpuWorkload:
    def __init__(self, n):
        import torch
        from oracle import native as ON
        from phifem_b200 import assemble, synthetic
        from phifem_b200.mesh import MeshTags
        self.ON, self.n = ON, n
        mesh = synthetic.box_mesh(n, device="cpu")
        self.x = mesh.x.numpy()
        self.cells = np.ascontiguousarray(mesh.cells.numpy())
        self.c2f = np.ascontiguousarray(mesh.c2f.numpy())
        self.f2c = np.ascontiguousarray(mesh.f2c.numpy())
        self.phi = synthetic.sphere_levelset(mesh.x).numpy()
        self.f = synthetic.ball_source(mesh.x).numpy()
        ct = ON.tag_cells_p1(self.x, self.cells, self.phi)
        ft = ON.tag_facets_p1(self.x, self.cells, self.c2f, self.f2c, self.phi, ct)
        # ds(100) entities + symbolic phase through the product's host plumbing on CPU tensors
        from oracle import tags as OT
        ents = OT.integration_entities(self.c2f, self.f2c, (ct == 1) | (ct == 2), ft == 4)
        plan = assemble.build_plan(mesh, MeshTags(mesh, 3, torch.from_numpy(ct)),
                                   MeshTags(mesh, 2, torch.from_numpy(ft)), ents)
        self.plan = {k: np.ascontiguousarray(getattr(plan, k).numpy())
                     for k in ("active", "slots_cells", "entities", "slots_boundary", "ghost", "slots_ghost")}
        self.nnz = plan.nnz
        self.num_cells = mesh.num_cells

    def step(self):
        ON, p = self.ON, self.plan
        ct = ON.tag_cells_p1(self.x, self.cells, self.phi)
        ON.tag_facets_p1(self.x, self.cells, self.c2f, self.f2c, self.phi, ct)
        ON.assemble_p1(self.x, self.cells, self.c2f, self.f2c, self.phi, self.f, ct, p["active"],
                       p["slots_cells"], p["entities"], p["slots_boundary"], p["ghost"],
                       p["slots_ghost"], 1.0, self.nnz)


def cpu_measure(n, steps, warmup):
    from oracle import native as ON
    w = CpuWorkload(n)
    for _ in range(warmup):
        w.step()
    t0 = time.perf_counter()
    for _ in range(steps):
        w.step()
    dt = (time.perf_counter() - t0) / steps
    return {"value": w.num_cells / dt, "unit": UNIT, "cores": ON.num_threads(), "kind": "port",
            "sample": "n=%d Kuhn unit cube (%d tetrahedra), sphere level set, tags + CSR assembly, "
                      "C/OpenMP oracle port; dolfinx/PETSc is not installable here" % (n, w.num_cells),
            "ms_per_step": dt * 1e3}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        # torchrun pins OMP_NUM_THREADS=1 for its workers; the CPU arm runs on rank 0 alone and uses every
        # host core (set before libgomp is loaded)
        os.environ["OMP_NUM_THREADS"] = str(os.cpu_count() or 1)
    res = cpu_measure(args.cpu_n, max(1, args.steps), max(1, args.warmup))
    line = {"metric": METRIC, "value": res["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": max(1, args.steps), "warmup": max(1, args.warmup), "ms_per_step": res["ms_per_step"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "impl": "reference",
            "config": {"workload": "synthetic 3D P1 phi-FEM Poisson, Kuhn tetrahedra, sphere level set "
                                   "(bounded sample n=%d of the n=%d configuration)" % (args.cpu_n, args.n)},
            "cpu_baseline": {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": res["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist

    from phifem_b200 import assemble, fem, mesh_scripts, synthetic

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the hot path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    n = args.n
    reorder_ms = None
    t_setup = time.perf_counter()
    if world > 1:
        from phifem_b200 import dist as pdist
        problem = pdist.SlabProblem(n, rank, world, dev, mode=args.dist_mode)
        mesh, phi, f = problem.mesh, problem.phi, problem.f
    else:
        problem = None
        if args.config.startswith("2d"):
            mesh = synthetic.rectangle_mesh(n, device=dev)
            ls_kw = dict(center=DISC_CENTER, radius=DISC_RADIUS)
        else:
            mesh = synthetic.box_mesh(n, device=dev)
            ls_kw = {}
        if args.mesh == "unstructured":
            # vertex jitter +-0.2 h (seed 0), random cell permutation (seed 1), random vertex
            # relabelling (seed 2); then -- as part of the mesh-level symbolic phase, like dolfinx's own reordering at
            # mesh creation -- renumbered along the Morton curve.  Level set and source are interpolated on THAT mesh.
            mesh = synthetic.unstructured_variant_device(mesh, jitter=0.2, seed=0)
            torch.cuda.synchronize()
            t_re = time.perf_counter()
            if not args.no_reorder:
                mesh = mesh.reordered()
            torch.cuda.synchronize()
            reorder_ms = (time.perf_counter() - t_re) * 1e3
        phi = synthetic.sphere_levelset(mesh.x, **ls_kw)
        f = synthetic.ball_source(mesh.x, **({"center": DISC_CENTER} if ls_kw else {}))
    mesh.c2f  # build the facet topology (mesh-level symbolic, once per mesh)
    mesh.detj_bounds()
    torch.cuda.synchronize()
    topo_s = time.perf_counter() - t_setup

    degree = 2 if args.config.endswith("p2") else 1
    phi_asm, f_asm = phi, f
    V = fem.functionspace_p1_device(mesh)
    fn = fem.Function(V, phi)
    if args.detection_degree > 1:
        # tags from a P_k level set with detection_degree k (the generic table-driven classifier) instead of
        # the demos' P1 detection level set
        Vd = fem.functionspace(mesh, args.detection_degree)
        phi_det = synthetic.sphere_levelset(Vd.dof_coordinates_dev(), **ls_kw)
        dls = mesh_scripts._DeviceLevelset(mesh, fem.Function(Vd, phi_det), args.detection_degree)
    else:
        dls = mesh_scripts._DeviceLevelset(mesh, fn, 1)
    ws = mesh_scripts.TagWorkspace(mesh)
    if problem is not None:
        problem.classify(dls, ws)
    else:
        mesh_scripts.classify(mesh, dls, ws=ws)
    torch.cuda.synchronize()
    counters = ws.counters.cpu().numpy()

    t0 = time.perf_counter()
    tdim = mesh.topology.dim
    from phifem_b200.mesh import MeshTags
    if problem is not None:
        plan = problem.build_plan(ws.cell_tags8, ws.facet_tags8)
        if args.dist_mode == "exchange":
            plan.method, plan.blocked, plan.rowsplan = "atomic", None, None
        data, b = problem.data, problem.b_local
    else:
        ctags, ftags = MeshTags(mesh, tdim, ws.cell_tags), MeshTags(mesh, tdim - 1, ws.facet_tags)
        ctags.tags8, ftags.tags8 = ws.cell_tags8, ws.facet_tags8
        ents = mesh_scripts._integration_entities_dev(mesh, ws.cell_tags8, ws.facet_tags8, 4, (1, 2))
        if degree == 2:
            # P2 trial/test space and P2 level set: phi_h / f_h = interpolants at the P2 nodes (main.py:85-90)
            Vw = fem.functionspace(mesh, 2)
            Xd = Vw.dof_coordinates_dev()
            phi_asm = synthetic.sphere_levelset(Xd, **ls_kw)
            f_asm = synthetic.ball_source(Xd, **({"center": DISC_CENTER} if ls_kw else {}))
            del Xd
            plan = assemble.build_plan(mesh, ctags, ftags, ents, V=Vw, V_phi=Vw)
        else:
            plan = assemble.build_plan(mesh, ctags, ftags, ents, method=args.scatter, capacity=args.capacity,
                                       order=args.order, geometry=args.geometry, cell_pass=args.cell_pass,
                                       rows_per_tile=args.rows_per_tile)
        data, b = plan.new_outputs()
    torch.cuda.synchronize()
    symbolic_ms = (time.perf_counter() - t0) * 1e3

    def step(events=None, split=False):
        """One pass of the hot path.  events: CUDA events recorded at the phase boundaries.  split=True (the
        untimed breakdown loop) runs the assembly pass by pass so that events separate its kernels."""
        k = 0

        def mark():
            nonlocal k
            if events is not None:
                events[k].record()
                k += 1
        mark()
        if problem is not None:   # the all-reduce of "any exterior cell" overlaps the interior-facet kernel
            problem.classify(dls, ws, mark=mark)
        else:
            mesh_scripts.classify_cells(mesh, dls, ws)
            mark()
            mesh_scripts.classify_facets(mesh, dls, ws)
            mark()
        if problem is not None:
            problem.assemble(1.0, marks=mark if split else None)
        else:
            assemble.assemble_into(plan, phi_asm, f_asm, 1.0, data, b, marks=mark if split else None)
        mark()

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(args.steps)]
    sampler = ClockSampler(local_rank)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    sampler.start()
    # an NVML query takes milliseconds, a step two: keep the device under the same load for a fixed number of
    # untimed steps (the same on every rank: a step holds a collective) so that the sampler has seen it, then time
    # exactly `steps` steps with the sampler still running
    for _ in range(15):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record()
    for i in range(args.steps):
        step(evs[i])
    stop.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    clocks = sampler.stop()
    total_ms = start.elapsed_time(stop)
    if world > 1:
        t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    ms_per_step = total_ms / args.steps

    # phase durations from the events recorded inside the timed region (the assembly is ONE call there: its
    # facet-once kernel overlaps the cell pass on a side stream) ...
    per = {nm: statistics.mean(evs[i][j].elapsed_time(evs[i][j + 1]) for i in range(args.steps))
           for j, nm in enumerate(["tag_cells", "tag_facets", "assembly"])}
    # ... and the assembly pass by pass from a short untimed loop
    n_split = 5
    evs2 = [[torch.cuda.Event(enable_timing=True) for _ in range(8)] for _ in range(n_split)]
    for i in range(n_split):
        step(evs2[i], split=True)
    torch.cuda.synchronize()
    names = ["tag_cells", "tag_facets", "zero", "assemble_cells", "assemble_boundary", "assemble_ghost", "exchange"]
    for j, nm in enumerate(names):
        if j >= 2:
            per[nm] = statistics.mean(evs2[i][j].elapsed_time(evs2[i][j + 1]) for i in range(n_split))
    if plan.method == "rows":
        per["assemble_surface"] = per.pop("assemble_ghost")      # ghost penalty + one-sided term, one pass
        per.pop("assemble_boundary")

    n_cells_local = problem.n_owned_cells if problem is not None else mesh.num_cells
    n_cells_total = n_cells_local
    if world > 1:
        t = torch.tensor([n_cells_local], dtype=torch.int64, device=dev)
        dist.all_reduce(t)
        n_cells_total = int(t.item())
    value = n_cells_total / (ms_per_step * 1e-3)

    act_v = torch.zeros(mesh.num_vertices, dtype=torch.bool, device=dev)
    act_v[mesh.cells[plan.active.long()].long().reshape(-1)] = True
    counts = {"Nc": mesh.num_cells, "Nv": mesh.num_vertices, "Nf": mesh.num_facets, "gdim": mesh.gdim,
              "nvpc": mesh.cells.shape[1], "Na": int(plan.active.numel()), "Ng": int(plan.ghost.numel()),
              "Nv_active": int(act_v.sum()), "nnz": plan.nnz, "Nrow": plan.n_rows,
              "nd": getattr(plan, "nd", mesh.cells.shape[1]),
              "Ndof_active": int((plan.indptr[1:] > plan.indptr[:-1]).sum()),
              "Ne_ds100": int(plan.entities.shape[0]),
              "halo_entries_sent": (sum(hi - lo for lo, hi in plan.send_ranges) if problem is not None else 0),
              "interior": int(counters[0]), "cut": int(counters[1]), "exterior": int(counters[2])}
    ab = algorithmic_bytes(counts)
    peak, peak_src = _peaks()
    dominant = max(("tag_cells", "tag_facets", "assemble_cells"), key=lambda k_: per[k_])
    # the row-gather kernel is the whole assembly (cells + ghost + one-sided terms, pattern read, CSR
    # values and b written): SURVEY.md 8(d) B_asm; the atomic cell kernel alone moves less
    asm_bytes = ab["assembly"] if plan.method in ("rows", "blocked") else ab["cells_kernel"]
    kbytes = {"tag_cells": ab["tags_cells"], "tag_facets": ab["tags_facets"],
              "assemble_cells": asm_bytes}[dominant]
    achieved = kbytes / (per[dominant] * 1e-3) / 1e9
    kernel_names = {"tag_cells": "k_tag_cells_p1", "tag_facets": "k_tag_facets",
                    "assemble_cells": {"rows": "k_assemble_rows_p1", "blocked": "k_assemble_blocked_p1",
                                       "atomic": "k_assemble_cells_p1",
                                       "pk-atomic": "k_assemble_cells_pk"}[plan.method]}
    roofline = {"bound": "hbm", "kernel": kernel_names[dominant], "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": None, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": kbytes,
                "step_achieved_gbs": ab["total"] / (ms_per_step * 1e-3) / 1e9,
                "step_frac": ab["total"] / (ms_per_step * 1e-3) / 1e9 / peak,
                "kernels_ms": per,
                "kernels_gbs": {"tag_cells": ab["tags_cells"] / per["tag_cells"] / 1e6,
                                "tag_facets": ab["tags_facets"] / per["tag_facets"] / 1e6,
                                "assemble_cells": asm_bytes / per["assemble_cells"] / 1e6}}
    if plan.method == "rows" and mesh.gdim == 3:
        # the cell pass is bound by fp64 issue, not by HBM: 146 fp64 instructions per (row, cell) record (SASS
        # count of cell_row<3>) against 64 fp64 lanes per clock per SM
        # from the coordinates; 85 with the cached cell geometry (cell_row_geom<3>, cut-only instructions excluded)
        per_record = 85 if plan.rowsplan.cell_geom is not None else 146
        tiles = plan.rowsplan.tiles
        if tiles is not None:   # cell-once pass: ~200 fp64 instructions per cell evaluation (SASS of cell_tensor<3>)
            per_record = 200
        n_eval = tiles.n_cell_slots if tiles is not None else plan.rowsplan.cells.n_records
        lanes = n_eval * float(per_record)
        peak_lanes = 148 * 64 * (clocks["sm_max_mhz"] or 1965) * 1e6
        roofline["fp64_issue"] = {"kernel": "k_assemble_tiles_p1" if tiles is not None else "k_assemble_rows_p1<cells>",
                                  "fp64_instructions_per_record": per_record, "records": n_eval,
                                  "achieved_tera_lane_instr_per_s": lanes / (per["assemble_cells"] * 1e-3) / 1e12,
                                  "peak_tera_lane_instr_per_s": peak_lanes / 1e12,
                                  "frac": lanes / (per["assemble_cells"] * 1e-3) / peak_lanes}
    traffic_file = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(traffic_file) and args.config == "3d-p1" and n == CONFIGS["3d-p1"][0] and world == 1:
        with open(traffic_file) as fh:
            roofline["traffic"] = json.load(fh).get(dominant)

    # ---- e2e: public API, pinned host buffers in, pinned host buffers out ---------------------------
    e2e = None
    if world == 1 and not args.no_e2e:
        phi_h = phi.cpu().pin_memory()
        f_h = f_asm.cpu().pin_memory()
        phi_asm_h = phi_asm.cpu().pin_memory() if degree == 2 else phi_h
        out_h = {"ct": torch.empty(mesh.num_cells, dtype=torch.int8).pin_memory(),
                 "ft": torch.empty(mesh.num_facets, dtype=torch.int8).pin_memory(),
                 "data": torch.empty(plan.nnz, dtype=torch.float64).pin_memory(),
                 "b": torch.empty(plan.n_rows, dtype=torch.float64).pin_memory()}
        import warnings
        side = torch.cuda.Stream()
        tags_done = torch.cuda.Event()
        f_done = torch.cuda.Event()

        def e2e_step():
            # host level set -> device once; the same device-resident Function feeds the tags and (for P1) the
            # assembly, as a user holding one phi_h would write it
            phi_d = phi_h.to(dev, non_blocking=True)
            # the source term follows the level set over PCIe on the side stream while the tag kernels run
            with torch.cuda.stream(side):
                f_d = f_h.to(dev, non_blocking=True)
                f_done.record()
            fn_h = fem.Function(V, phi_d)
            with warnings.catch_warnings():
                warnings.simplefilter("ignore", RuntimeWarning)
                ct_, ft_, _, ds_, _ = mesh_scripts.compute_tags_measures(mesh, fn_h, 1, box_mode=True)
            # the tags (one byte per entity, as the kernels write them; MeshTags widens to int32 on the host)
            # leave on a side stream while the assembly runs
            tags_done.record()
            with torch.cuda.stream(side):
                side.wait_event(tags_done)
                out_h["ct"].copy_(ct_.tags8, non_blocking=True)
                out_h["ft"].copy_(ft_.tags8, non_blocking=True)
            torch.cuda.current_stream().wait_event(f_done)
            A_, b_ = assemble.assemble_strong_dirichlet(plan, phi_d if degree == 1 else phi_asm_h, f_d,
                                                        stab_coef=1.0)
            f_d.record_stream(torch.cuda.current_stream())
            out_h["data"].copy_(A_.data, non_blocking=True)
            out_h["b"].copy_(b_, non_blocking=True)
            torch.cuda.synchronize()

        def e2e_pipelined(n_steps):
            """The same calls with two steps in flight: step i runs on stream i % 2 with its own pinned output
            buffers, so that the uploads, tag kernels and assembly of step i + 1 proceed under the device->host
            copies of step i (PCIe is full duplex; the D2H direction bounds the step).  One host thread; every step
            still uploads its inputs and downloads its results inside the timed region."""
            mains = [torch.cuda.Stream(), torch.cuda.Stream()]
            sides = [torch.cuda.Stream(), torch.cuda.Stream()]
            outs = [out_h, {k: torch.empty_like(v).pin_memory() for k, v in out_h.items()}]
            done, keep, asm_done = [None, None], [None, None], None
            torch.cuda.synchronize()
            t_start = time.perf_counter()
            for i in range(n_steps):
                j = i & 1
                if done[j] is not None:      # results of step i - 2 are on the host: its buffers may be reused
                    done[j].synchronize()
                    keep[j] = None
                s, sd, o = mains[j], sides[j], outs[j]
                with torch.cuda.stream(s):
                    phi_d = phi_h.to(dev, non_blocking=True)
                    f_ev, t_ev = torch.cuda.Event(), torch.cuda.Event()
                    with torch.cuda.stream(sd):
                        f_d = f_h.to(dev, non_blocking=True)
                        f_ev.record()
                    fn_h = fem.Function(V, phi_d)
                    with warnings.catch_warnings():
                        warnings.simplefilter("ignore", RuntimeWarning)
                        ct_, ft_, _, ds_, _ = mesh_scripts.compute_tags_measures(mesh, fn_h, 1, box_mode=True)
                    t_ev.record()
                    with torch.cuda.stream(sd):
                        sd.wait_event(t_ev)
                        o["ct"].copy_(ct_.tags8, non_blocking=True)
                        o["ft"].copy_(ft_.tags8, non_blocking=True)
                    s.wait_event(f_ev)
                    if asm_done is not None:  # the plan's facet-once scratch is shared by consecutive assemblies
                        s.wait_event(asm_done)
                    A_, b_ = assemble.assemble_strong_dirichlet(plan, phi_d if degree == 1 else phi_asm_h, f_d,
                                                                stab_coef=1.0)
                    asm_done = torch.cuda.Event()
                    asm_done.record()
                    o["data"].copy_(A_.data, non_blocking=True)
                    o["b"].copy_(b_, non_blocking=True)
                    s.wait_stream(sd)
                    done[j] = torch.cuda.Event()
                    done[j].record()
                    keep[j] = (phi_d, f_d, ct_, ft_, A_, b_)
            torch.cuda.synchronize()
            elapsed = time.perf_counter() - t_start
            for k in out_h:                   # both buffer sets hold the same results (bit for bit where the
                if plan.method == "rows" or outs[0][k].dtype != torch.float64:   # assembly sums in a fixed order)
                    assert torch.equal(outs[0][k], outs[1][k]), k
                else:
                    assert torch.allclose(outs[0][k], outs[1][k], rtol=1e-10, atol=1e-12 * float(outs[0][k].abs().max())), k
            return elapsed / n_steps

        e2e_steps = max(2, min(args.steps, 5))
        for _ in range(2):
            e2e_step()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            e2e_step()
        dt = (time.perf_counter() - t0) / e2e_steps
        pipelined_dt = None
        if args.e2e_pipeline:
            ref_data, ref_b = out_h["data"].clone(), out_h["b"].clone()
            e2e_pipelined(4)
            pipelined_dt = e2e_pipelined(2 * e2e_steps)
            if plan.method == "rows":
                assert torch.equal(out_h["data"], ref_data) and torch.equal(out_h["b"], ref_b)
        e2e = {"value": mesh.num_cells / dt, "unit": UNIT, "ms_per_step": dt * 1e3,
               "h2d_bytes_per_step": int(phi_h.numel() * 8 + (phi_asm_h.numel() * 8 if degree == 2 else 0)
                                         + f_h.numel() * 8),
               "d2h_bytes_per_step": int(sum(t.numel() * t.element_size() for t in out_h.values())),
               "api": "compute_tags_measures(box_mode=True) + assemble_strong_dirichlet(plan, ...) with "
                      "pinned host level set / source in and pinned host tags (1 byte per cell / facet) + CSR values "
                      "+ b out, source-term upload overlapped with the tag kernels, tag copies with the assembly; "
                      "assembly plan (symbolic phase) reused; one step at a time"}
        if pipelined_dt is not None:
            e2e["two_steps_in_flight_ms"] = pipelined_dt * 1e3

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu and args.config == "3d-p1":
        cpu = cpu_measure(args.cpu_n, 3, 1)
        cpu.pop("ms_per_step")

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": (CONFIGS[args.config][1] + "%s, tags + strong-Dirichlet CSR assembly")
                                       % (n_cells_local, n, " per unit cube joined by a thin tube across the "
                                          "partition boundaries" if world > 1 else ""),
                           "name": args.config, "mesh": args.mesh + (
                               "" if args.mesh == "structured" else
                               (": jitter 0.2 h, cells permuted, vertices relabelled" +
                                (", NOT renumbered" if args.no_reorder else ", renumbered along the Morton curve"))),
                           "cells_total": n_cells_total, "counts": counts,
                           "l2_policy": "inputs larger than L2 (%.1f GB streamed per step)"
                                        % (ab["total"] / 1e9),
                           "partition": ("single GPU" if world == 1 else
                                         "1 slab of the global box per rank, owned CSR rows, "
                                         + ("owner computes its rows from 2 redundantly classified ghost "
                                            "layers: one 8-byte all-reduce per step, no halo exchange"
                                            if args.dist_mode == "rows" else "NCCL halo exchange")),
                           "tags_dtype": "int8 on the device (int32 MeshTags.values widened on demand)",
                           "timed": "tag kernels + zeroing + assembly kernels; symbolic phase excluded; "
                                    "kernels_ms: tag_* and assembly from events inside the timed region, "
                                    "the assemble_* split from an untimed pass-by-pass loop"},
                "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "clocks": clocks,
                "gpu_launches": {"rows": 7, "blocked": 5, "atomic": 7, "pk-atomic": 9}[plan.method] * args.steps,
                "symbolic_ms": symbolic_ms, "topology_s": topo_s, "reorder_ms": reorder_ms,
                "scatter": {"method": plan.method,
                            **({"blocks": plan.blocked.n_blocks, "capacity": plan.blocked.capacity,
                                "bin_shape": plan.blocked.bin_shape,
                                "recompute_factor": plan.blocked.redundancy,
                                "plan_bytes": plan.blocked.index_bytes()} if plan.blocked else {}),
                            **({"order": plan.rowsplan.order, "max_row_nnz": plan.rowsplan.max_row_nnz,
                                "cell_pass": plan.rowsplan.cell_pass,
                                **({"rows_per_tile": plan.rowsplan.tiles.rows_per_tile,
                                    "tiles": plan.rowsplan.tiles.n_tiles, "chunks": plan.rowsplan.tiles.n_chunks,
                                    "cell_evaluations": plan.rowsplan.tiles.n_cell_slots,
                                    "recompute_factor": plan.rowsplan.tiles.recompute}
                                   if plan.rowsplan.tiles is not None else {}),
                                "rows_cells_surface": [plan.rowsplan.tiles.n_listed if plan.rowsplan.tiles is not None
                                                       else plan.rowsplan.cells.n_listed,
                                                       plan.rowsplan.surface.n_listed],
                                "records_cells_ghost_onesided": [plan.rowsplan.n_cell_records,
                                                                 plan.rowsplan.n_ghost_records,
                                                                 plan.rowsplan.n_entity_records],
                                "lane_padding": [plan.rowsplan.cells.padding(), plan.rowsplan.surface.padding()],
                                "cell_geometry": "cached per plan (64 B per active cell)"
                                if plan.rowsplan.cell_geom is not None else "from the vertex coordinates",
                                "plan_bytes": plan.rowsplan.index_bytes()} if plan.rowsplan else {})}}
        emit(line)
    if world > 1:
        dist.destroy_process_group()


_JSON_FD = None


def _quiet_stdout():
    """Library chatter (e.g. "NCCL version ..." from libnccl) must not share stdout with the ONE JSON line:
    fd 1 is pointed at stderr for the run and the line is written to the saved descriptor."""
    global _JSON_FD
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)


def emit(line):
    payload = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(payload.decode())
        sys.stdout.flush()
    else:
        sys.stdout.flush()
        os.write(_JSON_FD, payload)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--config", default="3d-p1", choices=sorted(CONFIGS),
                    help="workload; the default is the configuration BASELINE.json's metric is quoted on")
    ap.add_argument("--n", type=int, default=None,
                    help="cells per edge (6 n^3 tetrahedra per GPU / 2 n^2 triangles); default per config")
    ap.add_argument("--cpu-n", type=int, default=80, help="size of the bounded CPU sample")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--scatter", default="rows", choices=["rows", "blocked", "atomic"],
                    help="assembly strategy (row-gather / owner-computes blocks / fp64 reductions)")
    ap.add_argument("--capacity", type=int, default=None, help="contributions per block (blocked scatter)")
    ap.add_argument("--mesh", default="structured", choices=["structured", "unstructured"],
                    help="unstructured = SURVEY.md 8(d)'s variant (jitter, cell permutation, vertex relabelling), "
                         "renumbered along the Morton curve in the mesh-level symbolic phase")
    ap.add_argument("--no-reorder", action="store_true", help="unstructured mesh left in its random numbering")
    ap.add_argument("--cell-pass", default="rows", choices=["rows", "tiles"],
                    help="row-gather cell pass / cell-once tile pass (csrc/assemble_tiles.cu)")
    ap.add_argument("--rows-per-tile", type=int, default=256, choices=[128, 256])
    ap.add_argument("--order", default="auto", choices=["auto", "natural", "morton"],
                    help="row processing order of the row-gather assembly")
    ap.add_argument("--e2e-pipeline", action="store_true",
                    help="also measure e2e with two steps in flight (uploads and kernels of step i + 1 under the "
                         "downloads of step i): reported as e2e.two_steps_in_flight_ms, the e2e value stays the "
                         "one-step-at-a-time figure (the pipelined one varies from box to box: 14.1 / 20.5 ms seen)")
    ap.add_argument("--geometry", action="store_true",
                    help="row-gather cell pass from a per-plan geometry table instead of the vertex coordinates "
                         "(measured slower: profiles/round2_a_geometry_kernel.md)")
    ap.add_argument("--dist-mode", default="rows", choices=["rows", "exchange"],
                    help="multi-GPU numeric strategy: owner-computes rows / per-entity kernels + halo exchange")
    ap.add_argument("--detection-degree", type=int, default=1, choices=[1, 2],
                    help="degree of the detection level set and of the detection rule (1 = the demos' setting)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true", help="skip the end-to-end leg (profiling runs)")
    args = ap.parse_args()
    if args.n is None:
        args.n = CONFIGS[args.config][0]
    if args.config != "3d-p1" and (args.gpus > 1 or int(os.environ.get("WORLD_SIZE", "1")) > 1):
        raise SystemExit("bench.py: the multi-GPU path runs the 3d-p1 configuration")
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    _quiet_stdout()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
