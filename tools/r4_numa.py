"""Experiment: where do the host cores / memory of the box sit relative to the GPU?  (affinity masks, NUMA nodes, D2H rate of a
pinned buffer allocated under the GPU's own CPU affinity against the default)"""
import os
import time

import pynvml
import torch

pynvml.nvmlInit()
h = pynvml.nvmlDeviceGetHandleByIndex(0)
print("process affinity", sorted(os.sched_getaffinity(0)))
try:
    words = pynvml.nvmlDeviceGetCpuAffinity(h, 4)
    cpus = [64 * i + b for i, w in enumerate(words) for b in range(64) if (w >> b) & 1]
    print("gpu0 cpu affinity", cpus)
except Exception as exc:  # noqa: BLE001
    cpus = []
    print("nvmlDeviceGetCpuAffinity failed:", exc)
for node in sorted(os.listdir("/sys/devices/system/node")) if os.path.isdir("/sys/devices/system/node") else []:
    if node.startswith("node"):
        print(node, open("/sys/devices/system/node/%s/cpulist" % node).read().strip())
os.system("nvidia-smi topo -m 2>/dev/null | head -12")


def rate(label):
    dev = torch.device("cuda", 0)
    src = torch.empty(400 * 1024 * 1024, dtype=torch.uint8, device=dev)
    dst = torch.empty(src.numel(), dtype=torch.uint8).pin_memory()
    up = torch.empty(128 * 1024 * 1024, dtype=torch.uint8).pin_memory()
    upd = torch.empty(up.numel(), dtype=torch.uint8, device=dev)
    side = torch.cuda.Stream()
    for both in (False, True):
        for _ in range(2):
            dst.copy_(src, non_blocking=True)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(5):
            dst.copy_(src, non_blocking=True)
            if both:
                with torch.cuda.stream(side):
                    upd.copy_(up, non_blocking=True)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / 5
        print("%s: D2H %.1f GB/s%s" % (label, src.numel() / dt / 1e9, " with a concurrent upload" if both else ""))


rate("default affinity")
allowed = sorted(set(cpus) & os.sched_getaffinity(0))
if allowed and set(allowed) != os.sched_getaffinity(0):
    os.sched_setaffinity(0, allowed)
    rate("gpu-local affinity %s" % allowed)
else:
    print("the GPU's affinity mask does not narrow the process's: nothing to bind")
