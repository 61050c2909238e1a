"""D2H rate of pinned copies: one stream against two / four concurrent streams (is 45 GB/s the link or the engine?)."""
import time
import torch

dev = torch.device("cuda:0")
n = 409_000_000 // 8
src = torch.rand(n, dtype=torch.float64, device=dev)
dst = torch.empty(n, dtype=torch.float64).pin_memory()
for parts in (1, 2, 4):
    streams = [torch.cuda.Stream() for _ in range(parts)]
    chunks = [(i * n // parts, (i + 1) * n // parts) for i in range(parts)]
    best = 1e9
    for rep in range(5):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for s, (a, b) in zip(streams, chunks):
            with torch.cuda.stream(s):
                dst[a:b].copy_(src[a:b], non_blocking=True)
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    print("d2h %d stream(s): %.2f ms  %.1f GB/s" % (parts, best * 1e3, n * 8 / best / 1e9))
h = torch.rand(69_000_000 // 8, dtype=torch.float64).pin_memory()
d = torch.empty_like(h, device=dev)
best = 1e9
for rep in range(5):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    best = min(best, time.perf_counter() - t0)
print("h2d 69 MB: %.2f ms  %.1f GB/s" % (best * 1e3, h.numel() * 8 / best / 1e9))
# both directions at once
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
best = 1e9
for rep in range(5):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    with torch.cuda.stream(s1):
        dst.copy_(src, non_blocking=True)
    with torch.cuda.stream(s2):
        d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    best = min(best, time.perf_counter() - t0)
print("d2h 409 MB + h2d 69 MB concurrently: %.2f ms" % (best * 1e3))
