"""How fast does one B200 box take the step's inputs (c2: 587 MB of pinned host memory) over PCIe, as one copy on one
stream or split over several streams / copy engines?  (The end-to-end number of bench.py is bound by this.)"""
import time
import torch

n = 587_253_760 // 4
host = torch.empty(n, dtype=torch.float32).pin_memory()
host.normal_()
dev = torch.empty(n, dtype=torch.float32, device="cuda")
for parts in (1, 2, 3, 4, 8):
    streams = [torch.cuda.Stream() for _ in range(parts)]
    bounds = [n * i // parts for i in range(parts + 1)]

    def go():
        for i, s in enumerate(streams):
            with torch.cuda.stream(s):
                dev[bounds[i]:bounds[i + 1]].copy_(host[bounds[i]:bounds[i + 1]], non_blocking=True)
        for s in streams:
            s.synchronize()
    go()
    ts = []
    for _ in range(8):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        go()
        ts.append(time.perf_counter() - t0)
    ts.sort()
    print("%d stream(s): median %.2f ms = %.1f GB/s, best %.2f ms = %.1f GB/s" %
          (parts, ts[4] * 1e3, n * 4 / ts[4] / 1e9, ts[0] * 1e3, n * 4 / ts[0] / 1e9), flush=True)
