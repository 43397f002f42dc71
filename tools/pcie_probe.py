"""PCIe ceiling for the e2e (host-buffer) path: pinned H2D alone, D2H alone, both at once, per chunk size."""
import sys
import time

import torch

torch.cuda.set_device(0)
tot = 768 << 20
h_in = torch.empty(tot, dtype=torch.uint8).pin_memory()
h_out = torch.empty(tot, dtype=torch.uint8).pin_memory()
d_a = torch.empty(tot, dtype=torch.uint8, device="cuda")
d_b = torch.empty(tot, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(mode, chunk):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for off in range(0, tot, chunk):
        if mode in ("h2d", "both"):
            with torch.cuda.stream(s1):
                d_a[off:off + chunk].copy_(h_in[off:off + chunk], non_blocking=True)
        if mode in ("d2h", "both"):
            with torch.cuda.stream(s2):
                h_out[off:off + chunk].copy_(d_b[off:off + chunk], non_blocking=True)
    torch.cuda.synchronize()
    return tot / (time.perf_counter() - t0) / 1e9


for chunk in (8 << 20, 32 << 20, 128 << 20, 768 << 20):
    for mode in ("h2d", "d2h", "both"):
        run(mode, chunk)
        best = max(run(mode, chunk) for _ in range(3))
        print(f"chunk {chunk >> 20:4d} MiB  {mode:5s}  {best:6.1f} GB/s per direction", flush=True)
