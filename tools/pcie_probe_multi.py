"""Host <-> device copy ceiling of the box with 1, 2, 4, ... N GPUs copying AT THE SAME TIME (pinned buffers, both
directions), to back the end-to-end numbers of bench.py at N > 1.  One rank per GPU:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 tools/pcie_probe_multi.py
Rank 0 prints one JSON line: per phase (active ranks) and mode (h2d / d2h / both) the per-rank and the aggregate GB/s."""
import json
import os
import time

import torch
import torch.distributed as dist

rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("gloo")   # host-side barriers only: the copies under test own the PCIe links
tot, chunk = 512 << 20, 32 << 20
h_in = torch.empty(tot, dtype=torch.uint8).pin_memory()
h_out = torch.empty(tot, dtype=torch.uint8).pin_memory()
h_in.fill_(rank + 1)
d_a = torch.empty(tot, dtype=torch.uint8, device="cuda")
d_b = torch.empty(tot, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def one_pass(mode):
    for off in range(0, tot, chunk):
        if mode in ("h2d", "both"):
            with torch.cuda.stream(s1):
                d_a[off:off + chunk].copy_(h_in[off:off + chunk], non_blocking=True)
        if mode in ("d2h", "both"):
            with torch.cuda.stream(s2):
                h_out[off:off + chunk].copy_(d_b[off:off + chunk], non_blocking=True)


def measure(mode, active, passes=6):
    """All active ranks copy concurrently for `passes` passes; returns this rank's GB/s per direction (0 if idle)."""
    dist.barrier()
    rate = 0.0
    if active:
        one_pass(mode)
        torch.cuda.synchronize()
    dist.barrier()
    if active:
        t0 = time.perf_counter()
        for _ in range(passes):
            one_pass(mode)
        torch.cuda.synchronize()
        rate = passes * tot / (time.perf_counter() - t0) / 1e9
    return rate


out = {"world": world, "bytes_per_pass": tot, "chunk": chunk, "cpus": os.cpu_count(), "phases": {}}
n = 1
while n <= world:
    for mode in ("h2d", "d2h", "both"):
        r = torch.tensor([measure(mode, rank < n)], dtype=torch.float64)
        rates = [torch.zeros(1, dtype=torch.float64) for _ in range(world)]
        dist.all_gather(rates, r)
        per_rank = [round(float(x), 1) for x in rates[:n]]
        out["phases"][f"{n}gpu_{mode}"] = {"per_rank_GBs": per_rank, "aggregate_GBs_per_direction": round(sum(per_rank), 1)}
    n *= 2
if rank == 0:
    print(json.dumps(out), flush=True)
dist.destroy_process_group()
