"""Times the fused LowRes kernels on a B200 (device-resident batches) under different benchmark knobs.
Usage: python tools/sweep_lowres.py <workload,...> <name:KNOB=V,KNOB=V> ...   (see WORKLOADS / KNOBS below)"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from robust_object_detection_b200.batch import CorruptionPlan

torch.cuda.set_device(0)
POOL = [(765, 1360), (1050, 1400), (788, 1400), (1078, 1916), (1080, 1920), (1500, 2000), (540, 960), (360, 480),
        (765, 1361), (1079, 1917), (1499, 1999)]


def rate(fn, nbytes, steps=8, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return round(nbytes / (e0.elapsed_time(e1) / steps) / 1e6, 1)


WORKLOADS = {
    "1360x765_x256": [(765, 1360)] * 256,
    "1360x765_x64": [(765, 1360)] * 64,
    "1916x1079_x64": [(1079, 1916)] * 64,
    "1361x765_x128": [(765, 1361)] * 128,
    "1999x1499_x48": [(1499, 1999)] * 48,
    "1917x1080_x64": [(1080, 1917)] * 64,
    "1920x1080_x128": [(1080, 1920)] * 128,
    "1400x1050_x96": [(1050, 1400)] * 96,
    "1400x788_x128": [(788, 1400)] * 128,
    "1916x1078_x64": [(1078, 1916)] * 64,
    "2000x1500_x48": [(1500, 2000)] * 48,
    "960x540_x256": [(540, 960)] * 256,
    "480x360_x512": [(360, 480)] * 512,
    "visdrone_256": [POOL[i] for i in np.random.default_rng(3001).integers(0, 8, 256)],
    "mixed_256": [POOL[i] for i in np.random.default_rng(3000).integers(0, len(POOL), 256)],
    "odd_only_96": [POOL[8 + i % 3] for i in range(96)],
}
KNOBS = ("ROD_X2_PACKED", "ROD_X2P_CTAS", "ROD_X2_FLOAT_STAGED", "ROD_X2F_CTAS", "ROD_X2_ODD_STAGED", "ROD_X2G_CTAS",
         "ROD_X2G_BAND_DIV", "ROD_X2_REGULAR", "ROD_X2H_CTAS", "ROD_X2_BAND", "ROD_X2_ODD_REGULAR", "ROD_X2I_CTAS",
         "ROD_X2I_GRID", "ROD_X2P_GRID", "ROD_X2F_GRID", "ROD_X2H_GRID", "ROD_LR_LANES", "ROD_LR_X2I_LAST", "ROD_LR_HELPERS")

# usage: sweep_lowres.py <workload,workload,...> <name:K=V,K=V> <name:K=V> ...   (a bare "name:" is the default setting)
names = sys.argv[1].split(",") if len(sys.argv) > 1 else list(WORKLOADS)
settings = []
for a in sys.argv[2:]:
    sname, _, kvs = a.partition(":")
    settings.append((sname, dict(kv.split("=") for kv in kvs.split(",") if kv)))
if not settings:
    settings = [("default", {})]
out = {}
for wname in names:
    shapes = WORKLOADS[wname]
    src = dst = None
    for sname, env in settings:
        for k in KNOBS:
            os.environ.pop(k, None)
        os.environ.update(env)
        plan = CorruptionPlan.ragged(shapes)
        if src is None:
            src = torch.randint(0, 256, (plan.src_bytes,), dtype=torch.uint8, device="cuda")
            dst = torch.empty_like(src)
        out[f"{wname}:{sname}"] = rate(lambda: plan.lowres(src, dst), 2 * plan.payload_bytes)
        del plan
    del src, dst
    print(json.dumps(out), flush=True)
