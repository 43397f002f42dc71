"""Device JPEG encoder alone: ms per batch and images/s on VisDrone-shaped frames (smoothed and noisy content), next to
cv2.imencode on the host cores; and a cProfile of the philox-mode test-set driver.  Usage: python tools/time_jpeg.py [n]"""
import cProfile
import io
import json
import os
import pstats
import sys
import tempfile
import time
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cv2
import numpy as np
import torch

from robust_object_detection_b200.batch import CorruptionPlan
from robust_object_detection_b200.jpeg import JpegEncoder

SHAPES = [(765, 1360), (1050, 1400), (788, 1400), (1078, 1916), (1080, 1920), (1500, 2000), (540, 960), (360, 480)]
n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
rng = np.random.default_rng(4)
shapes = [SHAPES[int(rng.integers(0, 8))] for _ in range(n)]
out = {}
for kind in ("smooth", "noise"):
    imgs = []
    for h, w in shapes:
        im = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        imgs.append(cv2.GaussianBlur(im, (0, 0), 3.0) if kind == "smooth" else im)
    plan = CorruptionPlan.ragged(shapes)
    dev = torch.from_numpy(plan.pack(imgs)).cuda()
    enc = JpegEncoder(shapes, plan.src_offsets)
    files = enc.encode(dev)
    assert all(f == cv2.imencode(".jpg", im)[1].tobytes() for f, im in zip(files[:4], imgs[:4]))
    from robust_object_detection_b200 import _native as N
    from robust_object_detection_b200.batch import _ptr
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        N.check(N.lib().rod_jpeg_encode(enc._h, _ptr(dev), None), "enc")
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    for _ in range(3):   # the three page-locked download buffers of the ring are allocated by the first calls
        enc.encode(dev, copy=False)
    t0 = time.perf_counter()
    for _ in range(3):
        enc.encode(dev, copy=False)
    t_full = (time.perf_counter() - t0) / 3
    with ThreadPoolExecutor(16) as pool:
        t0 = time.perf_counter()
        list(pool.map(lambda im: cv2.imencode(".jpg", im)[1], imgs))
        t_cv = time.perf_counter() - t0
    mb = sum(len(f) for f in files) / 1e6
    out[kind] = {"images": n, "device_ms": round(ms, 2), "device_images_per_s": round(n / ms * 1e3), "encode_call_ms": round(t_full * 1e3, 1),
                 "encode_call_images_per_s": round(n / t_full), "cv2_16_threads_ms": round(t_cv * 1e3, 1), "cv2_images_per_s": round(n / t_cv),
                 "compressed_MB": round(mb, 1), "raw_MB": round(plan.payload_bytes / 1e6, 1)}
print(json.dumps(out))

if len(sys.argv) > 2:
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import time_testset_driver as T
    base = Path(tempfile.mkdtemp(prefix="rod_prof_", dir="/dev/shm"))
    T.make_tree(base / "src", 64)
    T.run_driver(base / "src", base / "warm", "philox")
    pr = cProfile.Profile()
    pr.enable()
    T.run_driver(base / "src", base / "p", "philox")
    pr.disable()
    s = io.StringIO()
    pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(28)
    print(s.getvalue()[:6000], file=sys.stderr)
