"""Device JPEG decoder alone: host preparation, device time per kernel and images/s on VisDrone-shaped files (smoothed and
noisy content, OpenCV's default quality 95), next to cv2.imdecode on 16 host threads.  Usage: python tools/time_jpegdec.py [n]"""
import json
import os
import sys
import time
from concurrent.futures import ThreadPoolExecutor

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cv2
import numpy as np
import torch

from robust_object_detection_b200.batch import CorruptionPlan
from robust_object_detection_b200.jpeg import JpegDecoder

SHAPES = [(765, 1360), (1050, 1400), (788, 1400), (1078, 1916), (1080, 1920), (1500, 2000), (540, 960), (360, 480)]
n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
rng = np.random.default_rng(4)
shapes = [SHAPES[int(rng.integers(0, 8))] for _ in range(n)]
out = {}
for kind in ("smooth", "noise"):
    files = []
    for h, w in shapes:
        im = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        files.append(cv2.imencode(".jpg", cv2.GaussianBlur(im, (0, 0), 3.0) if kind == "smooth" else im)[1].tobytes())
    plan = CorruptionPlan.ragged(shapes)
    pix = torch.empty(plan.src_bytes, dtype=torch.uint8, device="cuda")
    dec = JpegDecoder(files, plan.src_offsets, host_threads=16)   # warm-up: context, block cache
    dec.decode(pix)
    assert (dec.status() == 0).all()
    got = pix.cpu().numpy()
    for i in range(0, n, max(1, n // 8)):
        h, w = shapes[i]
        assert np.array_equal(got[plan.src_offsets[i]:plan.src_offsets[i] + 3 * h * w].reshape(h, w, 3),
                              cv2.imdecode(np.frombuffer(files[i], np.uint8), cv2.IMREAD_COLOR)), i
    t_create, t_total = [], []
    for _ in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        dec = JpegDecoder(files, plan.src_offsets, host_threads=16)
        t1 = time.perf_counter()
        dec.decode(pix)
        dec.status()
        t_total.append(time.perf_counter() - t0)
        t_create.append(t1 - t0)
    e = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    e[0].record()
    dec.decode(pix)
    e[1].record()
    torch.cuda.synchronize()
    with ThreadPoolExecutor(16) as pool:
        t0 = time.perf_counter()
        list(pool.map(lambda f: cv2.imdecode(np.frombuffer(f, np.uint8), cv2.IMREAD_COLOR), files))
        t_cv = time.perf_counter() - t0
    out[kind] = {"images": n, "compressed_MB": round(sum(len(f) for f in files) / 1e6, 1), "raw_MB": round(plan.payload_bytes / 1e6, 1),
                 "host_prepare_ms": round(min(t_create) * 1e3, 1), "decode_call_ms": round(min(t_total) * 1e3, 1),
                 "device_ms": round(e[0].elapsed_time(e[1]), 1), "decode_call_images_per_s": round(n / min(t_total)),
                 "cv2_16_threads_ms": round(t_cv * 1e3, 1), "cv2_images_per_s": round(n / t_cv)}
print(json.dumps(out))
