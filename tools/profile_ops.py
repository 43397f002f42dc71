"""Short driver for ncu captures: a few launches of each kernel on a 64-image 1360x765 batch
(399 MB of traffic per launch, larger than L2).  Not a benchmark."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from robust_object_detection_b200.batch import CorruptionPlan, draw_decisions

n, h, w = int(os.environ.get("ROD_PROFILE_N", "64")), 765, 1360
torch.cuda.set_device(0)
src = torch.randint(0, 256, (n, h, w, 3), dtype=torch.uint8, device="cuda")
dst = torch.empty_like(src)
plan = CorruptionPlan.uniform(n, h, w)
which = sys.argv[1:] or ["blur", "noise", "lowres", "letterbox", "mixed"]
for _ in range(3):
    if "blur" in which:
        plan.blur(src, dst)
    if "noise" in which:
        plan.noise(src, dst, None, 15.0, seed=1)
    if "lowres" in which:
        plan.lowres(src, dst)
if "lowres1080" in which:   # even x even shape: the packed-integer kernel
    p2 = CorruptionPlan.uniform(48, 1080, 1920)
    s2 = torch.randint(0, 256, (48, 1080, 1920, 3), dtype=torch.uint8, device="cuda")
    d2 = torch.empty_like(s2)
    for _ in range(3):
        p2.lowres(s2, d2)
    del s2, d2
if "lowresodd" in which:   # odd width: the staged odd-width kernel
    p3 = CorruptionPlan.uniform(48, 1079, 1917)
    s3 = torch.randint(0, 256, (48, 1079, 1917, 3), dtype=torch.uint8, device="cuda")
    d3 = torch.empty_like(s3)
    for _ in range(3):
        p3.lowres(s3, d3)
    del s3, d3
if "letterbox" in which:
    import random
    random.seed(42)
    ops = torch.from_numpy(draw_decisions(n)).cuda()
    if os.environ.get("ROD_PROFILE_OP"):   # every image on the same op
        ops = torch.full((n,), int(os.environ["ROD_PROFILE_OP"]), dtype=torch.uint8, device="cuda")
    f16 = torch.empty((n, 3, 640, 640), dtype=torch.float16, device="cuda")
    for _ in range(2):
        plan.corrupt_letterbox(src, ops, f16, 640, 640, 114, seed=1)
if "jpeg" in which:   # device JPEG encoder on the 64-image batch (uniform noise: every coefficient coded)
    from robust_object_detection_b200 import _native as N
    from robust_object_detection_b200.batch import _ptr
    from robust_object_detection_b200.jpeg import JpegEncoder
    enc = JpegEncoder([(h, w)] * n, [i * 3 * h * w for i in range(n)])
    for _ in range(2):
        N.check(N.lib().rod_jpeg_encode(enc._h, _ptr(src), None), "rod_jpeg_encode")
    torch.cuda.synchronize()
    del enc
if "jpegdec" in which:   # device JPEG decoder on 64 smooth files of the batch's size (one decode: every kernel of the chain)
    import cv2
    from robust_object_detection_b200.jpeg import JpegDecoder
    rng = np.random.default_rng(3)
    files = [cv2.imencode(".jpg", cv2.GaussianBlur(rng.integers(0, 256, (h, w, 3), dtype=np.uint8), (0, 0), 3.0))[1].tobytes() for _ in range(8)] * (n // 8)
    dec = JpegDecoder(files, [i * 3 * h * w for i in range(len(files))])
    dec.decode(dst)
    assert (dec.status() == 0).all()
    del dec
if "mixed" in which:
    shapes = [(1080, 1920), (1079, 1917), (1050, 1400), (1500, 2000)] * 8
    rp = CorruptionPlan.ragged(shapes)
    rsrc = torch.randint(0, 256, (rp.src_bytes,), dtype=torch.uint8, device="cuda")
    rdst = torch.empty_like(rsrc)
    for _ in range(2):
        rp.lowres(rsrc, rdst)
        rp.blur(rsrc, rdst)
torch.cuda.synchronize()
print("profile_ops done")
