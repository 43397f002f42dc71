"""End-to-end timing of BASELINE config 5 (the augmented-dataloader path): batches of 16 host frames (1360x765 BGR uint8)
-> 50 % random one-of-three corruption -> 640x640 letterbox -> RGB /255 fp16 NCHW on the GPU.

  ours       robust_object_detection_b200.training.CorruptionBatcher.run (pinned staging, one H2D per batch, fused
             corrupt + letterbox kernel; the fp16 batch stays on the device, where the detector consumes it)
  reference  the same per-image work with the reference's own library calls on the host (oracle/cv2_port.py for the
             corruption, cv2.resize / copyMakeBorder / /255 for the detector input -- what Ultralytics' LetterBox +
             preprocess_batch do), then one H2D of the fp16 batch.  Timed in ONE process with OpenCV's thread pool, and
             scaled to the reference's 8 DataLoader workers as an upper bound (x8, assuming perfect scaling).

Usage: python tools/time_training_hook.py [n_batches] > profiles/<tag>_training_hook.json   (imports oracle/: bench infra)"""
import json
import os
import random
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import cv2  # noqa: E402
import torch  # noqa: E402

from oracle import cv2_port  # noqa: E402
from robust_object_detection_b200.batch import draw_decisions  # noqa: E402
from robust_object_detection_b200.training import CorruptionBatcher  # noqa: E402

H, W, B, OUT = 765, 1360, 16, 640


def reference_batch(frames, ops):
    out = np.empty((len(frames), 3, OUT, OUT), dtype=np.float16)
    r = min(OUT / H, OUT / W)
    nh, nw = int(round(H * r)), int(round(W * r))
    top, left = (OUT - nh) // 2, (OUT - nw) // 2
    for i, (im, op) in enumerate(zip(frames, ops)):
        if op == 1:
            im = cv2_port.noise(im, 15)
        elif op == 2:
            im = cv2_port.blur(im, 9, 0)
        elif op == 3:
            im = cv2_port.lowres(im, 0.5)
        im = cv2.resize(im, (nw, nh), interpolation=cv2.INTER_LINEAR)
        im = cv2.copyMakeBorder(im, top, OUT - nh - top, left, OUT - nw - left, cv2.BORDER_CONSTANT, value=(114, 114, 114))
        out[i] = (im[:, :, ::-1].transpose(2, 0, 1).astype(np.float32) / 255.0).astype(np.float16)
    return torch.from_numpy(out).cuda()


def from_files(n_batches):
    """The same path starting from JPEG files on tmpfs (what a dataloader starts from): FileCorruptionBatcher (device JPEG
    decoder) against cv2.imread + the reference's per-image work in one process."""
    import shutil
    import tempfile
    from robust_object_detection_b200.training import FileCorruptionBatcher
    rng = np.random.default_rng(6)
    d = tempfile.mkdtemp(prefix="rod_hook_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    try:
        paths = []
        for i in range(2 * B):
            p = os.path.join(d, f"f{i:03d}.jpg")
            cv2.imwrite(p, cv2.GaussianBlur(rng.integers(0, 256, (H, W, 3), dtype=np.uint8), (0, 0), 3.0))
            paths.append(p)
        batches = [[paths[(b * B + i) % len(paths)] for i in range(B)] for b in range(n_batches)]
        fb = FileCorruptionBatcher(out_hw=(OUT, OUT), seed=42)
        random.seed(1)
        for x in fb.run(batches[:3]):
            pass
        torch.cuda.synchronize()
        random.seed(42)
        t0 = time.perf_counter()
        for x in fb.run(batches):
            pass
        torch.cuda.synchronize()
        ours_s = time.perf_counter() - t0
        ref_batches = min(n_batches, 4)
        random.seed(42)
        t0 = time.perf_counter()
        for b in range(ref_batches):
            reference_batch([cv2.imread(p) for p in batches[b]], draw_decisions(B))
        torch.cuda.synchronize()
        ref_s = time.perf_counter() - t0
        ours, ref = n_batches * B / ours_s, ref_batches * B / ref_s
        return {"workload": "batch 16 of 1360x765 JPEG files (tmpfs) -> decode -> random corruption -> 640 letterbox -> fp16 NCHW",
                "ours_images_per_s": ours, "ours_ms_per_batch": 1e3 * ours_s / n_batches, "file_bytes_per_image": os.path.getsize(paths[0]),
                "reference_images_per_s_one_process": ref, "reference_ms_per_batch": 1e3 * ref_s / ref_batches,
                "speedup_vs_one_process": ours / ref, "speedup_vs_8_workers_upper_bound": ours / (8 * ref)}
    finally:
        shutil.rmtree(d, ignore_errors=True)


def main():
    n_batches = int(sys.argv[1]) if len(sys.argv) > 1 else 24
    rng = np.random.default_rng(5)
    pool = [rng.integers(0, 256, (H, W, 3), dtype=np.uint8) for _ in range(2 * B)]
    batches = [[pool[(b * B + i) % len(pool)] for i in range(B)] for b in range(n_batches)]
    batcher = CorruptionBatcher(out_hw=(OUT, OUT), seed=42)
    random.seed(1)
    for x in batcher.run(batches[:3]):  # warm-up: plans, pinned slots, first launches
        pass
    torch.cuda.synchronize()
    random.seed(42)
    t0 = time.perf_counter()
    for x in batcher.run(batches):
        pass
    torch.cuda.synchronize()
    ours_s = time.perf_counter() - t0
    random.seed(42)
    ref_batches = min(n_batches, 6)
    np.random.seed(42)
    reference_batch(batches[0], draw_decisions(B))
    random.seed(42)
    t0 = time.perf_counter()
    for b in range(ref_batches):
        reference_batch(batches[b], draw_decisions(B))
    torch.cuda.synchronize()
    ref_s = time.perf_counter() - t0
    ours = n_batches * B / ours_s
    ref = ref_batches * B / ref_s
    files = from_files(n_batches)
    print(json.dumps({
        "workload": "config 5: batch 16 of 1360x765 host frames -> random corruption -> 640 letterbox -> fp16 NCHW on the GPU",
        "ours_images_per_s": ours, "ours_ms_per_batch": 1e3 * ours_s / n_batches,
        "h2d_bytes_per_batch": B * H * W * 3, "pcie_GBps": ours * H * W * 3 / 1e9,
        "reference_images_per_s_one_process": ref, "reference_ms_per_batch": 1e3 * ref_s / ref_batches,
        "reference_images_per_s_x8_workers_upper_bound": 8 * ref,
        "host_cores": len(os.sched_getaffinity(0)), "opencv_threads": cv2.getNumThreads(),
        "speedup_vs_one_process": ours / ref, "speedup_vs_8_workers_upper_bound": ours / (8 * ref), "from_files": files}))


if __name__ == "__main__":
    main()
