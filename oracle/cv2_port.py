"""CPU baseline port -- TEST/BENCH INFRASTRUCTURE, NOT PRODUCT CODE.

The reference's hot path (scripts/augmentations.py:21-45) is a handful of calls into OpenCV
and NumPy.  This file issues the same library calls (same arguments, same order) so that
bench.py can time "what the reference runs on the host cores" on a box where /root/reference
does not exist, and so tests can cross-check the numpy restatement in corruption_oracle.py
against the live wheels.  Imported only by bench.py (cpu_baseline / --impl reference) and tests/.
"""
import numpy as np

try:
    import cv2
except Exception:  # pragma: no cover - the image ships opencv 4.13.0
    cv2 = None


def available() -> bool:
    return cv2 is not None


def line_kernel(k: int, angle_deg: float) -> np.ndarray:
    """k x k float32 line kernel rotated about its centre and normalised (augmentations.py:21-27)."""
    base = np.zeros((k, k), np.float32)
    base[k // 2, :] = 1.0
    centre = (k / 2 - 0.5, k / 2 - 0.5)
    rot = cv2.warpAffine(base, cv2.getRotationMatrix2D(centre, angle_deg, 1.0), (k, k))
    return rot / (rot.sum() + 1e-8)


def noise(img: np.ndarray, sigma: float) -> np.ndarray:
    field = np.random.normal(0, sigma, img.shape).astype(np.float32)
    return np.clip(img.astype(np.float32) + field, 0, 255).astype(np.uint8)


def blur(img: np.ndarray, k: int, angle_deg: float) -> np.ndarray:
    return cv2.filter2D(img, -1, line_kernel(k, angle_deg))


def lowres(img: np.ndarray, factor: float) -> np.ndarray:
    h, w = img.shape[:2]
    small_size = (max(1, int(w * factor)), max(1, int(h * factor)))
    small = cv2.resize(img, small_size, interpolation=cv2.INTER_AREA)
    return cv2.resize(small, (w, h), interpolation=cv2.INTER_LINEAR)


OPS = {"noise": lambda im: noise(im, 15), "blur": lambda im: blur(im, 9, 0), "lowres": lambda im: lowres(im, 0.5)}


def _pool_init():
    cv2.setNumThreads(1)


_POOL_IMAGES = {}  # per worker process: the synthetic input stays resident between tasks (like a decoded frame)


def _pool_task(args):
    op, seed, h, w, reps = args
    img = _POOL_IMAGES.get((seed, h, w))
    if img is None:
        img = _POOL_IMAGES[(seed, h, w)] = np.random.default_rng(seed).integers(0, 256, (h, w, 3), dtype=np.uint8)
    np.random.seed(seed)
    fn = OPS[op]
    import time
    t0 = time.perf_counter()
    for _ in range(reps):
        fn(img)
    return time.perf_counter() - t0


def time_op(op: str, h: int, w: int, n_images: int, mode: str = "threads", workers: int = 0):
    """images/s of `op` over n_images synthetic HxW images.
    mode 'threads': one process, OpenCV's own thread pool (what the reference gets by default);
    mode 'pool'   : `workers` processes with cv2.setNumThreads(1) each (best case for noise)."""
    import os
    import time
    if mode == "threads":
        imgs = [np.random.default_rng(s).integers(0, 256, (h, w, 3), dtype=np.uint8) for s in range(min(n_images, 8))]
        fn = OPS[op]
        fn(imgs[0])
        t0 = time.perf_counter()
        for i in range(n_images):
            fn(imgs[i % len(imgs)])
        dt = time.perf_counter() - t0
        return n_images / dt, cv2.getNumThreads()
    import multiprocessing as mp
    workers = workers or len(os.sched_getaffinity(0))
    reps = max(1, n_images // workers)
    ctx = mp.get_context("spawn")
    with ctx.Pool(workers, initializer=_pool_init) as pool:
        pool.map(_pool_task, [(op, s, h, w, 1) for s in range(workers)])  # warm-up / page-in
        t0 = time.perf_counter()
        pool.map(_pool_task, [(op, s, h, w, reps) for s in range(workers)])
        dt = time.perf_counter() - t0
    return workers * reps / dt, workers
