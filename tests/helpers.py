"""Shared test helpers: seeded synthetic images (same generator as tests/golden/make_golden.py)."""
import hashlib

import numpy as np


def synth(seed, h, w, kind="uniform"):
    rng = np.random.default_rng(seed)
    if kind == "uniform":
        return rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    if kind == "binary":
        return (rng.integers(0, 2, (h, w, 3)) * 255).astype(np.uint8)
    raise ValueError(kind)


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


SMALL_SHAPES = [(1, 1), (1, 7), (7, 1), (2, 2), (2, 3), (3, 5), (4, 2), (5, 4), (8, 8), (9, 13),
                (17, 9), (16, 48), (33, 47), (64, 64), (31, 100), (48, 129)]
