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


def spawn_worker_corrupt(args):
    """Runs in a spawn()ed worker process (tests/test_gpu_parity.py::test_spawned_dataloader_workers): the drop-in
    functions initialise CUDA in the calling process, like a DataLoader worker started with the spawn method."""
    import os
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import random
    import numpy as np
    from robust_object_detection_b200 import augmentations as aug
    seed, h, w = args
    img = np.random.default_rng(seed).integers(0, 256, (h, w, 3), dtype=np.uint8)
    random.seed(seed)
    np.random.seed(seed)
    outs = [aug._apply_random_corruption(img) for _ in range(4)]
    return seed, [o.tobytes() for o in outs]


class CorruptDataset:
    """A torch-style map dataset whose __getitem__ runs the per-image drop-in hook, i.e. what the reference's DataLoader
    workers do (tests/test_gpu_parity.py::test_dataloader_workers_fork_and_spawn).  Module-level so that spawned workers can
    unpickle it."""

    def __init__(self, jobs):
        self.jobs = jobs

    def __len__(self):
        return len(self.jobs)

    def __getitem__(self, i):
        import os
        import sys
        sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
        import random
        import numpy as np
        from robust_object_detection_b200 import augmentations as aug
        seed, h, w = self.jobs[i]
        img = np.random.default_rng(seed).integers(0, 256, (h, w, 3), dtype=np.uint8)
        random.seed(seed)
        np.random.seed(seed)
        return seed, aug._apply_random_corruption(img).tobytes()
