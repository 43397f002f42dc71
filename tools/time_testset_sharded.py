"""BASELINE configs[3] with files on N GPUs: the drop-in test-set driver (robust-object-detection_b200/
build_corrupted_testsets.py) started once per GPU by torchrun, each rank owning a block of every image directory.

  python tools/time_testset_sharded.py [images_per_tree]                                   (one process, one GPU)
  python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/time_testset_sharded.py [n]

Rank 0 prints one JSON line: wall time of the slowest rank (NOISE_MODE = "philox", device JPEG encoder, files on tmpfs),
files per second, and a digest over every output file -- the digest of an N-rank run must equal the one-process run's
(Philox noise is keyed by the position in the directory listing).  gloo carries the two barriers; no pixel crosses ranks."""
import hashlib
import json
import os
import shutil
import sys
import time
from pathlib import Path

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from time_testset_driver import VARIANTS, make_tree  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
    decoder = sys.argv[2] if len(sys.argv) > 2 else "gpu"
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1:
        dist.init_process_group("gloo")
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    base = Path("/dev/shm" if os.path.isdir("/dev/shm") else "/tmp") / f"rod_sharded_{os.environ.get('MASTER_PORT', '0')}"

    def barrier():
        if world > 1:
            dist.barrier()

    if rank == 0:
        shutil.rmtree(base, ignore_errors=True)
        make_tree(base / "src", n)
    barrier()
    from robust_object_detection_b200 import build_corrupted_testsets as drv
    drv.YOLO_SRC, drv.COCO_SRC, drv.NOISE_MODE, drv.ENCODER = base / "src" / "yolo", base / "src" / "coco", "philox", "gpu"
    drv.DECODER = decoder
    keep = sys.stdout
    times = []
    parts = []   # (variant, seconds) of every _process_images call of the timed pass
    inner = drv._process_images

    def timed_process(src_dir, dst_dir, variant, run=None):
        t = time.perf_counter()
        inner(src_dir, dst_dir, variant, run)
        if run is not None:
            run.drain()
        parts.append((variant, round(time.perf_counter() - t, 4)))

    drv._process_images = timed_process
    try:
        for tag in ("warm", "timed"):
            drv.OUT_ROOT = base / tag
            parts.clear()
            barrier()
            t0 = time.perf_counter()
            sys.stdout = open(os.devnull, "w")
            try:
                drv.main()
            finally:
                sys.stdout.close()
                sys.stdout = keep
            mine = time.perf_counter() - t0
            barrier()
            times.append((mine, time.perf_counter() - t0))
        per_rank, per_rank_parts = [times[1][0]], [list(parts)]
        if world > 1:
            got = [None] * world
            dist.all_gather_object(got, (times[1][0], list(parts)))
            per_rank, per_rank_parts = [g[0] for g in got], [g[1] for g in got]
        if rank == 0:
            digest, files = hashlib.sha256(), 0
            for p in sorted((base / "timed").rglob("*.jpg")):
                digest.update(str(p.relative_to(base / "timed")).encode())
                digest.update(hashlib.sha256(p.read_bytes()).digest())
                files += 1
            assert files == 2 * len(VARIANTS) * n, files
            wall = times[1][1]
            print(json.dumps({"n_gpus": world, "images_per_tree": n, "output_files": files, "wall_s": wall,
                              "files_per_s": files / wall, "per_rank_s": per_rank, "per_rank_variant_s": per_rank_parts, "noise_mode": "philox", "encoder": "gpu", "decoder": decoder,
                              "host_cores": len(os.sched_getaffinity(0)), "files_sha256": digest.hexdigest()}))
    finally:
        barrier()
        if rank == 0:
            shutil.rmtree(base, ignore_errors=True)
        if world > 1:
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
