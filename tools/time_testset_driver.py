"""Wall-clock comparison on real JPEG files (BASELINE config 4 with file I/O):

  reference loop   the serial loop of scripts/build_corrupted_testsets.py:85-173 restated with the same library calls
                   (oracle/cv2_port.py: cv2.imread -> corruption -> cv2.imwrite, np.random.seed(42) once)
  drop-in driver   robust-object-detection_b200/build_corrupted_testsets.py (GPU corruption, host codec on a thread
                   pool), NOISE_MODE = "compat" (byte-identical files) and "philox" (in-kernel RNG)

on a synthetic YOLO + COCO tree of N VisDrone-shaped JPEGs per tree.  Also checks that every file the compat-mode driver
wrote is byte-identical to the reference loop's.  Usage: python tools/time_testset_driver.py [N] > profiles/<tag>_testset_driver.json
(test / bench infrastructure: imports oracle/)."""
import json
import os
import shutil
import sys
import tempfile
import time
from pathlib import Path

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import cv2  # noqa: E402

from oracle import cv2_port  # noqa: E402

SHAPES = [(765, 1360), (1050, 1400), (788, 1400), (1078, 1916), (1080, 1920), (1500, 2000), (540, 960), (360, 480)]
VARIANTS = ["Test_Clean", "Test_Noise", "Test_Blur", "Test_LowRes"]


def make_tree(root: Path, n: int):
    rng = np.random.default_rng(4)
    for tree in ("yolo", "coco"):
        d = root / tree / "images" / "val"
        d.mkdir(parents=True)
        for i in range(n):
            h, w = SHAPES[int(rng.integers(0, len(SHAPES)))]
            img = cv2.GaussianBlur(rng.integers(0, 256, (h, w, 3), dtype=np.uint8), (0, 0), 3.0)
            cv2.imwrite(str(d / f"frame_{i:04d}.jpg"), img)
    (root / "yolo" / "labels" / "val").mkdir(parents=True)
    (root / "coco" / "annotations").mkdir(parents=True)
    (root / "coco" / "annotations" / "instances_val.json").write_text('{"images": [], "annotations": []}')


def reference_loop(root: Path, out: Path):
    """build_corrupted_testsets.py:169-173: set_seed(42); YOLO tree then COCO tree; variant-major; glob order; serial."""
    np.random.seed(42)
    ops = {"Test_Clean": lambda im: im, "Test_Noise": lambda im: cv2_port.noise(im, 15),
           "Test_Blur": lambda im: cv2_port.blur(im, 9, 0), "Test_LowRes": lambda im: cv2_port.lowres(im, 0.5)}
    n = 0
    for tree, sub in (("yolo", "yolo6"), ("coco", "coco6")):
        for v in VARIANTS:
            dst = out / sub / v / "images" / "val"
            dst.mkdir(parents=True)
            for p in (root / tree / "images" / "val").glob("*.*"):
                img = cv2.imread(str(p))
                if img is None:
                    continue
                cv2.imwrite(str(dst / p.name), ops[v](img))
                n += 1
    return n


def run_driver(root: Path, out: Path, mode: str, encoder: str = "gpu"):
    from robust_object_detection_b200 import build_corrupted_testsets as drv
    drv.YOLO_SRC, drv.COCO_SRC, drv.OUT_ROOT, drv.NOISE_MODE, drv.ENCODER = root / "yolo", root / "coco", out, mode, encoder
    keep, sys.stdout = sys.stdout, open(os.devnull, "w")  # the driver prints the reference's progress lines
    try:
        drv.main()
    finally:
        sys.stdout.close()
        sys.stdout = keep


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    base = Path(tempfile.mkdtemp(prefix="rod_testset_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None))
    try:
        make_tree(base / "src", n)
        files = 2 * len(VARIANTS) * n
        run_driver(base / "src", base / "warm", "philox")  # CUDA context, library load, first-launch costs
        shutil.rmtree(base / "warm")
        res = {"images_per_tree": n, "output_files": files, "shapes": "VisDrone frame sizes, smoothed synthetic content",
               "host_cores": len(os.sched_getaffinity(0)), "io": "tmpfs" if str(base).startswith("/dev/shm") else "disk"}
        t0 = time.perf_counter()
        assert reference_loop(base / "src", base / "ref") == files
        res["reference_loop_s"] = time.perf_counter() - t0
        ref_files = {p.relative_to(base / "ref"): p.read_bytes() for p in (base / "ref").rglob("*.jpg")}
        for encoder in ("gpu", "host"):
            for mode in ("compat", "philox"):
                tag = f"driver_{mode}" + ("" if encoder == "gpu" else "_hostcodec")
                out = base / tag
                t0 = time.perf_counter()
                run_driver(base / "src", out, mode, encoder)
                res[f"{tag}_s"] = time.perf_counter() - t0
                res[f"{tag}_files_per_s"] = files / res[f"{tag}_s"]
                # compat: every file must equal the reference loop's; philox: every file but Test_Noise's
                same = diff = 0
                for rel, want in ref_files.items():
                    if mode == "philox" and "Test_Noise" in rel.parts:
                        continue
                    if (out / rel).read_bytes() == want:
                        same += 1
                    else:
                        diff += 1
                res[f"{tag}_files_identical"], res[f"{tag}_files_different"] = same, diff
                shutil.rmtree(out)
        res["compat_files_identical"], res["compat_files_different"] = res["driver_compat_files_identical"], res["driver_compat_files_different"]
        res["encoder"] = "driver_* : JPEG encoded on the GPU (rod_jpeg_encode); driver_*_hostcodec : cv2.imwrite on the I/O threads"
        res["reference_loop_files_per_s"] = files / res["reference_loop_s"]
        res["speedup_compat"] = res["reference_loop_s"] / res["driver_compat_s"]
        res["speedup_philox"] = res["reference_loop_s"] / res["driver_philox_s"]
        print(json.dumps(res))
    finally:
        shutil.rmtree(base, ignore_errors=True)


if __name__ == "__main__":
    main()
