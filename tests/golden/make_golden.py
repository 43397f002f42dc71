"""Generate golden vectors from the UNMODIFIED reference (run in the build container only).

    python tests/golden/make_golden.py

Imports /root/reference/scripts/augmentations.py (numpy 2.3.5 + opencv 4.13.0, the
versions the reference pins) and records its outputs for seeded synthetic inputs:

  * golden_small.npz   -- full input/output arrays for small and odd shapes
  * golden_hashes.json -- sha256 of the outputs at BASELINE.json sizes, the
                          random-apply decision sequences, and library versions

Inputs are regenerated from their seeds by the tests, so only outputs are stored for
the large cases.  /root/reference does not exist on the GPU box; the tests read only
these two files.
"""
import hashlib
import json
import os
import random
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = "/root/reference"

SMALL_SHAPES = [(1, 1), (1, 7), (7, 1), (2, 2), (2, 3), (3, 5), (4, 2), (5, 4), (8, 8), (9, 13),
                (17, 9), (16, 48), (33, 47), (64, 64), (31, 100), (48, 129)]
BIG_CASES = [  # (name, seed, h, w)
    ("visdrone_765x1360", 12345, 765, 1360),
    ("cfg2_img0", 2000, 765, 1360),
    ("odd_1079x1917", 3001, 1079, 1917),
    ("even_1080x1920", 3002, 1080, 1920),
    ("mixed_1050x1400", 3003, 1050, 1400),
    ("max_1500x2000", 3004, 1500, 2000),
    ("max_1499x1999", 3005, 1499, 1999),
    ("mosaic_1024", 3006, 1024, 1024),
    ("patch_256", 3007, 256, 256),
    ("wide_9x1361", 3008, 9, 1361),
]


def synth(seed, h, w, kind="uniform"):
    rng = np.random.default_rng(seed)
    if kind == "uniform":
        return rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    if kind == "binary":
        return (rng.integers(0, 2, (h, w, 3)) * 255).astype(np.uint8)
    raise ValueError(kind)


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    sys.path.insert(0, REF_ROOT)
    import cv2
    from scripts import augmentations as ref

    small = {}
    for i, (h, w) in enumerate(SMALL_SHAPES):
        for kind in ("uniform", "binary"):
            img = synth(100 + i, h, w, kind)
            tag = f"{kind}_{h}x{w}"
            small[f"in_{tag}"] = img
            small[f"blur9_{tag}"] = ref.apply_motion_blur(img, 9, 0)
            small[f"blur5_{tag}"] = ref.apply_motion_blur(img, 5, 0)
            small[f"lowres_{tag}"] = ref.apply_lowres(img, 0.5)
            np.random.seed(7 + i)
            small[f"noise_{tag}"] = ref.apply_noise(img, 15)
    # strided (non-contiguous crop) input, as train_restoration.py:84 passes
    base = synth(999, 80, 120)
    crop = base[5:53, 7:91]
    small["in_crop_base"] = base
    small["blur9_crop"] = ref.apply_motion_blur(crop, 9, 0)
    small["lowres_crop"] = ref.apply_lowres(crop, 0.5)
    # other factors on a small image
    img = synth(555, 45, 70)
    small["in_factor"] = img
    for f in (0.25, 0.3, 0.75):
        small[f"lowres_f{f}"] = ref.apply_lowres(img, f)
    # truncation / clipping cases of SURVEY 8d config 1
    const = np.full((32, 40, 3), 128, np.uint8)
    np.random.seed(11)
    small["noise_const128"] = ref.apply_noise(const, 15)
    rails = synth(556, 32, 40, "binary")
    small["in_rails"] = rails
    np.random.seed(12)
    small["noise_rails"] = ref.apply_noise(rails, 15)
    # the motion-blur kernel itself
    small["kernel_9_0"] = ref._motion_blur_kernel(9, 0)
    np.savez_compressed(os.path.join(HERE, "golden_small.npz"), **small)

    hashes = {"versions": {"numpy": np.__version__, "cv2": cv2.__version__},
              "big": {}, "decisions": {}, "noise_sequence": {}}
    for name, seed, h, w in BIG_CASES:
        img = synth(seed, h, w)
        np.random.seed(42)
        hashes["big"][name] = {
            "seed": seed, "h": h, "w": w, "in": sha(img),
            "blur9": sha(ref.apply_motion_blur(img, 9, 0)),
            "lowres": sha(ref.apply_lowres(img, 0.5)),
            "noise_seed42": sha(ref.apply_noise(img, 15)),
        }
    # config 1: all 64 sequential apply_noise calls after one np.random.seed(42) (BASELINE.json configs[0])
    np.random.seed(42)
    seq = []
    for i in range(64):
        seq.append(sha(ref.apply_noise(synth(1000 + i, 765, 1360), 15)))
    hashes["noise_sequence"] = {"seed": 42, "first_image_seed": 1000, "sha": seq}
    # random-apply decisions (a6-a8): which op each of 64 consecutive calls picks
    names = {"noise": 1, "blur": 2, "lowres": 3}
    for gate in ("ultralytics", "pil"):
        random.seed(42)
        ops = []
        for _ in range(64):
            r = random.random()
            applied = (r < 0.5) if gate == "ultralytics" else not (r > 0.5)
            ops.append(names[random.choice(["noise", "blur", "lowres"])] if applied else 0)
        hashes["decisions"][gate] = ops
    # _apply_random_corruption end to end on a small image (python + numpy streams)
    random.seed(3)
    np.random.seed(3)
    img = synth(777, 40, 56)
    outs = [sha(ref._apply_random_corruption(img)) for _ in range(12)]
    hashes["random_corruption"] = {"py_seed": 3, "np_seed": 3, "img_seed": 777, "h": 40, "w": 56, "sha": outs}
    # RandomCorruption (PIL, RGB) end to end
    from PIL import Image
    random.seed(4)
    np.random.seed(4)
    pil = Image.fromarray(synth(778, 40, 56))
    t = ref.RandomCorruption(p=0.5)
    hashes["pil_transform"] = {"py_seed": 4, "np_seed": 4, "img_seed": 778, "h": 40, "w": 56,
                               "sha": [sha(np.array(t(pil))) for _ in range(12)]}
    with open(os.path.join(HERE, "golden_hashes.json"), "w") as f:
        json.dump(hashes, f, indent=1)
    print("wrote", len(small), "arrays and", len(hashes["big"]), "big-case hashes")


ANGLE_CASES = [(9, 45), (9, 135), (9, 30), (9, 90), (9, 179.5), (5, 60), (3, 45), (7, 33), (11, 45), (11, 100)]
ANGLE_SHAPES = [(1, 1), (2, 3), (5, 4), (9, 13), (33, 47), (31, 100), (48, 129), (8, 6), (20, 46), (3, 2)]
ANGLE_BIG = [("visdrone_765x1360", 12345, 765, 1360), ("tail3_540x961", 4001, 540, 961), ("tail2_301x402", 4002, 301, 402)]


def make_angles():
    """apply_motion_blur at angles other than 0 (general 2-D float kernel).  Stores the rotated
    kernels the reference builds (cv2.getRotationMatrix2D + warpAffine, augmentations.py:21-27) and its outputs."""
    sys.path.insert(0, REF_ROOT)
    from scripts import augmentations as ref
    arrs, hashes = {}, {}
    for k, ang in ANGLE_CASES:
        arrs[f"kernel_{k}_{ang}"] = ref._motion_blur_kernel(k, ang)
        for i, (h, w) in enumerate(ANGLE_SHAPES):
            for kind in ("uniform", "binary"):
                img = synth(300 + i, h, w, kind)
                arrs[f"out_{k}_{ang}_{kind}_{h}x{w}"] = ref.apply_motion_blur(img, k, ang)
        for name, seed, h, w in ANGLE_BIG:
            hashes[f"{k}_{ang}_{name}"] = sha(ref.apply_motion_blur(synth(seed, h, w), k, ang))
    np.savez_compressed(os.path.join(HERE, "golden_angles.npz"), **arrs)
    with open(os.path.join(HERE, "golden_angles.json"), "w") as f:
        json.dump({"cases": ANGLE_CASES, "shapes": ANGLE_SHAPES, "big": ANGLE_BIG, "sha": hashes}, f, indent=1)
    print("wrote", len(arrs), "angle arrays and", len(hashes), "hashes")


LETTERBOX_CASES = [  # (seed, h, w, out_h, out_w)
    (6000, 765, 1360, 640, 640), (6001, 1080, 1920, 640, 640), (6002, 360, 480, 320, 320), (6003, 640, 640, 640, 640),
    (6004, 500, 375, 640, 640), (6005, 1050, 1400, 1024, 1024), (6006, 97, 133, 64, 96), (6007, 720, 1280, 640, 640),
    (6008, 1078, 1916, 640, 640), (6009, 31, 45, 64, 64)]


def make_letterbox():
    """Config 5's formatting stage (Ultralytics LetterBox + Format + preprocess_batch) is not in /root/reference and
    Ultralytics is not installable here, so it cannot be pinned to Ultralytics itself.  These vectors pin it to LIVE
    OpenCV PRIMITIVES instead -- cv2.resize(INTER_LINEAR) to the documented LetterBox size, cv2.copyMakeBorder with
    value (114, 114, 114) and the documented rounding of the padding, BGR -> RGB, HWC -> CHW, / 255 -> float16 -- so the
    oracle's letterbox_norm_f16 (a NumPy restatement) and the GPU kernel are no longer only checked against each other.
    The GEOMETRY formula (r = min(oh/h, ow/w); new = round(w r), round(h r); top = round(dh - 0.1)) is recalled from
    Ultralytics 8.3.x and remains unverified against it."""
    import cv2
    out = {}
    for seed, h, w, oh, ow in LETTERBOX_CASES:
        img = synth(seed, h, w)
        r = min(oh / h, ow / w)
        new_w, new_h = int(round(w * r)), int(round(h * r))
        dw, dh = (ow - new_w) / 2, (oh - new_h) / 2
        res = img if (h, w) == (new_h, new_w) else cv2.resize(img, (new_w, new_h), interpolation=cv2.INTER_LINEAR)
        top, bottom = int(round(dh - 0.1)), int(round(dh + 0.1))
        left, right = int(round(dw - 0.1)), int(round(dw + 0.1))
        canvas = cv2.copyMakeBorder(res, top, bottom, left, right, cv2.BORDER_CONSTANT, value=(114, 114, 114))
        assert canvas.shape[:2] == (oh, ow), (canvas.shape, oh, ow)
        chw = np.ascontiguousarray(canvas[:, :, ::-1].transpose(2, 0, 1))
        f16 = (chw.astype(np.float32) / np.float32(255.0)).astype(np.float16)
        out[f"{seed}_{h}x{w}_{oh}x{ow}"] = sha(f16)
    with open(os.path.join(HERE, "golden_letterbox.json"), "w") as f:
        json.dump({"cases": LETTERBOX_CASES, "sha": out, "cv2": cv2.__version__}, f, indent=1)
    print("wrote", len(out), "letterbox hashes")


def make_restoration():
    """(corrupted, clean) pairs of the unmodified RestorationDataset.__getitem__
    (scripts/train_restoration.py:104-129) on synthetic JPEG-free inputs: cv2.imread is replaced by a stub that
    returns the synthetic frame, everything else (crop, flip, corruption choice, formatting) is the reference's."""
    sys.path.insert(0, REF_ROOT)
    import cv2
    from pathlib import Path
    from scripts import train_restoration as tr
    shapes = [(120, 200), (97, 133), (64, 64), (300, 180), (81, 90), (200, 201)]
    frames = {f"img{i}.jpg": synth(5000 + i, h, w) for i, (h, w) in enumerate(shapes)}
    real_imread = cv2.imread
    cv2.imread = lambda path, *a: frames[Path(path).name].copy()
    try:
        out = {}
        for is_train in (True, False):
            ds = tr.RestorationDataset(Path("/nonexistent"), patch_size=64, is_train=is_train)
            ds.img_paths = [Path(n) for n in frames]
            random.seed(2)
            np.random.seed(21)
            for i in range(len(frames)):
                cor, clean = ds[i]
                out[f"{'train' if is_train else 'val'}_cor_{i}"] = cor.numpy()
                out[f"{'train' if is_train else 'val'}_clean_{i}"] = clean.numpy()
    finally:
        cv2.imread = real_imread
    np.savez_compressed(os.path.join(HERE, "golden_restoration.npz"), **out)
    print("wrote", len(out), "restoration arrays")
    # frames smaller than the patch: the resize-first branch (train_restoration.py:79-81, 88-90)
    small = [(40, 80), (100, 50), (30, 30), (64, 20), (70, 64), (63, 65)]
    frames = {f"img{i}.jpg": synth(5100 + i, h, w) for i, (h, w) in enumerate(small)}
    cv2.imread = lambda path, *a: frames[Path(path).name].copy()
    try:
        out = {}
        for is_train in (True, False):
            ds = tr.RestorationDataset(Path("/nonexistent"), patch_size=64, is_train=is_train)
            ds.img_paths = [Path(n) for n in frames]
            random.seed(4)
            np.random.seed(41)
            for i in range(len(frames)):
                cor, clean = ds[i]
                out[f"{'train' if is_train else 'val'}_cor_{i}"] = cor.numpy()
                out[f"{'train' if is_train else 'val'}_clean_{i}"] = clean.numpy()
    finally:
        cv2.imread = real_imread
    np.savez_compressed(os.path.join(HERE, "golden_restoration_small.npz"), **out)
    print("wrote", len(out), "restoration arrays (frames smaller than the patch)")


JPEG_CASES = [(9300, 16, 16, "noise"), (9301, 37, 53, "smooth"), (9302, 97, 133, "noise"), (9303, 120, 200, "binary"),
              (9304, 360, 480, "smooth"), (9305, 765, 1360, "noise"), (9306, 540, 960, "smooth"), (9307, 17, 33, "noise")]


def jpeg_case_image(seed, h, w, kind):
    import cv2
    img = synth(seed, h, w)
    if kind == "smooth":
        img = cv2.GaussianBlur(img, (0, 0), 2.5)
    elif kind == "binary":
        img = (img > 127).astype(np.uint8) * 255
    return img


def make_jpeg():
    """sha256 of the files cv2.imencode('.jpg', img) -- the encoder behind the reference's
    cv2.imwrite(str(dst / name), out), build_corrupted_testsets.py:124 -- produces in the build container."""
    import cv2
    out = {f"{seed}_{h}x{w}_{kind}": hashlib.sha256(cv2.imencode(".jpg", jpeg_case_image(seed, h, w, kind))[1].tobytes()).hexdigest()
           for seed, h, w, kind in JPEG_CASES}
    with open(os.path.join(HERE, "golden_jpeg.json"), "w") as f:
        json.dump({"cases": JPEG_CASES, "sha": out, "cv2": cv2.__version__}, f, indent=1)
    print("wrote", len(out), "JPEG hashes")


JPEGDEC_CASES = [  # (seed, h, w, kind, imencode parameters as (name, value) pairs)
    (9400, 37, 53, "smooth", []), (9401, 64, 48, "noise", [("QUALITY", 60)]), (9402, 17, 33, "noise", [("OPTIMIZE", 1)]),
    (9403, 97, 133, "smooth", [("SAMPLING_FACTOR", "422")]), (9404, 40, 41, "noise", [("SAMPLING_FACTOR", "444"), ("QUALITY", 100)]),
    (9405, 50, 70, "grey", []), (9406, 120, 200, "smooth", [("RST_INTERVAL", 5)]), (9407, 9, 5, "noise", [("QUALITY", 20)]),
    (9408, 33, 64, "binary", [("RST_INTERVAL", 1), ("SAMPLING_FACTOR", "422")])]


def jpegdec_case_file(seed, h, w, kind, params):
    import cv2
    img = jpeg_case_image(seed, h, w, "noise" if kind == "grey" else kind)
    if kind == "grey":
        img = np.ascontiguousarray(img[:, :, 0])
    flat = []
    for name, val in params:
        flat.append(getattr(cv2, "IMWRITE_JPEG_" + name))
        flat.append(getattr(cv2, "IMWRITE_JPEG_SAMPLING_FACTOR_" + val) if name == "SAMPLING_FACTOR" else val)
    return cv2.imencode(".jpg", img, flat)[1].tobytes()


def make_jpegdec():
    """SURVEY 8f rank 1, reading side: small JPEG files written in the build container (OpenCV's encoder, several layouts)
    together with the sha256 of the pixels cv2.imdecode(file, IMREAD_COLOR) -- the decoder behind the reference's
    cv2.imread(str(img_path)), build_corrupted_testsets.py:109 -- returns there.  The files themselves are stored (base64):
    the check does not depend on the local OpenCV build at all."""
    import base64
    import cv2
    cases = []
    for seed, h, w, kind, params in JPEGDEC_CASES:
        data = jpegdec_case_file(seed, h, w, kind, params)
        pix = cv2.imdecode(np.frombuffer(data, np.uint8), cv2.IMREAD_COLOR)
        cases.append({"name": f"{seed}_{h}x{w}_{kind}", "params": params, "shape": list(pix.shape), "file_b64": base64.b64encode(data).decode(),
                      "pixels_sha256": hashlib.sha256(np.ascontiguousarray(pix).tobytes()).hexdigest()})
    with open(os.path.join(HERE, "golden_jpegdec.json"), "w") as f:
        json.dump({"cases": cases, "cv2": cv2.__version__}, f, indent=1)
    print("wrote", len(cases), "JPEG files with the hashes of their decoded pixels")


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "jpegdec":
        make_jpegdec()
    elif len(sys.argv) > 1 and sys.argv[1] == "jpeg":
        make_jpeg()
    elif len(sys.argv) > 1 and sys.argv[1] == "restoration":
        make_restoration()
    elif len(sys.argv) > 1 and sys.argv[1] == "angles":
        make_angles()
    elif len(sys.argv) > 1 and sys.argv[1] == "letterbox":
        make_letterbox()
    else:
        main()
        make_angles()
        make_restoration()
        make_letterbox()
        make_jpeg()
        make_jpegdec()
