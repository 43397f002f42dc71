"""CPU: the oracle restatement vs golden vectors recorded from the unmodified reference
(tests/golden/make_golden.py).  Bit-exact (integer/byte work)."""
import os
import random

import numpy as np
import pytest

from oracle import corruption_oracle as orc
from tests.helpers import SMALL_SHAPES, sha, synth


@pytest.mark.parametrize("kind", ["uniform", "binary"])
def test_small_shapes_all_ops(golden_small, kind):
    for i, (h, w) in enumerate(SMALL_SHAPES):
        tag = f"{kind}_{h}x{w}"
        img = golden_small[f"in_{tag}"]
        assert np.array_equal(img, synth(100 + i, h, w, kind))
        assert np.array_equal(orc.apply_motion_blur(img, 9, 0), golden_small[f"blur9_{tag}"]), tag
        assert np.array_equal(orc.apply_motion_blur(img, 5, 0), golden_small[f"blur5_{tag}"]), tag
        assert np.array_equal(orc.apply_lowres(img, 0.5), golden_small[f"lowres_{tag}"]), tag
        np.random.seed(7 + i)
        assert np.array_equal(orc.apply_noise(img, 15), golden_small[f"noise_{tag}"]), tag


def test_strided_crop_and_factors(golden_small):
    crop = golden_small["in_crop_base"][5:53, 7:91]
    assert not crop.flags["C_CONTIGUOUS"]
    assert np.array_equal(orc.apply_motion_blur(crop, 9, 0), golden_small["blur9_crop"])
    assert np.array_equal(orc.apply_lowres(crop, 0.5), golden_small["lowres_crop"])
    img = golden_small["in_factor"]
    for f in (0.25, 0.3, 0.75):
        assert np.array_equal(orc.apply_lowres(img, f), golden_small[f"lowres_f{f}"])


def test_noise_truncation_and_rails(golden_small):
    np.random.seed(11)
    out = orc.apply_noise(np.full((32, 40, 3), 128, np.uint8), 15)
    assert np.array_equal(out, golden_small["noise_const128"])
    # truncation bias: mean(out - in) ~ -0.5 on mid-range input
    assert abs(float(out.astype(np.float64).mean()) - 127.5) < 0.6
    np.random.seed(12)
    assert np.array_equal(orc.apply_noise(golden_small["in_rails"], 15), golden_small["noise_rails"])


def test_blur_kernel_is_nine_equal_taps(golden_small):
    k = golden_small["kernel_9_0"]
    assert k.dtype == np.float32 and k.shape == (9, 9)
    assert np.count_nonzero(k) == 9 and np.all(k[4] == np.float32(1.0) / np.float32(9.0))
    assert orc.motion_blur_taps(9, 0) == 9
    with pytest.raises(NotImplementedError):
        orc.motion_blur_taps(9, 45)


def test_big_cases_sha(golden_hashes):
    for name, g in golden_hashes["big"].items():
        img = synth(g["seed"], g["h"], g["w"])
        assert sha(img) == g["in"], name
        assert sha(orc.apply_motion_blur(img, 9, 0)) == g["blur9"], name
        assert sha(orc.apply_lowres(img, 0.5)) == g["lowres"], name
        if g["h"] * g["w"] <= 1100000:
            np.random.seed(42)
            assert sha(orc.apply_noise(img, 15)) == g["noise_seed42"], name


def test_noise_sequence_config1(golden_hashes):
    g = golden_hashes["noise_sequence"]
    np.random.seed(g["seed"])
    for i, want in enumerate(g["sha"][:2]):
        assert sha(orc.apply_noise(synth(g["first_image_seed"] + i, 765, 1360), 15)) == want


def test_letterbox_oracle_vs_cv2_primitive_golden():
    """The formatting stage of config 5 (no reference code: it lives in Ultralytics) -- the oracle's NumPy restatement
    against vectors built in the build container from LIVE cv2.resize + cv2.copyMakeBorder(value=114) + / 255 -> float16
    (tests/golden/make_golden.py make_letterbox).  Pins the arithmetic to OpenCV primitives; the LetterBox geometry
    formula itself stays unverified against Ultralytics."""
    import json
    import os
    g = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "golden_letterbox.json")))
    for seed, h, w, oh, ow in g["cases"]:
        got = orc.letterbox_norm_f16(synth(seed, h, w), oh, ow)
        assert got.dtype == np.float16 and got.shape == (3, oh, ow)
        assert sha(got) == g["sha"][f"{seed}_{h}x{w}_{oh}x{ow}"], (h, w, oh, ow)


def test_decisions_and_random_corruption(golden_hashes):
    for gate in ("ultralytics", "pil"):
        random.seed(42)
        assert orc.draw_decisions(64, gate) == golden_hashes["decisions"][gate]
    g = golden_hashes["random_corruption"]
    random.seed(g["py_seed"])
    np.random.seed(g["np_seed"])
    img = synth(g["img_seed"], g["h"], g["w"])
    assert [sha(orc.apply_random_corruption(img)) for _ in g["sha"]] == g["sha"]


def test_general_angle_filter2d_vs_reference():
    """the oracle's filter2d() against the reference's apply_motion_blur at angles != 0
    (kernels and outputs recorded by tests/golden/make_golden.py from the unmodified reference)."""
    import json
    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    g = np.load(os.path.join(here, "golden_angles.npz"))
    meta = json.load(open(os.path.join(here, "golden_angles.json")))
    for k, ang in meta["cases"]:
        kern = g[f"kernel_{k}_{ang}"]
        for i, (h, w) in enumerate(meta["shapes"]):
            for kind in ("uniform", "binary"):
                img = synth(300 + i, h, w, kind)
                want = g[f"out_{k}_{ang}_{kind}_{h}x{w}"]
                assert np.array_equal(orc.filter2d(img, kern), want), (k, ang, h, w, kind)
        for name, seed, h, w in meta["big"]:
            if (k, ang) in ((9, 45), (11, 45), (9, 179.5)) or name.startswith("tail2"):
                assert sha(orc.filter2d(synth(seed, h, w), kern)) == meta["sha"][f"{k}_{ang}_{name}"], (k, ang, name)


RESTORATION_SHAPES = [(120, 200), (97, 133), (64, 64), (300, 180), (81, 90), (200, 201)]


def restoration_inputs(is_train):
    """Decisions and noise fields of six RestorationDataset.__getitem__ calls, consuming `random` / `np.random` in the
    reference's order (train_restoration.py:79-121, augmentations.py:31): returns (frames, decisions, fields)."""
    from robust_object_detection_b200.batch import draw_restoration_decisions
    frames = [synth(5000 + i, h, w) for i, (h, w) in enumerate(RESTORATION_SHAPES)]
    random.seed(2)
    np.random.seed(21)
    dec, fields = [], []
    for (h, w) in RESTORATION_SHAPES:
        d = draw_restoration_decisions(h, w, 64, is_train=is_train)
        dec.append(d)
        fields.append(orc.draw_noise_field((64, 64, 3), 15) if d[3] == orc.OP_NOISE else np.zeros((64, 64, 3), np.float32))
    return frames, dec, fields


@pytest.mark.parametrize("is_train", [True, False])
def test_restoration_pairs_vs_reference_dataset(is_train):
    """crop / flip / choice order + the three corruptions + RGB f32 CHW / 255, against pairs produced
    by the unmodified RestorationDataset (tests/golden/golden_restoration.npz)."""
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_restoration.npz"))
    frames, dec, fields = restoration_inputs(is_train)
    tag = "train" if is_train else "val"
    for i, (img, (y, x, flip, op)) in enumerate(zip(frames, dec)):
        patch = img[y:y + 64, x:x + 64]
        patch = np.ascontiguousarray(patch[:, ::-1] if flip else patch)
        cor = orc.add_noise_field(patch, fields[i]) if op == orc.OP_NOISE else orc.apply_op(patch, op)
        assert np.array_equal((patch[:, :, ::-1].astype(np.float32) / 255.0).transpose(2, 0, 1), g[f"{tag}_clean_{i}"]), i
        assert np.array_equal((cor[:, :, ::-1].astype(np.float32) / 255.0).transpose(2, 0, 1), g[f"{tag}_cor_{i}"]), (i, op)


RESTORATION_SMALL_SHAPES = [(40, 80), (100, 50), (30, 30), (64, 20), (70, 64), (63, 65)]


@pytest.mark.parametrize("is_train", [True, False])
def test_restoration_resize_first_branch_vs_reference_dataset(is_train):
    """Frames smaller than the patch (train_restoration.py:79-81, 88-90): the oracle's INTER_LINEAR enlargement, then
    the usual crop / flip / choice / corruption order, against pairs of the unmodified RestorationDataset
    (tests/golden/golden_restoration_small.npz)."""
    from robust_object_detection_b200.batch import draw_restoration_decisions
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_restoration_small.npz"))
    random.seed(4)
    np.random.seed(41)
    tag = "train" if is_train else "val"
    for i, (h, w) in enumerate(RESTORATION_SMALL_SHAPES):
        img = synth(5100 + i, h, w)
        if h < 64 or w < 64:
            img = orc.resize_linear(img, max(h, 64), max(w, 64))
        y, x, flip, op = draw_restoration_decisions(h, w, 64, is_train=is_train)
        patch = img[y:y + 64, x:x + 64]
        patch = np.ascontiguousarray(patch[:, ::-1] if flip else patch)
        cor = orc.apply_op(patch, op)   # noise draws from np.random's global stream, like the reference
        assert np.array_equal((patch[:, :, ::-1].astype(np.float32) / 255.0).transpose(2, 0, 1), g[f"{tag}_clean_{i}"]), i
        assert np.array_equal((cor[:, :, ::-1].astype(np.float32) / 255.0).transpose(2, 0, 1), g[f"{tag}_cor_{i}"]), (i, op)
