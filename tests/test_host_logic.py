"""CPU: the C-ABI library loads and exports every declared symbol; host-side logic
(random-apply protocol, layouts, sharding incl. a world_size-2 gloo run).  No compute calls."""
import os
import re
import random
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "robust-object-detection_b200", "librod_b200.so")


@pytest.fixture(scope="module")
def built():
    if not os.path.exists(LIB):
        import __graft_entry__ as g
        g.build()
    return LIB


def test_abi_exports_every_declared_symbol(built):
    from robust_object_detection_b200 import _native
    header = open(os.path.join(ROOT, "include", "rod_b200.h")).read()
    declared = set(re.findall(r"ROD_API[^;(]*?\b(rod_\w+)\s*\(", header))
    assert len(declared) >= 16
    assert declared == set(_native.SYMBOLS), declared ^ set(_native.SYMBOLS)
    lib = _native.lib()
    for name in declared:
        assert getattr(lib, name) is not None
    exported = subprocess.check_output(["nm", "-D", "--defined-only", built], text=True)
    exported = set(re.findall(r" T (rod_\w+)", exported))
    assert declared <= exported
    assert b"sm_100a" in lib.rod_version()


def test_no_cpu_fallback_without_device(built):
    from robust_object_detection_b200 import _native
    if _native.lib().rod_device_count() > 0:
        pytest.skip("a GPU is visible")
    from robust_object_detection_b200 import augmentations as aug
    with pytest.raises(_native.RodError):
        aug.apply_motion_blur(np.zeros((4, 4, 3), np.uint8), 9, 0)


def test_argument_validation_precedes_device_work(built):
    from robust_object_detection_b200 import augmentations as aug
    with pytest.raises(NotImplementedError):
        aug.apply_motion_blur(np.zeros((4, 4, 3), np.uint8), 13, 45)  # DFT territory in OpenCV: rejected before any device work
    with pytest.raises(NotImplementedError):
        aug.apply_motion_blur(np.zeros((4, 4, 3), np.uint8), 8, 0)
    with pytest.raises(ValueError):
        aug.apply_lowres(np.zeros((4, 4), np.uint8), 0.5)
    with pytest.raises(ValueError):
        aug.apply_noise(np.zeros((4, 4, 3), np.float32), 15)


def test_constants_and_kernel(golden_small):
    from robust_object_detection_b200 import augmentations as aug
    assert (aug.NOISE_SIGMA, aug.BLUR_KERNEL, aug.BLUR_ANGLE_DEG, aug.DOWNSCALE_FACTOR) == (15, 9, 0, 0.5)
    k = aug._motion_blur_kernel(9, 0)
    assert k.dtype == np.float32 and np.array_equal(k, golden_small["kernel_9_0"])


def test_draw_decisions_matches_reference_streams(golden_hashes):
    from robust_object_detection_b200.batch import draw_decisions
    for gate in ("ultralytics", "pil"):
        random.seed(42)
        assert draw_decisions(64, gate).tolist() == golden_hashes["decisions"][gate]


def test_shard_ranges_cover_everything_once():
    from robust_object_detection_b200.sharding import shard_by_bytes, shard_range
    for n in (1, 7, 256, 1610):
        for world in (1, 2, 4, 8):
            seen = []
            for r in range(world):
                lo, hi = shard_range(n, r, world)
                seen += list(range(lo, hi))
            assert seen == list(range(n))
    sizes = [3 * h * w for h, w in [(765, 1360), (1080, 1920), (360, 480), (1500, 2000)] * 50]
    for world in (1, 2, 4, 8):
        blocks = shard_by_bytes(sizes, world)
        assert blocks[0][0] == 0 and blocks[-1][1] == len(sizes)
        assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
        loads = [sum(sizes[lo:hi]) for lo, hi in blocks]
        assert max(loads) <= 1.1 * sum(sizes) / world + max(sizes)


def test_testset_driver_shards(tmp_path, monkeypatch):
    """Which glob positions a rank of the sharded test-set build owns (BASELINE configs[3]): contiguous blocks that cover
    the directory once, torchrun's environment by default, the serial compat noise stream on rank 0 alone."""
    from robust_object_detection_b200 import build_corrupted_testsets as drv
    paths = []
    for i in range(23):
        q = tmp_path / f"f{i:02d}.jpg"
        q.write_bytes(b"x" * (1000 + 137 * (i % 5) * (i % 5)))
        paths.append(q)
    monkeypatch.setattr(drv, "NOISE_MODE", "philox")
    for world in (1, 2, 4, 8):
        for v in drv.VARIANTS:
            seen = []
            for r in range(world):
                monkeypatch.setenv("RANK", str(r))
                monkeypatch.setenv("WORLD_SIZE", str(world))
                lo, hi = drv._shard_of(paths, v)
                seen += list(range(lo, hi))
            assert seen == list(range(len(paths))), (world, v)
    monkeypatch.setattr(drv, "NOISE_MODE", "compat")
    monkeypatch.setenv("WORLD_SIZE", "4")
    monkeypatch.setenv("RANK", "0")
    assert drv._shard_of(paths, "Test_Noise") == (0, 23) and drv._shard_of(paths, "Test_Blur")[0] == 0
    monkeypatch.setenv("RANK", "3")
    assert drv._shard_of(paths, "Test_Noise") == (0, 0) and drv._shard_of(paths, "Test_Blur")[1] == 23
    monkeypatch.setattr(drv, "SHARD", (1, 2))   # explicit knob wins over the environment
    assert drv._rank_world() == (1, 2)
    monkeypatch.setattr(drv, "SHARD", None)
    assert drv._rank_world() == (0, 1) and drv._shard_of(paths, "Test_LowRes") == (0, 23)
    monkeypatch.setattr(drv, "SHARD", (2, 2))
    with pytest.raises(ValueError):
        drv._rank_world()


def test_testset_driver_write_slices(tmp_path):
    """File writes of one batch go to the I/O threads in slices (one future per thread, not per file): every file is
    written once, a failed write surfaces as IOError when the batch is drained, an empty batch is fine."""
    import threading
    from robust_object_detection_b200 import build_corrupted_testsets as drv
    run = drv._TreeRun()
    try:
        lock, seen = threading.Lock(), []

        def write(path, payload):
            with lock:
                seen.append((path, payload))
            with open(path, "wb") as f:
                f.write(payload)
            return True

        jobs = [(write, str(tmp_path / f"o{i:03d}.jpg"), bytes([i]) * (i + 1)) for i in range(37)]
        futs = run.submit_writes(jobs)
        assert 1 <= len(futs) <= run.pool._max_workers
        run.pending.append(futs)
        run.pending.append(run.submit_writes([]))
        run.drain(keep=1)
        assert len(run.pending) == 1
        run.drain()
        assert sorted(seen) == sorted((q, d) for _, q, d in jobs)
        for _, q, d in jobs:
            assert open(q, "rb").read() == d
        run.pending.append(run.submit_writes([(write, str(tmp_path / "a.jpg"), b"a"), (lambda q, d: False, "b", b"")]))
        with pytest.raises(IOError):
            run.drain()
    finally:
        run.pending.clear()
        run.close()


_GLOO_WORKER = r"""
import os, sys, json
sys.path.insert(0, sys.argv[1])
import torch.distributed as dist
from robust_object_detection_b200.sharding import shard_range, gather_records
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%s" % sys.argv[2], rank=int(sys.argv[3]), world_size=2)
lo, hi = shard_range(1610, dist.get_rank(), 2)
recs = gather_records({"rank": dist.get_rank(), "lo": lo, "hi": hi, "ms": 1.0 + dist.get_rank()})
if dist.get_rank() == 0:
    print(json.dumps(recs))
dist.destroy_process_group()
"""


def test_two_rank_gloo_sharding(tmp_path):
    import json
    import socket
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    script = tmp_path / "worker.py"
    script.write_text(_GLOO_WORKER)
    procs = [subprocess.Popen([sys.executable, str(script), ROOT, str(port), str(r)], stdout=subprocess.PIPE, text=True)
             for r in range(2)]
    outs = [p.communicate(timeout=120)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs)
    recs = json.loads(outs[0].strip().splitlines()[-1])
    assert [r["rank"] for r in recs] == [0, 1]
    assert recs[0]["lo"] == 0 and recs[0]["hi"] == recs[1]["lo"] and recs[1]["hi"] == 1610
    assert max(r["ms"] for r in recs) == 2.0


def test_forked_child_guard_and_spawn_switch(monkeypatch):
    """Process model of the reference's callers (train_yolo_augmented.py:16-19,33: the hook runs in workers=8 DataLoader
    processes).  (1) A process whose pid differs from the one that first ran a corruption here -- a fork()ed worker --
    gets a RuntimeError that names the fix, before any CUDA call.  (2) patch_ultralytics_augmentations() switches
    multiprocessing to 'spawn' (each worker then owns a CUDA context), unless ROD_KEEP_START_METHOD=1.  No GPU needed."""
    import multiprocessing
    import sys
    import types
    from robust_object_detection_b200 import augmentations as aug
    monkeypatch.setattr(aug, "_owner_pid", 1)           # "the parent"
    with pytest.raises(RuntimeError, match="spawn"):
        aug._check_process()
    with pytest.raises(RuntimeError, match="fork"):
        aug.apply_motion_blur(np.zeros((4, 4, 3), np.uint8), 9, 0)
    monkeypatch.setattr(aug, "_owner_pid", None)

    class Albumentations:
        def __call__(self, labels):
            return labels

    mod = types.ModuleType("ultralytics.data.augment")
    mod.Albumentations = Albumentations
    pkg, data = types.ModuleType("ultralytics"), types.ModuleType("ultralytics.data")
    pkg.data, data.augment = data, mod
    for name, m in (("ultralytics", pkg), ("ultralytics.data", data), ("ultralytics.data.augment", mod)):
        monkeypatch.setitem(sys.modules, name, m)
    before = multiprocessing.get_start_method(allow_none=True)
    try:
        monkeypatch.setenv("ROD_KEEP_START_METHOD", "1")
        multiprocessing.set_start_method("fork", force=True)
        aug.patch_ultralytics_augmentations()
        assert multiprocessing.get_start_method() == "fork"
        monkeypatch.setenv("ROD_KEEP_START_METHOD", "0")
        aug.patch_ultralytics_augmentations()
        assert multiprocessing.get_start_method() == "spawn"
    finally:
        multiprocessing.set_start_method(before, force=True)


def test_torch_library_ops_are_registered_and_traceable(built):
    """rod::noise / blur / lowres / corrupt_batch / corrupt_letterbox exist as torch.library custom ops, trace through
    their fake implementations (shapes / dtypes), and have no CPU kernel."""
    import torch
    from torch._subclasses.fake_tensor import FakeTensorMode
    import robust_object_detection_b200.torch_ops  # noqa: F401
    for name in ("noise", "blur", "lowres", "corrupt_batch", "corrupt_letterbox"):
        assert hasattr(torch.ops.rod, name), name
    with FakeTensorMode():
        x = torch.empty((2, 8, 12, 3), dtype=torch.uint8, device="cuda")
        ops = torch.empty(2, dtype=torch.uint8, device="cuda")
        for y in (torch.ops.rod.noise(x, 15.0, 1, 0), torch.ops.rod.blur(x, 9, 0.0), torch.ops.rod.lowres(x, 0.5),
                  torch.ops.rod.corrupt_batch(x, ops, 15.0, 9, 0.5, 1, 0)):
            assert y.shape == x.shape and y.dtype == torch.uint8 and y.device.type == "cuda"
        z = torch.ops.rod.corrupt_letterbox(x, ops, 64, 96, 114, 15.0, 9, 0.5, 1, 0)
        assert z.shape == (2, 3, 64, 96) and z.dtype == torch.float16
    with pytest.raises(NotImplementedError):
        torch.ops.rod.blur(torch.zeros((4, 4, 3), dtype=torch.uint8), 9, 0.0)


def test_numpy_legacy_normal_stream_bit_exact():
    """csrc/np_legacy_rng.cpp (host code, no GPU): the compat-mode field generator must be NumPy's global legacy stream
    bit for bit -- np.random.normal(0, sigma, shape).astype(float32) -- for any count (odd counts leave a cached
    Gaussian), from any generator state (fresh seed: pos = 624; mid-block; with a cached Gaussian pending), and must
    leave np.random in exactly the state the NumPy call would, so that mixed use continues on the same stream."""
    from robust_object_detection_b200 import augmentations as aug
    sizes = [1, 2, 3, 7, 155, 156, 157, 311, 312, 313, 1000, 4097, 8112, 12740, 12745, 16224, 65536, 100001, (37, 53, 3),
             (765, 1360, 3)]   # (a work item is 52 blocks of 624 words = 8112 candidate pairs, ~12 740 outputs)
    for seed in (0, 42, 123456789):
        np.random.seed(seed)
        want = [np.random.normal(0, 15, s).astype(np.float32) for s in sizes]
        want_tail = np.random.normal(0, 1, 9)
        want_uni = np.random.random(5)
        np.random.seed(seed)
        got = [aug.legacy_normal_f32(15, s) for s in sizes]
        assert all(g.dtype == np.float32 and g.shape == w.shape and np.array_equal(g, w) for g, w in zip(got, want)), seed
        assert np.array_equal(np.random.normal(0, 1, 9), want_tail) and np.array_equal(np.random.random(5), want_uni)
    # interleaved with other consumers of the global stream (uniforms move pos, standard_normal leaves a cached value)
    np.random.seed(7)
    a = [np.random.random(3), np.random.standard_normal(1), np.random.normal(0, 2.5, 11).astype(np.float32),
         np.random.randint(0, 100, 5), np.random.normal(0, 15, 622).astype(np.float32), np.random.standard_normal(2)]
    np.random.seed(7)
    b = [np.random.random(3), np.random.standard_normal(1), aug.legacy_normal_f32(2.5, 11),
         np.random.randint(0, 100, 5), aug.legacy_normal_f32(15, 622), np.random.standard_normal(2)]
    assert all(np.array_equal(x, y) for x, y in zip(a, b))
    # other sigmas, including 0 and a float32-inexact one
    for sigma in (0.0, 1.0, 0.1, 33.3):
        np.random.seed(5)
        w = np.random.normal(0, sigma, 5001).astype(np.float32)
        np.random.seed(5)
        assert np.array_equal(aug.legacy_normal_f32(sigma, 5001), w), sigma
    with pytest.raises(ValueError):
        aug.legacy_normal_f32(-1.0, 4)
    # caller-provided (page-locked when a GPU is there) output buffer, as apply_noise uses it
    np.random.seed(3)
    w = np.random.normal(0, 15, (37, 53, 3)).astype(np.float32)
    np.random.seed(3)
    buf = aug._pinned_field(37 * 53 * 3)
    got = aug.legacy_normal_f32(15, (37, 53, 3), out=buf)
    assert np.array_equal(got, w) and np.shares_memory(got, buf) and aug._pinned_field(37 * 53 * 3) is buf
    with pytest.raises(ValueError):
        aug.legacy_normal_f32(15, (4,), out=np.zeros(5, np.float32))
    # any thread count (1 = the calling thread alone produces and consumes) gives the same field and final state
    import ctypes
    from robust_object_detection_b200 import _native as N
    np.random.seed(11)
    n = 3 * 300 * 401 + 1
    want = np.random.normal(0, 15, n).astype(np.float32)
    want_state = np.random.get_state(legacy=True)
    for threads in (1, 2, 3, 7):
        np.random.seed(11)
        st = np.random.get_state(legacy=True)
        key = np.array(st[1], dtype=np.uint32)
        pos, has, cached = ctypes.c_int32(st[2]), ctypes.c_int32(st[3]), ctypes.c_double(st[4])
        out = np.empty(n, np.float32)
        assert N.lib().rod_numpy_legacy_normal_f32(key.ctypes.data, ctypes.byref(pos), ctypes.byref(has), ctypes.byref(cached),
                                                   15.0, n, out.ctypes.data, threads) == 0
        assert np.array_equal(out, want), threads
        assert np.array_equal(key, want_state[1]) and pos.value == want_state[2], threads
        assert has.value == want_state[3] and cached.value == want_state[4], threads


def test_numpy_legacy_normal_other_instruction_sets(tmp_path):
    """The generator picks AVX-512 / AVX2 / plain variants at run time; the two a machine with AVX-512 never takes are
    built here with the variant pinned (csrc/np_legacy_rng.cpp ROD_RNG_ISA_*) and checked against np.random too."""
    import ctypes
    import shutil
    if shutil.which("g++") is None:
        pytest.skip("no g++")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = os.path.join(root, "robust-object-detection_b200", "csrc", "np_legacy_rng.cpp")
    flags = open("/proc/cpuinfo").read() if os.path.exists("/proc/cpuinfo") else ""
    for isa in ("DEFAULT", "AVX2"):
        if isa == "AVX2" and " avx2" not in flags:
            continue
        so = str(tmp_path / f"rng_{isa.lower()}.so")
        subprocess.check_call(["g++", "-O3", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-fno-math-errno", "-pthread",
                               f"-DROD_RNG_ISA_{isa}", "-I", os.path.join(root, "include"), src, "-o", so])
        fn = ctypes.CDLL(so).rod_numpy_legacy_normal_f32
        fn.argtypes = [ctypes.c_void_p] * 4 + [ctypes.c_double, ctypes.c_uint64, ctypes.c_void_p, ctypes.c_int]
        for seed, pre, sizes, threads in ((1, 0, (3, 12741, 100001), 1), (2, 333, (8113, 1, 360901), 3), (3, 624, (40000, 7), 0)):
            np.random.seed(seed)
            np.random.random(pre)
            want = [np.random.normal(0, 15, n).astype(np.float32) for n in sizes]
            want_state = np.random.get_state(legacy=True)
            np.random.seed(seed)
            np.random.random(pre)
            for n, w in zip(sizes, want):
                st = np.random.get_state(legacy=True)
                key = np.array(st[1], dtype=np.uint32)
                pos, has, cached = ctypes.c_int32(st[2]), ctypes.c_int32(st[3]), ctypes.c_double(st[4])
                out = np.empty(n, np.float32)
                assert fn(key.ctypes.data, ctypes.byref(pos), ctypes.byref(has), ctypes.byref(cached), 15.0, n, out.ctypes.data, threads) == 0
                np.random.set_state(("MT19937", key, pos.value, has.value, cached.value))
                assert np.array_equal(out, w), (isa, seed, n)
            got_state = np.random.get_state(legacy=True)
            assert np.array_equal(got_state[1], want_state[1]) and got_state[2:] == want_state[2:], (isa, seed)


def test_numpy_legacy_normal_random_call_sequences():
    """Property test (hypothesis): any interleaving of np.random consumers and legacy_normal_f32 calls of random sizes,
    from any seed, produces the same values as the all-NumPy run and leaves the same generator state."""
    from hypothesis import given, settings, strategies as st
    from robust_object_detection_b200 import augmentations as aug

    step = st.one_of(st.tuples(st.just("normal"), st.integers(1, 5000), st.sampled_from([15.0, 1.0, 2.5])),
                     st.tuples(st.just("uniform"), st.integers(1, 700), st.just(0.0)),
                     st.tuples(st.just("std"), st.integers(1, 9), st.just(0.0)))

    @settings(max_examples=25, deadline=None)
    @given(seed=st.integers(0, 2 ** 32 - 1), steps=st.lists(step, min_size=1, max_size=8))
    def run(seed, steps):
        def play(ours):
            np.random.seed(seed)
            out = []
            for kind, n, sigma in steps:
                if kind == "normal":
                    out.append(aug.legacy_normal_f32(sigma, n) if ours else np.random.normal(0, sigma, n).astype(np.float32))
                elif kind == "uniform":
                    out.append(np.random.random(n))
                else:
                    out.append(np.random.standard_normal(n))
            return out, np.random.get_state(legacy=True)
        a, sa = play(False)
        b, sb = play(True)
        assert all(np.array_equal(x, y) for x, y in zip(a, b))
        assert np.array_equal(sa[1], sb[1]) and sa[2:] == sb[2:]

    run()


def test_jpeg_probe_through_the_c_abi(built):
    """rod_jpegdec_probe is host code (no device needed): the frame size for the layouts the device decoder takes, None
    for the ones it reports (the driver then reads the file with cv2.imread)."""
    import cv2
    from robust_object_detection_b200.jpeg import probe
    rng = np.random.default_rng(21)
    img = rng.integers(0, 256, (37, 53, 3), dtype=np.uint8)
    S = cv2.IMWRITE_JPEG_SAMPLING_FACTOR
    for params in ([], [cv2.IMWRITE_JPEG_QUALITY, 30], [cv2.IMWRITE_JPEG_OPTIMIZE, 1], [S, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_422],
                   [S, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_444], [cv2.IMWRITE_JPEG_RST_INTERVAL, 2]):
        assert probe(cv2.imencode(".jpg", img, params)[1].tobytes()) == (37, 53), params
    assert probe(cv2.imencode(".jpg", img[:, :, 0])[1].tobytes()) == (37, 53)
    for params in ([cv2.IMWRITE_JPEG_PROGRESSIVE, 1], [S, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_440],
                   [S, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_411]):
        assert probe(cv2.imencode(".jpg", img, params)[1].tobytes()) is None, params
    assert probe(cv2.imencode(".jpg", img[:, :4])[1].tobytes()) is None           # subsampled and narrower than 5 pixels
    assert probe(cv2.imencode(".png", img)[1].tobytes()) is None
    assert probe(b"") is None and probe(b"\xff\xd8\xff") is None
    # EXIF orientation other than 1 (cv2.imread rotates such files): spliced in as an APP1 segment behind SOI
    base = cv2.imencode(".jpg", img)[1].tobytes()
    def exif(orientation):
        tiff = b"II*\x00\x08\x00\x00\x00" + b"\x01\x00" + b"\x12\x01\x03\x00\x01\x00\x00\x00" + bytes([orientation, 0, 0, 0]) + b"\x00\x00\x00\x00"
        body = b"Exif\x00\x00" + tiff
        return base[:2] + b"\xff\xe1" + (len(body) + 2).to_bytes(2, "big") + body + base[2:]
    assert probe(exif(1)) == (37, 53)
    assert probe(exif(6)) is None
