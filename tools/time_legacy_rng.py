"""Times rod_numpy_legacy_normal_f32 (csrc/np_legacy_rng.cpp, host code: the field of compat-mode noise) against
np.random.normal on this machine's cores and checks the two bit for bit.  No GPU, no torch.
Usage: python tools/time_legacy_rng.py [threads ...]"""
import ctypes
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = ctypes.CDLL(os.path.join(ROOT, "robust-object-detection_b200", "librod_b200.so"))
fn = lib.rod_numpy_legacy_normal_f32
fn.argtypes = [ctypes.c_void_p] * 4 + [ctypes.c_double, ctypes.c_uint64, ctypes.c_void_p, ctypes.c_int]
fn.restype = ctypes.c_int
n = 765 * 1360 * 3
out = np.empty(n, np.float32)


def run(threads):
    np.random.seed(42)
    st = np.random.get_state(legacy=True)
    key = np.array(st[1], dtype=np.uint32, copy=True)
    pos, has, cached = ctypes.c_int32(int(st[2])), ctypes.c_int32(0), ctypes.c_double(0.0)
    t0 = time.perf_counter()
    assert fn(key.ctypes.data, ctypes.byref(pos), ctypes.byref(has), ctypes.byref(cached), 15.0, n, out.ctypes.data, threads) == 0
    return time.perf_counter() - t0


np.random.seed(42)
t0 = time.perf_counter()
want = np.random.normal(0, 15, n).astype(np.float32)
res = {"cores": os.cpu_count(), "field": "1360x765x3 float32", "numpy_ms": round((time.perf_counter() - t0) * 1e3, 2)}
for t in [int(a) for a in sys.argv[1:]] or [1, 8, 0]:
    res[f"threads_{t or 'all'}_ms"] = round(min(run(t) for _ in range(7)) * 1e3, 2)
    res[f"threads_{t or 'all'}_equal"] = bool(np.array_equal(out, want))
print(json.dumps(res))
