"""CPU oracle for the corruption pipeline -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module.  The product path (robust-object-detection_b200/) never
does; it fails loudly when the CUDA library is missing.

What this restates
------------------
The hot path of the reference is scripts/augmentations.py:30-56 (duplicated in
scripts/build_corrupted_testsets.py:41-59).  Those functions are three-line calls
into two third-party native libraries whose sources are NOT under /root/reference:

  * NumPy 2.3.5  (pinned in detr_env_requirements.txt)  -- apply_noise arithmetic
  * OpenCV 4.13.0 (opencv-python 4.13.0.92)             -- cv2.filter2D,
    cv2.resize(INTER_AREA), cv2.resize(INTER_LINEAR)

This module restates the *published algorithm* of those library routines in plain
numpy integer / float32 arithmetic (no cv2 import), one function per reference
call site.  Parity pinning: tests/golden/make_golden.py imports the unmodified
reference module from /root/reference in the build container and records its
outputs (full arrays for small shapes, sha256 for BASELINE-sized shapes) in
tests/golden/; tests/test_oracle_golden.py checks every function below against
those vectors bit-exactly.  tests/test_oracle_vs_cv2.py additionally compares it
with the live cv2/numpy wheels when they are importable.

Conventions: images are HWC uint8, 3 channels (BGR for the reference callers, but
every operation here is per-channel), possibly non-contiguous; results are fresh
C-contiguous uint8 arrays.
"""
from __future__ import annotations

import math
import random as _pyrandom

import numpy as np

# Module constants of the reference (scripts/augmentations.py:14-17).
NOISE_SIGMA = 15
BLUR_KERNEL = 9
BLUR_ANGLE_DEG = 0
DOWNSCALE_FACTOR = 0.5

OP_NONE, OP_NOISE, OP_BLUR, OP_LOWRES = 0, 1, 2, 3
_CHOICES = ("noise", "blur", "lowres")  # scripts/augmentations.py:50


# ----------------------------------------------------------------------------
# a1  apply_noise  (scripts/augmentations.py:30-33)
# ----------------------------------------------------------------------------
def add_noise_field(img: np.ndarray, noise_f32: np.ndarray) -> np.ndarray:
    """uint8 + supplied float32 field -> uint8, exactly as augmentations.py:32-33:
    one float32 round-to-nearest add, clamp to [0,255], then C-style truncation."""
    assert img.dtype == np.uint8 and noise_f32.dtype == np.float32
    s = img.astype(np.float32) + noise_f32
    s = np.minimum(np.maximum(s, np.float32(0)), np.float32(255))
    return np.ascontiguousarray(s.astype(np.uint8))  # astype truncates toward zero


def draw_noise_field(shape, sigma: float) -> np.ndarray:
    """The field augmentations.py:31 draws: NumPy *global legacy* RNG, float64,
    C order over (H, W, 3), then cast to float32."""
    return np.random.normal(0, sigma, shape).astype(np.float32)


def apply_noise(img: np.ndarray, sigma: float) -> np.ndarray:
    return add_noise_field(img, draw_noise_field(img.shape, sigma))


# ----------------------------------------------------------------------------
# Philox4x32-10 + quantile-table pairs / Box-Muller (the GPU "philox" noise mode; no reference twin --
# this restates OUR kernels' documented streams so tests can check them exactly).
# ----------------------------------------------------------------------------
_PHILOX_M0 = np.uint64(0xD2511F53)
_PHILOX_M1 = np.uint64(0xCD9E8D57)
_PHILOX_W0 = 0x9E3779B9
_PHILOX_W1 = 0xBB67AE85


def philox4x32_10(ctr: np.ndarray, key: np.ndarray, rounds: int = 10) -> np.ndarray:
    """ctr: uint32[N,4], key: uint32[N,2] (or [2]) -> uint32[N,4].  Salmon et al. 2011 (Philox4x32-`rounds`)."""
    c = ctr.astype(np.uint64).copy()
    k0 = np.broadcast_to(key[..., 0], (c.shape[0],)).astype(np.uint64).copy()
    k1 = np.broadcast_to(key[..., 1], (c.shape[0],)).astype(np.uint64).copy()
    mask = np.uint64(0xFFFFFFFF)
    for _ in range(rounds):
        p0 = _PHILOX_M0 * c[:, 0]
        p1 = _PHILOX_M1 * c[:, 2]
        hi0, lo0 = p0 >> np.uint64(32), p0 & mask
        hi1, lo1 = p1 >> np.uint64(32), p1 & mask
        n0 = hi1 ^ c[:, 1] ^ k0
        n1 = lo1
        n2 = hi0 ^ c[:, 3] ^ k1
        n3 = lo0
        c[:, 0], c[:, 1], c[:, 2], c[:, 3] = n0, n1, n2, n3
        k0 = (k0 + np.uint64(_PHILOX_W0)) & mask
        k1 = (k1 + np.uint64(_PHILOX_W1)) & mask
    return c.astype(np.uint32)


_PHILOX_TAIL_FLIP = 0x80000000
_NOISE_K_PER_SIGMA = np.float32(1.17741002251547466)       # sqrt(2 ln 2), as the fp32 constant the kernel uses
_ANGLE_SCALE = float(np.float32(9.58737992428525768e-05))  # float32(2 pi / 65536)
_ANGLE_BIAS = float(np.float32(-804.2476806640625))        # float32((0.5 - 2^23) * 2 pi / 65536)


GAUSS_TABLE_MAX_SIGMA = 20.0   # rod_core.h ROD_GAUSS_TABLE_MAX_SIGMA
GAUSS_TABLE_MIN_SIGMA = 3.0    # rod_core.h ROD_GAUSS_TABLE_MIN_SIGMA
GAUSS_H4_STRETCH_A = -1.4657745851475e-3   # rod_core.h ROD_GAUSS_H4_STRETCH_A / _B
GAUSS_H4_STRETCH_B = 2.4950155916569e-5


def gauss_table(sigma: float) -> np.ndarray:
    """int64[256]: X[i] = round(128 * float32(sigma) * y_i), the 256-point discretisation of N(0, (sigma/2)^2) in 1/256
    pixel units of the table generator (rod_tables.h build_gauss_table): z_i = mean of N(0,1) over the i-th of 256
    equiprobable cells, y_i = z_i (1 + A z_i^4 + B z_i^8) normalised to unit variance (A, B make the 4th and 6th
    moments 3 and 15).  scipy's ndtri is the independent inverse normal CDF here."""
    from scipy.special import ndtri
    q = ndtri(np.arange(128, 257, dtype=np.float64) / 256.0)      # cell edges of the upper half; q[0] = 0, q[128] = inf
    q[0] = 0.0
    ph = np.exp(-0.5 * q[:-1] ** 2) * 0.39894228040143267794
    ph = np.concatenate([ph, [0.0]])
    z = 256.0 * (ph[:-1] - ph[1:])
    z4 = (z * z) * (z * z)
    y = z * (1.0 + GAUSS_H4_STRETCH_A * z4 + GAUSS_H4_STRETCH_B * (z4 * z4))
    norm = 128.0 * float(np.float32(sigma)) / np.sqrt(np.sum(y * y) / 128.0)
    up = np.floor(norm * y + 0.5).astype(np.int64)
    return np.concatenate([-up[::-1], up])


def philox_noise_field_table(n_elems: int, sigma: float, seed: int, image_index: int,
                             offset: int = 0, rounds: int = 10) -> np.ndarray:
    """Restatement of the TABLE generator of Philox mode (rod_core.h / noise.cu noise_table_kernel), the default
    for 3 <= sigma <= 20.  Integer arithmetic only; one Philox block per group of 16 elements (g = e >> 4, counter
    (g, image lo, image hi, offset), key = seed); word r_q of the block gives four 8-bit draws
    xa = X[r_q & 255], xb = X[(r_q >> 8) & 255], xc = X[(r_q >> 16) & 255], xd = X[r_q >> 24] and their Hadamard mix
      element 16g + 4q + 0 : k = floor((xa + xb + xc + xd) / 256)
      element 16g + 4q + 1 : k = floor((xa - xb + xc - xd) / 256)
      element 16g + 4q + 2 : k = floor((xa + xb - xc - xd) / 256)
      element 16g + 4q + 3 : k = floor((xa - xb - xc + xd) / 256)
    Returns k as float64 (an integer: add_philox_noise's floor is then the identity).  rounds = 7 restates
    ROD_GAUSS_TABLE_PHILOX7."""
    n_groups = (n_elems + 15) // 16
    g = np.arange(n_groups, dtype=np.uint64)
    ctr = np.empty((n_groups, 4), dtype=np.uint32)
    ctr[:, 0] = (g & np.uint64(0xFFFFFFFF)).astype(np.uint32)
    ctr[:, 1] = np.uint32(image_index & 0xFFFFFFFF)
    ctr[:, 2] = np.uint32((image_index >> 32) & 0xFFFFFFFF)
    ctr[:, 3] = np.uint32(offset & 0xFFFFFFFF)
    key = np.array([seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF], dtype=np.uint32)
    r = philox4x32_10(ctr, key, rounds).astype(np.int64)          # [n_groups, 4]
    X = gauss_table(sigma)
    xa, xb, xc, xd = X[r & 0xFF], X[(r >> 8) & 0xFF], X[(r >> 16) & 0xFF], X[r >> 24]
    k = np.empty((n_groups, 4, 4), dtype=np.float64)
    k[:, :, 0] = (xa + xb + xc + xd) >> 8
    k[:, :, 1] = (xa - xb + xc - xd) >> 8
    k[:, :, 2] = (xa + xb - xc - xd) >> 8
    k[:, :, 3] = (xa - xb - xc + xd) >> 8
    return k.reshape(-1)[:n_elems]


def philox_noise_field(n_elems: int, sigma: float, seed: int, image_index: int,
                       offset: int = 0, generator: str = "auto") -> np.ndarray:
    """The field Philox mode adds: generator "auto" (what rod_noise_u8 / rod_corrupt_batch_u8 use: the table
    generator for 3 <= sigma <= 20, else Box-Muller), "table", "table7" (ROD_GAUSS_TABLE_PHILOX7: like auto with the
    table generator on Philox4x32-7) or "boxmuller" (always used by the training path, rod_corrupt_letterbox_f16)."""
    in_range = GAUSS_TABLE_MIN_SIGMA <= float(np.float32(sigma)) <= GAUSS_TABLE_MAX_SIGMA
    if generator == "table" or (generator in ("auto", "table7") and in_range):
        return philox_noise_field_table(n_elems, sigma, seed, image_index, offset, 7 if generator == "table7" else 10)
    return philox_noise_field_boxmuller(n_elems, sigma, seed, image_index, offset)


def philox_noise_field_boxmuller(n_elems: int, sigma: float, seed: int, image_index: int,
                                 offset: int = 0) -> np.ndarray:
    """float64 restatement of the Box-Muller generator of Philox mode for one image
    (robust-object-detection_b200/csrc/rod_core.h: philox4x32_10 + gauss8).

    Element e (flat HWC index) belongs to group g = e // 8; word p = (e % 8) // 2 of the
    group's Philox block makes the Box-Muller pair (8g + 2p, 8g + 2p + 1):
      ctr = [g, image_index & 0xffffffff, image_index >> 32, offset], key = [seed lo, seed hi]
      r = Philox4x32-10(ctr, key)
      u  = ((r_p >> 16) + 0.5) * 2^-16                         radius uniform (stratified, 16 bits)
           if r_p >> 16 == 0: u = (t_p + 0.5) * 2^-48 with t = Philox block at ctr[2] ^ 0x80000000
      th = (2^23 + (r_p & 0xffff)) * float32(2 pi / 65536) + float32((0.5 - 2^23) 2 pi / 65536)
      s0 = sqrt(-log2 u) cos th, s1 = sqrt(-log2 u) sin th
      noise[e] = K * s with K = float32(sigma) * float32(sqrt(2 ln 2)).
    The kernel's output is add_philox_noise(img, noise): clamp(v + floor(noise), 0, 255).
    """
    n_groups = (n_elems + 7) // 8
    g = np.arange(n_groups, dtype=np.uint64)
    ctr = np.empty((n_groups, 4), dtype=np.uint32)
    ctr[:, 0] = (g & np.uint64(0xFFFFFFFF)).astype(np.uint32)
    ctr[:, 1] = np.uint32(image_index & 0xFFFFFFFF)
    ctr[:, 2] = np.uint32((image_index >> 32) & 0xFFFFFFFF)
    ctr[:, 3] = np.uint32(offset & 0xFFFFFFFF)
    key = np.array([seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF], dtype=np.uint32)
    r = philox4x32_10(ctr, key)
    ctr_t = ctr.copy()
    ctr_t[:, 2] ^= np.uint32(_PHILOX_TAIL_FLIP)
    t = philox4x32_10(ctr_t, key)
    K = float(np.float32(sigma) * _NOISE_K_PER_SIGMA)
    s = np.empty((n_groups, 8), dtype=np.float64)
    for p in range(4):
        hi = (r[:, p] >> np.uint32(16)).astype(np.float64)
        lo = (r[:, p] & np.uint32(0xFFFF)).astype(np.float64)
        u = (hi + 0.5) * 2.0 ** -16
        u = np.where(hi == 0, (t[:, p].astype(np.float64) + 0.5) * 2.0 ** -48, u)
        rad = np.sqrt(-np.log2(u))
        th = (8388608.0 + lo) * _ANGLE_SCALE + _ANGLE_BIAS
        s[:, 2 * p] = rad * np.cos(th)
        s[:, 2 * p + 1] = rad * np.sin(th)
    return (K * s).reshape(-1)[:n_elems]


def add_philox_noise(img: np.ndarray, noise: np.ndarray) -> np.ndarray:
    """Philox-mode output: clamp(v + floor(noise), 0, 255).  Equals the reference's
    trunc(clip(v + noise, 0, 255)) in real arithmetic (v is an integer); the reference's float32
    rounding of v + noise moves about 4e-6 of the elements across an integer, which Philox mode
    (validated statistically, not bit-exactly) does not reproduce."""
    x = img.astype(np.int64) + np.floor(np.asarray(noise, dtype=np.float64)).astype(np.int64).reshape(img.shape)
    return np.clip(x, 0, 255).astype(np.uint8)


# ----------------------------------------------------------------------------
# a2/a3  _motion_blur_kernel + apply_motion_blur  (scripts/augmentations.py:21-38)
# ----------------------------------------------------------------------------
def reflect101(i: np.ndarray, n: int) -> np.ndarray:
    """cv::borderInterpolate(BORDER_REFLECT_101), periodic form (valid for any reach)."""
    if n == 1:
        return np.zeros_like(i)
    period = 2 * (n - 1)
    m = np.mod(i, period)
    return np.where(m < n, m, period - m)


def motion_blur_taps(k: int, angle_deg: float) -> int:
    """At angle 0 the warpAffine in augmentations.py:24-25 is the identity, so the
    kernel is k taps of float32(1/k) on row k//2 (anchor = k//2).  Returns k."""
    if float(angle_deg) != 0.0:
        raise NotImplementedError("oracle covers angle_deg == 0 only (BASELINE scope)")
    if k < 1:
        raise ValueError("k must be >= 1")
    return int(k)


DFT_FILTER_ELEMS = 130  # cv::filter2D switches 8U kernels with >= 130 elements to DFT-based convolution


def filter2d(img: np.ndarray, kernel: np.ndarray) -> np.ndarray:
    """cv2.filter2D(img, -1, kernel) for a small float32 kernel (fewer than 130 elements: the direct,
    non-DFT FilterEngine of OpenCV 4.13.0, modules/imgproc/src/filter.simd.hpp Filter2D + FilterVec_8u),
    as called by augmentations.py:38 with the rotated line kernel of augmentations.py:21-27.

    Per output byte (channels interleaved, BORDER_REFLECT_101 on both axes): the non-zero taps are
    accumulated in float32 in row-major kernel order starting from 0.  The vectorised part of each row
    (flat byte index < 4 * (3W // 4)) uses fused multiply-add (v_muladd on an FMA3 host), the scalar tail
    (the last 3W % 4 bytes of the row) a separate multiply and add; the sum is rounded half-to-even
    (cvRound) and saturated.  Pinned by tests/golden/golden_angles.npz (outputs of the reference)."""
    assert img.dtype == np.uint8 and img.ndim == 3
    kernel = np.asarray(kernel, dtype=np.float32)
    k = kernel.shape[0]
    assert kernel.shape == (k, k) and k % 2 == 1 and k * k < DFT_FILTER_ELEMS
    h, w, c = img.shape
    pad = k // 2
    ys = reflect101(np.arange(-pad, h + pad), h)
    xs = reflect101(np.arange(-pad, w + pad), w)
    src = img[ys][:, xs].astype(np.float32).reshape(h + 2 * pad, (w + 2 * pad) * c)
    n = w * c
    fma = np.zeros((h, n), np.float32)
    sep = np.zeros((h, n), np.float32)
    for dy in range(k):
        for dx in range(k):
            wgt = kernel[dy, dx]
            if wgt == 0:
                continue
            v = src[dy:dy + h, c * dx:c * dx + n]
            # float32 fma emulated in float64: the product is exact, the sum is rounded to float32 once
            fma = (fma.astype(np.float64) + v.astype(np.float64) * np.float64(wgt)).astype(np.float32)
            sep = sep + v * wgt
    n_vec = n & ~3
    s = np.concatenate([fma[:, :n_vec], sep[:, n_vec:]], axis=1)
    return np.ascontiguousarray(np.clip(np.rint(s), 0, 255).astype(np.uint8).reshape(h, w, c))


def apply_motion_blur(img: np.ndarray, k: int, angle_deg: float, kernel: np.ndarray = None) -> np.ndarray:
    """cv2.filter2D(img, -1, kernel) with the kernel of augmentations.py:21-27.
    angle 0: horizontal k-tap box, BORDER_REFLECT_101, out = (2S + k) // (2k) (integer; equals OpenCV's float
    path because S/k is never within 1/(2k) of a tie).  Other angles: `kernel` must be the rotated k x k
    float32 kernel (built by the caller with cv2, or taken from the golden file) -> filter2d()."""
    if float(angle_deg) != 0.0 or kernel is not None:
        if kernel is None:
            raise ValueError("angle != 0 needs the rotated kernel (cv2.getRotationMatrix2D + warpAffine)")
        return filter2d(img, kernel)
    k = motion_blur_taps(k, angle_deg)
    assert img.dtype == np.uint8 and img.ndim == 3
    h, w, c = img.shape
    anchor = k // 2
    xs = np.arange(w)
    acc = np.zeros((h, w, c), dtype=np.int32)
    for j in range(k):
        acc += img[:, reflect101(xs + j - anchor, w), :]
    return np.ascontiguousarray(((2 * acc + k) // (2 * k)).astype(np.uint8))


# ----------------------------------------------------------------------------
# a4  cv2.resize(..., INTER_AREA)  (scripts/augmentations.py:44)
# ----------------------------------------------------------------------------
def lowres_small_size(h: int, w: int, factor: float):
    """augmentations.py:43 -- note int() truncation and the max(1, .) floor."""
    return max(1, int(h * factor)), max(1, int(w * factor))


def area_table(ssize: int, dsize: int):
    """OpenCV computeResizeAreaTab (modules/imgproc/src/resize.cpp, 4.13.0):
    list of (dst index, src index, float32 alpha) in emission order."""
    scale = ssize / dsize  # double
    tab = []
    for d in range(dsize):
        f1 = d * scale
        f2 = f1 + scale
        cell = min(scale, ssize - f1)
        s1 = math.ceil(f1)
        s2 = min(math.floor(f2), ssize - 1)
        s1 = min(s1, s2)
        if s1 - f1 > 1e-3:
            tab.append((d, s1 - 1, np.float32((s1 - f1) / cell)))
        for s in range(s1, s2):
            tab.append((d, s, np.float32(1.0 / cell)))
        if f2 - s2 > 1e-3:
            tab.append((d, s2, np.float32(min(min(f2 - s2, 1.0), cell) / cell)))
    return tab


def area_table_padded(ssize: int, dsize: int):
    """Same table as rectangular arrays: first[d], count[d], alpha[d, t] (float32)."""
    tab = area_table(ssize, dsize)
    first = np.zeros(dsize, dtype=np.int32)
    count = np.zeros(dsize, dtype=np.int32)
    rows = [[] for _ in range(dsize)]
    for d, s, a in tab:
        if not rows[d]:
            first[d] = s
        else:
            assert s == first[d] + len(rows[d])  # taps are consecutive source indices
        rows[d].append(a)
    for d in range(dsize):
        count[d] = len(rows[d])
    maxt = max(int(count.max()), 1)
    alpha = np.zeros((dsize, maxt), dtype=np.float32)
    for d in range(dsize):
        alpha[d, : count[d]] = rows[d]
    return first, count, alpha


def is_area_fast2(h: int, w: int, nh: int, nw: int) -> bool:
    """cv::resize picks resizeAreaFast_ (the integer 2x2 mean) only when BOTH axis
    scales are the exact integer 2."""
    return w == 2 * nw and h == 2 * nh


def resize_area(img: np.ndarray, nh: int, nw: int) -> np.ndarray:
    assert img.dtype == np.uint8 and img.ndim == 3
    h, w, c = img.shape
    if nh > h or nw > w:
        raise NotImplementedError("INTER_AREA upscaling is outside the path")
    if nh == h and nw == w:
        return np.ascontiguousarray(img)
    if is_area_fast2(h, w, nh, nw):
        a = img.astype(np.int32)
        s = a[0::2, 0::2] + a[0::2, 1::2] + a[1::2, 0::2] + a[1::2, 1::2]
        return np.ascontiguousarray(((s + 2) >> 2).astype(np.uint8))
    sx = w / nw
    sy = h / nh
    if abs(sx - round(sx)) < 2.3e-16 and abs(sy - round(sy)) < 2.3e-16:
        # other exact integer scales go through resizeAreaFast_ too: int sum * float(1/area)
        ix, iy = int(round(sx)), int(round(sy))
        a = img.astype(np.int32).reshape(nh, iy, nw, ix, c)
        s = a.sum(axis=(1, 3)).astype(np.float32)
        r = s * np.float32(1.0 / (ix * iy))
        return np.ascontiguousarray(np.clip(np.rint(r), 0, 255).astype(np.uint8))
    xf, xc, xa = area_table_padded(w, nw)
    yf, yc, ya = area_table_padded(h, nh)
    src = img.astype(np.float32)
    # horizontal pass: buf = 0; buf = buf + S*alpha in tap order (separate mul and add)
    buf = np.zeros((h, nw, c), dtype=np.float32)
    for t in range(xa.shape[1]):
        sel = np.nonzero(xc > t)[0]
        if sel.size == 0:
            break
        prod = src[:, xf[sel] + t, :] * xa[sel, t][None, :, None]
        buf[:, sel, :] = buf[:, sel, :] + prod
    # vertical pass: first tap sum = beta*buf, later taps sum = sum + beta*buf
    acc = np.zeros((nh, nw, c), dtype=np.float32)
    for t in range(ya.shape[1]):
        sel = np.nonzero(yc > t)[0]
        if sel.size == 0:
            break
        prod = buf[yf[sel] + t, :, :] * ya[sel, t][:, None, None]
        if t == 0:
            acc[sel] = prod
        else:
            acc[sel] = acc[sel] + prod
    return np.ascontiguousarray(np.clip(np.rint(acc), 0, 255).astype(np.uint8))


# ----------------------------------------------------------------------------
# a5  cv2.resize(..., INTER_LINEAR) for 8-bit  (scripts/augmentations.py:45)
# ----------------------------------------------------------------------------
COEF_BITS = 11
COEF_ONE = 1 << COEF_BITS


def linear_table(ssize: int, dsize: int, clamp: bool):
    """Per-destination (s0, s1, a0, a1).  clamp=True is the x axis (cv::resize
    clamps index AND zeroes the fraction); clamp=False is the y axis (fraction
    kept, only the two row indices are clipped)."""
    scale = 1.0 / (dsize / ssize)  # double, as cv::resize computes scale_x
    d = np.arange(dsize, dtype=np.float64)
    f = ((d + 0.5) * scale - 0.5).astype(np.float32)
    s = np.floor(f).astype(np.int32)
    f = (f - s.astype(np.float32)).astype(np.float32)
    if clamp:
        lo = s < 0
        s = np.where(lo, 0, s)
        f = np.where(lo, np.float32(0), f)
        hi = s >= ssize - 1
        s = np.where(hi, ssize - 1, s)
        f = np.where(hi, np.float32(0), f).astype(np.float32)
    a0 = np.rint((np.float32(1.0) - f) * np.float32(COEF_ONE)).astype(np.int32)
    a1 = np.rint(f * np.float32(COEF_ONE)).astype(np.int32)
    s0 = np.clip(s, 0, ssize - 1).astype(np.int32)
    s1 = np.clip(s + 1, 0, ssize - 1).astype(np.int32)
    return s0, s1, a0, a1


def resize_linear(img: np.ndarray, nh: int, nw: int) -> np.ndarray:
    assert img.dtype == np.uint8 and img.ndim == 3
    h, w, c = img.shape
    if nh == h and nw == w:
        return np.ascontiguousarray(img)
    if h == 2 * nh and w == 2 * nw:
        # cv::resize rewrites INTER_LINEAR to INTER_AREA for an exact 2x2 decimation
        return resize_area(img, nh, nw)
    xs0, xs1, xa0, xa1 = linear_table(w, nw, clamp=True)
    ys0, ys1, yb0, yb1 = linear_table(h, nh, clamp=False)
    src = img.astype(np.int32)
    hrow = src[:, xs0, :] * xa0[None, :, None] + src[:, xs1, :] * xa1[None, :, None]
    h4 = hrow >> 4
    t0 = (yb0[:, None, None] * h4[ys0]) >> 16
    t1 = (yb1[:, None, None] * h4[ys1]) >> 16
    out = (t0 + t1 + 2) >> 2
    return np.ascontiguousarray(np.clip(out, 0, 255).astype(np.uint8))


def apply_lowres(img: np.ndarray, factor: float) -> np.ndarray:
    """augmentations.py:41-45: INTER_AREA down to (int(w f), int(h f)), INTER_LINEAR back."""
    h, w = img.shape[:2]
    nh, nw = lowres_small_size(h, w, factor)
    return resize_linear(resize_area(img, nh, nw), h, w)


# ----------------------------------------------------------------------------
# a6-a8  random one-of-three + the two 50 % gates
# ----------------------------------------------------------------------------
def draw_choice() -> int:
    """random.choice(["noise","blur","lowres"]) (augmentations.py:50) -> OP_* code.
    Consumes the Python global `random` stream exactly like the reference."""
    return 1 + _CHOICES.index(_pyrandom.choice(list(_CHOICES)))


def apply_op(img: np.ndarray, op: int) -> np.ndarray:
    if op == OP_NOISE:
        return apply_noise(img, NOISE_SIGMA)
    if op == OP_BLUR:
        return apply_motion_blur(img, BLUR_KERNEL, BLUR_ANGLE_DEG)
    if op == OP_LOWRES:
        return apply_lowres(img, DOWNSCALE_FACTOR)
    return np.ascontiguousarray(img)


def apply_random_corruption(img: np.ndarray) -> np.ndarray:
    """_apply_random_corruption, augmentations.py:48-56."""
    return apply_op(img, draw_choice())


def draw_decisions(n: int, gate: str = "ultralytics", p: float = 0.5):
    """Op-codes for n consecutive hook calls.  gate='ultralytics' applies iff
    random.random() < 0.5 (augmentations.py:93); gate='pil' skips iff
    random.random() > p (augmentations.py:70)."""
    ops = []
    for _ in range(n):
        r = _pyrandom.random()
        applied = (r < 0.5) if gate == "ultralytics" else not (r > p)
        ops.append(draw_choice() if applied else OP_NONE)
    return ops


# ----------------------------------------------------------------------------
# Config 5: detector-input formatting (Ultralytics 8.3.x LetterBox + Format +
# preprocess_batch).  NOT in /root/reference -- builder-written restatement from
# the cv2 primitives above; labelled as such in DESIGN.md.
# ----------------------------------------------------------------------------
def letterbox_geometry(h: int, w: int, out_h: int, out_w: int):
    r = min(out_h / h, out_w / w)
    new_w, new_h = int(round(w * r)), int(round(h * r))
    dw, dh = (out_w - new_w) / 2, (out_h - new_h) / 2
    top, left = int(round(dh - 0.1)), int(round(dw - 0.1))
    return new_h, new_w, top, left


def letterbox_norm_f16(img_bgr: np.ndarray, out_h: int = 640, out_w: int = 640,
                       pad: int = 114) -> np.ndarray:
    """uint8 HWC BGR -> float16 CHW RGB in [0,1]: INTER_LINEAR resize-to-fit, constant
    pad, channel swap, half(float(u8) / 255.f)."""
    h, w = img_bgr.shape[:2]
    new_h, new_w, top, left = letterbox_geometry(h, w, out_h, out_w)
    content = resize_linear(img_bgr, new_h, new_w)
    canvas = np.full((out_h, out_w, 3), pad, dtype=np.uint8)
    canvas[top:top + new_h, left:left + new_w] = content
    rgb = canvas[:, :, ::-1].transpose(2, 0, 1)
    return (rgb.astype(np.float32) / np.float32(255.0)).astype(np.float16)
