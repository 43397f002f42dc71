"""CPU: sequential replay of the CUDA kernels' tile loops (tests/emu/emu_harness.cpp, built from
the same rod_core.h / rod_tables.h the kernels compile) against the oracle.  Validates the
host-built resize tables, the index math and the per-chunk arithmetic without a GPU.
Bit-exact for blur / lowres / compat noise; Philox field within 2e-3 (float vs float64 math)."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from oracle import corruption_oracle as orc
from tests.helpers import sha, synth

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "emu", "emu_harness.cpp")


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("emu") / "libemu.so")
    subprocess.check_call(["g++", "-O1", "-ffp-contract=off", "-shared", "-fPIC", "-o", out, SRC])
    lib = ctypes.CDLL(out)
    u8p, f32p = ctypes.POINTER(ctypes.c_uint8), ctypes.POINTER(ctypes.c_float)
    lib.emu_blur.argtypes = [u8p, u8p, ctypes.c_int, ctypes.c_int, ctypes.c_long, ctypes.c_long, ctypes.c_int, ctypes.c_int]
    lib.emu_lowres.argtypes = [u8p, u8p, ctypes.c_int, ctypes.c_int, ctypes.c_long, ctypes.c_long, ctypes.c_double, ctypes.c_int]
    lib.emu_lowres_tiled.argtypes = lib.emu_lowres.argtypes
    lib.emu_filter2d.argtypes = [u8p, u8p, ctypes.c_int, ctypes.c_int, ctypes.c_long, ctypes.c_long, f32p, ctypes.c_int]
    lib.emu_lowres_x2w.argtypes = [u8p, u8p, ctypes.c_int, ctypes.c_int, ctypes.c_long, ctypes.c_long, ctypes.c_double, ctypes.c_int]
    lib.emu_noise.argtypes = [u8p, u8p, f32p, f32p, ctypes.c_long, ctypes.c_float, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_uint32]
    lib.emu_noise_table.argtypes = [u8p, u8p, f32p, ctypes.c_long, ctypes.c_float, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_uint32, ctypes.c_int]
    lib.emu_gauss_table.argtypes = [ctypes.c_float, ctypes.POINTER(ctypes.c_int32)]
    lib.emu_letterbox_u8.argtypes = [u8p, ctypes.c_int, ctypes.c_int, ctypes.c_long, u8p, ctypes.c_int, ctypes.c_int, ctypes.c_int]
    return lib


def _p(a, t=ctypes.c_uint8):
    return a.ctypes.data_as(ctypes.POINTER(t))


SHAPES = [(1, 1), (1, 2), (2, 1), (3, 3), (2, 5), (7, 4), (5, 9), (9, 13), (17, 21), (33, 47), (40, 128), (64, 129),
          (65, 255), (70, 257), (31, 1361), (100, 700), (77, 350), (50, 16), (36, 48), (66, 172), (35, 344),
          (34, 346), (64, 1024), (41, 684)]


@pytest.mark.parametrize("k", [9, 3, 5, 15, 31])
def test_emu_blur(emu, k):
    for n, (h, w) in enumerate(SHAPES):
        img = synth(400 + n, h, w)
        want = orc.apply_motion_blur(img, k, 0)
        for phase in (0, 5, 15):
            got = np.zeros_like(img)
            assert emu.emu_blur(_p(img), _p(got), h, w, 3 * w, 3 * w, k, phase) == 0
            assert np.array_equal(got, want), (h, w, k, phase)


@pytest.mark.parametrize("factor", [0.5, 0.25, 0.3, 0.75, 1.0])
def test_emu_lowres(emu, factor):
    for n, (h, w) in enumerate(SHAPES):
        img = synth(500 + n, h, w)
        want = orc.apply_lowres(img, factor)
        for phase in (0, 7):
            got = np.zeros_like(img)
            for fn in (emu.emu_lowres, emu.emu_lowres_tiled):
                got[:] = 0
                rc = fn(_p(img), _p(got), h, w, 3 * w, 3 * w, factor, phase)
                assert rc == 0, (h, w, factor, rc)
                assert np.array_equal(got, want), (h, w, factor, phase)


def test_emu_lowres_visdrone_shape(emu):
    img = synth(12345, 765, 1360)
    got = np.zeros_like(img)
    assert emu.emu_lowres(_p(img), _p(got), 765, 1360, 4080, 4080, 0.5, 0) == 0
    assert np.array_equal(got, orc.apply_lowres(img, 0.5))
    for h, w in [(1080, 1920), (540, 962), (333, 1916)]:
        img = synth(h + w, h, w)
        got = np.zeros_like(img)
        assert emu.emu_lowres(_p(img), _p(got), h, w, 3 * w, 3 * w, 0.5, 0) == 0
        assert np.array_equal(got, orc.apply_lowres(img, 0.5)), (h, w)


def test_emu_noise_compat_and_philox(emu):
    img = synth(600, 37, 53)
    n = img.size
    np.random.seed(9)
    field = orc.draw_noise_field(img.shape, 15)
    got = np.zeros_like(img)
    emu.emu_noise(_p(img), _p(got), _p(field, ctypes.c_float), None, n, 15.0, 0, 0, 0)
    assert np.array_equal(got, orc.add_noise_field(img, field))
    # philox stream, Box-Muller generator: host replay (double math) vs numpy restatement
    out = np.zeros(n, np.float32)
    emu.emu_noise(_p(img), _p(got), None, _p(out, ctypes.c_float), n, 15.0, 0x1234567890ABCDEF, 5, 3)
    want = orc.philox_noise_field(n, 15.0, 0x1234567890ABCDEF, 5, 3, generator="boxmuller")
    assert np.max(np.abs(out - want)) < 2e-3
    assert np.mean(got != orc.add_philox_noise(img, want)) < 1e-3


def test_emu_noise_table_generator(emu):
    """The TABLE generator (default for 3 <= sigma <= 20): the C++ table equals the oracle's scipy-based one entry by
    entry and has the designed moments; the host replay of the kernel's 32-bit two-form arithmetic equals the numpy
    restatement exactly, for Philox4x32-10 and Philox4x32-7."""
    for sigma in (15.0, 3.0, 20.0, 7.25, 4.5):
        tab = np.zeros(256, np.int32)
        assert emu.emu_gauss_table(ctypes.c_float(sigma), _p(tab, ctypes.c_int32)) == 0
        X = orc.gauss_table(sigma)
        assert np.array_equal(tab.astype(np.int64), X), sigma
        assert np.array_equal(X, -X[::-1]) and np.all(np.diff(X) > 0) and 4 * int(X.max()) < 32768
        y = X / (128.0 * sigma)          # unit-variance draws: 2nd / 4th / 6th moments of N(0, 1)
        assert abs((y ** 2).mean() - 1) < 2e-4 and abs((y ** 4).mean() - 3) < 3e-3 and abs((y ** 6).mean() - 15) < 0.03, sigma
    n = 1 << 20
    img = np.random.default_rng(5).integers(0, 256, n, dtype=np.uint8)
    got = np.zeros_like(img)
    out = np.zeros(n, np.float32)
    for rounds, gen in ((10, "auto"), (7, "table7")):
        assert emu.emu_noise_table(_p(img), _p(got), _p(out, ctypes.c_float), n, ctypes.c_float(15.0), 0xABCDEF0123456789, 7, 2, rounds) == 0
        want = orc.philox_noise_field(n, 15.0, 0xABCDEF0123456789, 7, 2, generator=gen)          # auto -> table
        assert np.array_equal(want, orc.philox_noise_field_table(n, 15.0, 0xABCDEF0123456789, 7, 2, rounds))
        assert np.array_equal(out.astype(np.float64), want)
        assert np.array_equal(got, orc.add_philox_noise(img, want))
    assert emu.emu_noise_table(_p(img), _p(got), None, 8, ctypes.c_float(22.0), 0, 0, 0, 10) == 1   # out of the table's range
    assert emu.emu_noise_table(_p(img), _p(got), None, 8, ctypes.c_float(1.0), 0, 0, 0, 10) == 1


def test_table_generator_exact_distribution():
    """The exact law of one element of the table generator -- the 4-fold convolution of the 256-atom table law on
    the 1/256 grid, floored -- against the cell probabilities of floor(N(0, sigma^2)): the chi-square non-centrality a
    16.7 M-sample test would see stays far below its own standard deviation sqrt(2 * cells)."""
    from scipy.special import ndtr
    for sigma, bound in ((15.0, 6.0), (3.0, 12.0), (20.0, 6.0), (8.0, 8.0)):
        X = orc.gauss_table(sigma)
        off = 8192
        p = np.zeros(16384)
        np.add.at(p, X + off, 1.0 / 256)
        s = np.fft.irfft(np.fft.rfft(p, 65536) ** 4, 65536)
        s[s < 0] = 0
        k = np.floor((np.arange(65536) - 4 * off) / 256.0).astype(int)
        pk = np.bincount(k - k.min(), weights=s)
        ks = np.arange(k.min(), k.max() + 1)
        ideal = ndtr((ks + 1) / sigma) - ndtr(ks / sigma)
        N = 16.7e6
        big = ideal * N >= 50
        ncp = N * ((pk[big] - ideal[big]) ** 2 / ideal[big]).sum()
        assert ncp < bound, (sigma, ncp)
        mean, var = (pk * ks).sum(), (pk * ks * ks).sum() - (pk * ks).sum() ** 2
        assert abs(mean + 127.5 / 256) < 1e-4 and abs(np.sqrt(var) - np.sqrt(sigma ** 2 + 1 / 12)) < 4e-4 * sigma, sigma
        assert np.abs(ks[pk > 0]).max() > 5.5 * sigma     # tails beyond 5.5 sigma exist


def test_philox_known_answer():
    # Random123 KAT for philox4x32_10: ctr = key = 0 and the all-ones vector
    r = orc.philox4x32_10(np.zeros((1, 4), np.uint32), np.zeros((1, 2), np.uint32))[0]
    assert [hex(int(x)) for x in r] == ["0x6627e8d5", "0xe169c58d", "0xbc57ac4c", "0x9b00dbd8"]
    r = orc.philox4x32_10(np.full((1, 4), 0xFFFFFFFF, np.uint32), np.full((1, 2), 0xFFFFFFFF, np.uint32))[0]
    assert [hex(int(x)) for x in r] == ["0x408f276d", "0x41c83b0e", "0xa20bc7c6", "0x6d5451fd"]


@pytest.mark.parametrize("shape", [(765, 1360), (720, 1280), (360, 480), (640, 640), (333, 517)])
def test_emu_letterbox(emu, shape):
    h, w = shape
    img = synth(700 + h, h, w)
    canvas = np.zeros((640, 640, 3), np.uint8)
    assert emu.emu_letterbox_u8(_p(img), h, w, 3 * w, _p(canvas), 640, 640, 114) == 0
    want = orc.letterbox_norm_f16(img, 640, 640, 114)
    got = (canvas[:, :, ::-1].transpose(2, 0, 1).astype(np.float32) / np.float32(255)).astype(np.float16)
    assert np.array_equal(got, want)


def test_philox_stream_statistics_cpu():
    """The restated Philox-mode stream on the CPU: moments, tails, the truncation bias and
    clip fractions of add_philox_noise, and the 2^-16 tail-refinement branch."""
    from scipy import stats
    n = 1 << 21
    f = orc.philox_noise_field(n, 15.0, 42, 0, generator="boxmuller")
    z = f / 15.0
    assert abs(z.mean()) < 4 / np.sqrt(n) and abs(z.std() - 1.0) < 3e-3
    assert abs(stats.kurtosis(z)) < 0.02 and abs(stats.skew(z)) < 0.01
    assert stats.kstest(z[::7], "norm").pvalue > 1e-3
    # pairs are uncorrelated, neighbouring groups too
    assert abs(np.corrcoef(z[0::2], z[1::2])[0, 1]) < 5e-3 and abs(np.corrcoef(z[:-8], z[8:])[0, 1]) < 5e-3
    # the stratified 16-bit radius alone stops at sqrt(2 ln 2^17) = 4.855 sigma: the refinement reaches beyond
    big = orc.philox_noise_field(1 << 24, 1.0, 7, 3, generator="boxmuller")
    assert np.abs(big).max() > 4.9 and abs((np.abs(big) > 4.0).mean() - 2 * stats.norm.sf(4.0)) < 2e-5
    # truncation bias and clipping of the fused output
    img = synth(1000, 256, 1024)
    out = orc.add_philox_noise(img, f[:img.size]).astype(np.int32)
    mid = (img >= 70) & (img <= 185)
    d = (out - img)[mid]
    assert abs(d.mean() + 0.5) < 0.05 and abs(d.std() - 15.0) < 0.05
    assert 0.020 < (out == 0).mean() < 0.029 and 0.020 < (out == 255).mean() < 0.029


@pytest.mark.parametrize("generator", ["auto", "table7"])
def test_philox_table_stream_statistics_cpu(generator):
    """Same checks for the TABLE generator (the default at sigma = 15; Philox4x32-10 and -7): k = floor(noise) exactly,
    so the moments are those of floor(N(0, sigma^2)): mean -0.5, variance sigma^2 + 1/12; the histogram of k out to
    4.1 sigma against the exact cell probabilities Phi((k+1)/sigma) - Phi(k/sigma) (chi-square, 8 M samples); tails
    beyond 4.5 sigma present.  JOINT behaviour of the four elements one Philox word makes (a Hadamard mix of four
    draws) and of neighbouring words / groups: lag-1..16 autocorrelation, correlation of squares for every pair inside
    a word, and a coarse 2-D chi-square of (element 4q, element 4q+1) and (element 4q, element 4q+2) against the
    product of the marginals."""
    from scipy import stats
    n = 1 << 23
    k = orc.philox_noise_field(n, 15.0, 42, 0, generator=generator)
    assert np.array_equal(k, np.floor(k))
    assert abs(k.mean() + 0.5) < 4 * 15 / np.sqrt(n) and abs(k.std() - np.sqrt(225 + 1 / 12)) < 0.02
    z = (k + 0.5) / 15.0
    assert abs(stats.kurtosis(z)) < 0.01 and abs(stats.skew(z)) < 0.005
    cells = np.arange(-62, 62)
    expect = n * (stats.norm.cdf((cells + 1) / 15.0) - stats.norm.cdf(cells / 15.0))
    obs = np.array([(k == c).sum() for c in cells], dtype=np.float64)
    chi2 = ((obs - expect) ** 2 / expect).sum()
    assert chi2 < stats.chi2.ppf(1 - 1e-4, len(cells)), chi2
    tol = 4.0 / np.sqrt(n / 4)                                   # 4 sigma of a sample correlation over n/4 pairs
    for lag in (1, 2, 3, 4, 8, 16):
        assert abs(np.corrcoef(z[:-lag], z[lag:])[0, 1]) < tol, lag
        assert abs(np.corrcoef(z[:-lag] ** 2, z[lag:] ** 2)[0, 1]) < tol, lag
    w = z.reshape(-1, 4)                                         # the four elements of one Philox word
    for a in range(4):
        for b in range(a + 1, 4):
            assert abs(np.corrcoef(w[:, a], w[:, b])[0, 1]) < tol, (a, b)
            assert abs(np.corrcoef(w[:, a] ** 2, w[:, b] ** 2)[0, 1]) < tol, (a, b)      # independent, not just uncorrelated
            assert abs(np.corrcoef(np.abs(w[:, a]), np.abs(w[:, b]))[0, 1]) < tol, (a, b)
    edges = stats.norm.ppf(np.linspace(0, 1, 13))[1:-1]          # 12 x 12 equiprobable cells
    for a, b in ((0, 1), (0, 2), (1, 3), (2, 3)):
        ia, ib = np.searchsorted(edges, w[:, a]), np.searchsorted(edges, w[:, b])
        tab2 = np.bincount(ia * 12 + ib, minlength=144).reshape(12, 12).astype(np.float64)
        exp2 = np.outer(tab2.sum(1), tab2.sum(0)) / tab2.sum()
        chi2d = ((tab2 - exp2) ** 2 / exp2).sum()
        assert chi2d < stats.chi2.ppf(1 - 1e-4, 121), (a, b, chi2d)
    assert abs((np.abs(z) > 4.0).mean() - 2 * stats.norm.sf(4.0)) < 2e-5 and np.abs(z).max() > 4.5
    img = synth(1000, 256, 1024)
    out = orc.add_philox_noise(img, k[:img.size]).astype(np.int32)
    mid = (img >= 70) & (img <= 185)
    d = (out - img)[mid]
    assert abs(d.mean() + 0.5) < 0.05 and abs(d.std() - 15.0) < 0.05
    assert 0.020 < (out == 0).mean() < 0.029 and 0.020 < (out == 255).mean() < 0.029


X2W_SHAPES = [(765, 1360), (360, 480), (100, 8), (9, 4), (2, 4), (5, 12), (64, 64), (65, 128), (131, 36), (201, 1400),
              (97, 1916), (540, 960), (33, 2000), (40, 20), (77, 1364), (50, 240), (51, 244), (52, 248)]


@pytest.mark.parametrize("band_rows", [24, 56, 512])
def test_emu_lowres_warp_marching(emu, band_rows):
    """lowres_x2w_kernel replayed lane by lane on the CPU: strips of 30 chunks with halo lanes, neighbour pixels by
    shuffle, border replication, two-pixel last chunks (w % 8 == 4), exact-2x and odd heights, pitched rows."""
    for i, (h, w) in enumerate(X2W_SHAPES):
        img = synth(700 + i, h, w)
        want = orc.apply_lowres(img, 0.5)
        pitch = 3 * w + (0 if i % 3 else 20)
        buf = np.full((h, pitch), 0xAB, np.uint8)
        buf[:, :3 * w] = img.reshape(h, 3 * w)
        got = np.zeros_like(img)
        rc = emu.emu_lowres_x2w(_p(buf), _p(got), h, w, pitch, 3 * w, 0.5, band_rows)
        assert rc == 0, (h, w, rc)
        assert np.array_equal(got, want), (h, w, band_rows)


@pytest.mark.parametrize("band_rows", [24, 56, 512])
def test_emu_lowres_float_staged_kernel(emu, band_rows):
    """lowres_x2f_kernel's arithmetic (carried pair sums, paired vertical stage with the packed integer finish) replayed on
    the CPU against the oracle on the odd-height shapes of X2W_SHAPES."""
    emu.emu_lowres_x2f.argtypes = emu.emu_lowres_x2w.argtypes
    n_run = 0
    for i, (h, w) in enumerate(X2W_SHAPES):
        if h % 2 == 0:
            continue
        img = synth(700 + i, h, w)
        if i % 2:
            img = (img > 127).astype(np.uint8) * 255
        want = orc.apply_lowres(img, 0.5)
        pitch = 3 * w + (0 if i % 3 else 20)
        buf = np.full((h, pitch), 0xAB, np.uint8)
        buf[:, :3 * w] = img.reshape(h, 3 * w)
        got = np.zeros_like(img)
        rc = emu.emu_lowres_x2f(_p(buf), _p(got), h, w, pitch, 3 * w, 0.5, band_rows)
        assert rc == 0, (h, w, rc)
        assert np.array_equal(got, want), (h, w, band_rows, int((got != want).sum()))
        n_run += 1
    assert n_run >= 6


@pytest.mark.parametrize("band_rows", [24, 56, 512])
def test_emu_lowres_regular_three_tap_kernel(emu, band_rows):
    """lowres_x2h_kernel's loop over low-res rows (fixed source-row pairs, carried pair sums, the per-low-res-row emission
    schedule clipped to the band) replayed on the CPU against the oracle on odd heights; every row is written exactly once."""
    emu.emu_lowres_x2h.argtypes = emu.emu_lowres_x2w.argtypes
    n_run = 0
    shapes = [s for s in X2W_SHAPES if s[0] % 2 == 1] + [(3, 8), (5, 244), (7, 36), (765, 1360), (1079, 1916), (1999, 16), (2001, 8)]
    for i, (h, w) in enumerate(shapes):
        img = synth(900 + i, h, w)
        if i % 2:
            img = (img > 127).astype(np.uint8) * 255
        pitch = 3 * w + (0 if i % 3 else 20)
        buf = np.full((h, pitch), 0xAB, np.uint8)
        buf[:, :3 * w] = img.reshape(h, 3 * w)
        got = np.full_like(img, 0x5A)
        rc = emu.emu_lowres_x2h(_p(buf), _p(got), h, w, pitch, 3 * w, 0.5, band_rows)
        if rc == 3:     # not eligible (h < 3, or tap rows outside the regular pattern): the float-tap kernel takes it
            continue
        assert rc == 0, (h, w, rc)
        want = orc.apply_lowres(img, 0.5)
        assert np.array_equal(got, want), (h, w, band_rows, int((got != want).sum()))
        n_run += 1
    assert n_run >= 8


X2G_SHAPES = [(765, 1361), (360, 481), (100, 9), (9, 5), (2, 3), (5, 13), (64, 65), (65, 129), (131, 37), (201, 1401), (97, 1917),
              (540, 961), (33, 1999), (40, 21), (77, 1363), (50, 241), (51, 243), (52, 247), (8, 7), (3, 11), (10, 15), (64, 17)]


@pytest.mark.parametrize("band_rows", [24, 56, 512])
def test_emu_lowres_odd_width_kernel(emu, band_rows):
    """lowres_x2g_kernel's arithmetic (odd widths at factor 0.5: three-tap float INTER_AREA x pass, carried tap row,
    border replication for chunks with 0..3 valid low-res pixels, coefficient + slip driven INTER_LINEAR x stage, general
    y stage) replayed lane by lane on the CPU against the oracle; odd and even heights, uniform and binary content."""
    emu.emu_lowres_x2g.argtypes = [ctypes.POINTER(ctypes.c_uint8), ctypes.POINTER(ctypes.c_uint8), ctypes.c_int, ctypes.c_int,
                                   ctypes.c_long, ctypes.c_long, ctypes.c_int]
    n_run = 0
    for i, (h, w) in enumerate(X2G_SHAPES):
        if band_rows != 56 and h * w > 300000:
            continue
        img = synth(1100 + i, h, w)
        if i % 4 == 1:
            img = (img > 127).astype(np.uint8) * 255
        want = orc.apply_lowres(img, 0.5)
        pitch = 3 * w + (0 if i % 3 else 7)
        buf = np.full((h, pitch), 0xAB, np.uint8)
        buf[:, :3 * w] = img.reshape(h, 3 * w)
        got = np.zeros_like(img)
        rc = emu.emu_lowres_x2g(_p(buf), _p(got), h, w, pitch, 3 * w, band_rows)
        if rc == 3 and min(h, w) <= 3:
            continue                      # integer-scale corner cases (resizeAreaFast_) are not this kernel's
        assert rc == 0, (h, w, rc)
        assert np.array_equal(got, want), (h, w, band_rows, int((got != want).sum()))
        n_run += 1
    assert n_run >= (19 if band_rows == 56 else 14)
    assert emu.emu_lowres_x2g(_p(buf), _p(got), 64, 64, 192, 192, 56) == 3      # even width: not this kernel's


X2P_SHAPES = [(360, 480), (100, 8), (2, 4), (4, 4), (64, 64), (130, 36), (200, 1400), (96, 1916), (540, 960), (34, 2000), (40, 20),
              (76, 1364), (50, 240), (52, 244), (52, 248), (1078, 1916), (1050, 1400)]


@pytest.mark.parametrize("band_rows", [24, 56, 512])
def test_emu_lowres_odd_width_regular_kernel(emu, band_rows):
    """lowres_x2i_kernel: the odd-width arithmetic inside the loop over low-res rows (odd heights: carried tap row; even
    heights: two fresh tap rows) with the per-low-res-row emission schedule, replayed on the CPU against the oracle."""
    emu.emu_lowres_x2i.argtypes = [ctypes.POINTER(ctypes.c_uint8), ctypes.POINTER(ctypes.c_uint8), ctypes.c_int, ctypes.c_int,
                                   ctypes.c_long, ctypes.c_long, ctypes.c_int]
    n_carry = n_plain = 0
    for i, (h, w) in enumerate(X2G_SHAPES):
        img = synth(1100 + i, h, w)
        if i % 2:
            img = (img > 127).astype(np.uint8) * 255
        pitch = 3 * w + (0 if i % 3 else 13)
        buf = np.full((h, pitch), 0xAB, np.uint8)
        buf[:, :3 * w] = img.reshape(h, 3 * w)
        got = np.full_like(img, 0x5A)
        rc = emu.emu_lowres_x2i(_p(buf), _p(got), h, w, pitch, 3 * w, band_rows)
        if rc == 3:
            continue
        assert rc == 0, (h, w, rc)
        want = orc.apply_lowres(img, 0.5)
        assert np.array_equal(got, want), (h, w, band_rows, int((got != want).sum()))
        n_carry += h % 2
        n_plain += 1 - h % 2
    assert n_carry >= 5 and n_plain >= 5


@pytest.mark.parametrize("band_rows", [24, 56, 512])
def test_emu_lowres_packed_kernel(emu, band_rows):
    """lowres_x2p_kernel (even w and h: the all-integer packed pipeline) replayed lane by lane on the CPU against the
    oracle: halo lanes, neighbour words by shuffle, border replication, two-pixel last chunks (w % 8 == 4), band
    boundaries, pitched rows; uniform and binary {0, 255} content (every rounding boundary of the >> 2 stages)."""
    emu.emu_lowres_x2p.argtypes = [ctypes.POINTER(ctypes.c_uint8), ctypes.POINTER(ctypes.c_uint8), ctypes.c_int, ctypes.c_int,
                                   ctypes.c_long, ctypes.c_long, ctypes.c_int]
    for i, (h, w) in enumerate(X2P_SHAPES):
        if band_rows != 56 and h * w > 600000:
            continue
        img = synth(900 + i, h, w)
        if i % 4 == 1:
            img = (img > 127).astype(np.uint8) * 255
        want = orc.apply_lowres(img, 0.5)
        pitch = 3 * w + (0 if i % 3 else 20)
        buf = np.full((h, pitch), 0xAB, np.uint8)
        buf[:, :3 * w] = img.reshape(h, 3 * w)
        got = np.zeros_like(img)
        rc = emu.emu_lowres_x2p(_p(buf), _p(got), h, w, pitch, 3 * w, band_rows)
        assert rc == 0, (h, w, rc)
        assert np.array_equal(got, want), (h, w, band_rows, int((got != want).sum()))
    assert emu.emu_lowres_x2p(_p(buf), _p(got), 765, 1360, 4080, 4080, 56) == 3      # odd height: not this kernel's
    assert emu.emu_lowres_x2p(_p(buf), _p(got), 100, 1362, 4086, 4086, 56) == 3      # w % 4 != 0


def test_emu_filter2d_general_angles(emu):
    """filter2d_kernel's tiling / halo / tap order / FMA-vs-tail arithmetic replayed on the CPU against the outputs
    the reference produced at angles != 0 (tests/golden/golden_angles.npz)."""
    import json
    here = os.path.join(HERE, "golden")
    g = np.load(os.path.join(here, "golden_angles.npz"))
    meta = json.load(open(os.path.join(here, "golden_angles.json")))
    for k, ang in meta["cases"]:
        kern = np.ascontiguousarray(g[f"kernel_{k}_{ang}"], dtype=np.float32)
        for i, (h, w) in enumerate(meta["shapes"]):
            img = synth(300 + i, h, w)
            got = np.zeros_like(img)
            assert emu.emu_filter2d(_p(img), _p(got), h, w, 3 * w, 3 * w, _p(kern, ctypes.c_float), k) == 0
            assert np.array_equal(got, g[f"out_{k}_{ang}_uniform_{h}x{w}"]), (k, ang, h, w)
    name, seed, h, w = meta["big"][2]  # several tiles in x, 2-byte row tail
    kern = np.ascontiguousarray(g["kernel_9_45"], dtype=np.float32)
    img = synth(seed, h, w)
    got = np.zeros_like(img)
    emu.emu_filter2d(_p(img), _p(got), h, w, 3 * w, 3 * w, _p(kern, ctypes.c_float), 9)
    assert sha(got) == meta["sha"][f"9_45_{name}"]


def jpeg_header(h, w):
    """The bytes SOI .. SOS that OpenCV writes for an h x w BGR image with its default parameters."""
    import cv2
    ok, buf = cv2.imencode(".jpg", np.zeros((h, w, 3), np.uint8))
    assert ok
    b = buf.tobytes()
    i = 2
    while True:
        assert b[i] == 0xFF
        L = (b[i + 2] << 8) | b[i + 3]
        if b[i + 1] == 0xDA:
            return b[:i + 2 + L]
        i += 2 + L


def test_emu_jpeg_encoder_matches_cv2(emu):
    """The encoder arithmetic of csrc/rod_jpeg.h (colour conversion, h2v2 chroma, islow DCT, reciprocal quantisation,
    Huffman coding with libjpeg's dummy blocks, stuffing) replayed on the CPU: the file must equal cv2.imencode's byte for
    byte -- sizes that are not multiples of 8 / 16, uniform noise (every coefficient busy), smooth and constant content."""
    import cv2
    emu.emu_jpeg_encode.argtypes = [ctypes.POINTER(ctypes.c_uint8), ctypes.c_int, ctypes.c_int, ctypes.c_long,
                                    ctypes.POINTER(ctypes.c_uint8), ctypes.c_long, ctypes.POINTER(ctypes.c_uint8), ctypes.c_long]
    emu.emu_jpeg_encode.restype = ctypes.c_long
    shapes = [(16, 16), (8, 8), (1, 1), (37, 53), (64, 48), (17, 33), (100, 9), (9, 100), (120, 200), (97, 133), (765, 1360), (38, 40)]
    for i, (h, w) in enumerate(shapes):
        variants = [synth(8000 + i, h, w)]
        if h * w < 100000:
            variants += [cv2.GaussianBlur(synth(8100 + i, h, w), (0, 0), 2.5), np.full((h, w, 3), 255 if i % 2 else 0, np.uint8),
                         (synth(8200 + i, h, w) > 127).astype(np.uint8) * 255]
        hdr = np.frombuffer(jpeg_header(h, w), np.uint8).copy()
        for v, img in enumerate(variants):
            want = cv2.imencode(".jpg", img)[1].tobytes()
            pitch = 3 * w + (0 if v % 2 else 7)
            buf = np.full((h, pitch), 0xAB, np.uint8)
            buf[:, :3 * w] = img.reshape(h, 3 * w)
            out = np.zeros(3 * h * w + 4096, np.uint8)
            n = emu.emu_jpeg_encode(_p(buf), h, w, pitch, _p(hdr), len(hdr), _p(out), len(out))
            assert n > 0, (h, w, v, n)
            got = out[:n].tobytes()
            assert len(got) == len(want) and got == want, (h, w, v, len(got), len(want),
                                                           next((k for k in range(min(len(got), len(want))) if got[k] != want[k]), -1))


def test_emu_jpeg_encoder_matches_golden(emu):
    """... and against the hashes tests/golden/make_golden.py recorded from cv2.imencode in the build container
    (golden_jpeg.json), so the check does not rest on the local OpenCV build alone."""
    import hashlib
    import json
    import cv2
    emu.emu_jpeg_encode.argtypes = [ctypes.POINTER(ctypes.c_uint8), ctypes.c_int, ctypes.c_int, ctypes.c_long,
                                    ctypes.POINTER(ctypes.c_uint8), ctypes.c_long, ctypes.POINTER(ctypes.c_uint8), ctypes.c_long]
    emu.emu_jpeg_encode.restype = ctypes.c_long
    g = json.load(open(os.path.join(HERE, "golden", "golden_jpeg.json")))
    for seed, h, w, kind in g["cases"]:
        img = synth(seed, h, w)
        if kind == "smooth":
            img = cv2.GaussianBlur(img, (0, 0), 2.5)
        elif kind == "binary":
            img = (img > 127).astype(np.uint8) * 255
        hdr = np.frombuffer(jpeg_header(h, w), np.uint8).copy()
        out = np.zeros(3 * h * w + 4096, np.uint8)
        n = emu.emu_jpeg_encode(_p(np.ascontiguousarray(img)), h, w, 3 * w, _p(hdr), len(hdr), _p(out), len(out))
        assert n > 0
        assert hashlib.sha256(out[:n].tobytes()).hexdigest() == g["sha"][f"{seed}_{h}x{w}_{kind}"], (seed, h, w, kind)


def _emu_decode(emu, data, mode=1, rounds=None):
    """mode 0: sequential scan decoder; mode 1: the device's passes (self-synchronising subsequences, slowest schedule)"""
    emu.emu_jpeg_decode_mode.argtypes = [ctypes.POINTER(ctypes.c_uint8), ctypes.c_long, ctypes.POINTER(ctypes.c_uint8), ctypes.c_long,
                                         ctypes.POINTER(ctypes.c_int), ctypes.c_int]
    b = np.frombuffer(data, np.uint8).copy()
    out = np.zeros(2100 * 2100 * 3, np.uint8)
    hw = (ctypes.c_int * 3)()
    rc = emu.emu_jpeg_decode_mode(_p(b), len(b), _p(out), len(out), hw, mode)
    if rounds is not None:
        rounds.append(hw[2])
    return rc, (out[:hw[0] * hw[1] * 3].reshape(hw[0], hw[1], 3).copy() if rc == 0 else None)


def test_emu_jpeg_decoder_matches_cv2(emu):
    """The decoder arithmetic of csrc/rod_jpegdec.h (Huffman decoding, dequantisation + islow IDCT, fancy h2v2 chroma
    upsampling with libjpeg's edge rows / columns, YCbCr -> BGR) and the host side of csrc/rod_jpegdec_host.h (markers,
    derived tables, unstuffing) replayed on the CPU: the pixels must equal cv2.imdecode's -- every small size from 5 pixels
    of width on (MCU-partial right / bottom edges, one-row images), several qualities incl. 100, optimised Huffman tables,
    noise / smooth / flat content, VisDrone frame sizes."""
    import cv2
    rng = np.random.default_rng(11)
    cases = [(h, w) for h in (1, 2, 3, 7, 8, 9, 15, 16, 17, 31, 33) for w in (5, 6, 7, 8, 9, 15, 16, 17, 18, 31, 32, 33, 47, 65)]
    cases += [(765, 1360), (540, 960), (1078, 1916), (97, 133), (300, 180)]
    for i, (h, w) in enumerate(cases):
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        kind = i % 4
        if kind == 1:
            img = cv2.GaussianBlur(img, (0, 0), 2.0)
        elif kind == 2:
            img[:] = rng.integers(0, 256, 3, dtype=np.uint8)
        params = [[], [cv2.IMWRITE_JPEG_QUALITY, 100], [cv2.IMWRITE_JPEG_QUALITY, int(rng.integers(5, 95))], [cv2.IMWRITE_JPEG_OPTIMIZE, 1]][(i // 4) % 4]
        # chroma layouts: 4:2:0 (OpenCV's default) mostly, 4:2:2, 4:4:4, greyscale
        lay = (i // 3) % 5
        if lay == 3:
            params = params + [cv2.IMWRITE_JPEG_SAMPLING_FACTOR, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_422]
        elif lay == 4:
            params = params + [cv2.IMWRITE_JPEG_SAMPLING_FACTOR, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_444]
        elif i % 11 == 5:
            img = np.ascontiguousarray(img[:, :, 0])
        if i % 5 == 2:   # restart markers: every interval is a bit stream of its own (DC predictions from 0)
            params = params + [cv2.IMWRITE_JPEG_RST_INTERVAL, 1 + i % 7]
        ok, enc = cv2.imencode(".jpg", img, params)
        assert ok
        want = cv2.imdecode(enc, cv2.IMREAD_COLOR)
        for mode in (0, 1):
            rc, got = _emu_decode(emu, enc.tobytes(), mode)
            assert rc == 0, (h, w, params, mode, rc)
            assert np.array_equal(got, want), (h, w, kind, params, mode)


def test_emu_jpeg_decoder_self_synchronisation(emu):
    """The parallel Huffman stage on streams that synchronise badly: constant images (every block is the same few bits, a
    decoder started at a wrong block of the MCU never falls into step: the states propagate one subsequence per round),
    uniform noise at quality 100 (long codes), and a corrupt stream, which must be reported, not decoded."""
    import cv2
    rounds = []
    for img, params in ((np.full((300, 500, 3), 77, np.uint8), []), (np.zeros((64, 2000, 3), np.uint8), [cv2.IMWRITE_JPEG_OPTIMIZE, 1]),
                        (synth(8400, 240, 320), [cv2.IMWRITE_JPEG_QUALITY, 100]), (synth(8401, 200, 200) // 128 * 255, [])):
        enc = cv2.imencode(".jpg", img, params)[1]
        rc, got = _emu_decode(emu, enc.tobytes(), 1, rounds)
        assert rc == 0 and np.array_equal(got, cv2.imdecode(enc, cv2.IMREAD_COLOR))
    assert max(rounds) > 8, rounds          # the slow-propagation case was really exercised
    whole = bytearray(cv2.imencode(".jpg", synth(8402, 120, 160))[1].tobytes())
    scan = bytes(whole).index(b"\xff\xda") + 14
    cut = bytes(whole[:scan + (len(whole) - scan) // 3]) + b"\xff\xd9"
    assert _emu_decode(emu, cut, 1)[0] in (4, 5) and _emu_decode(emu, cut, 0)[0] in (4, 5)


def test_emu_jpeg_decoder_reports_other_layouts(emu):
    """Files the device decoder does not take are reported, not approximated: vertical-only / 4:1:1 chroma sampling,
    progressive, tiny widths of subsampled files (libjpeg-turbo's upsampler reads its padding there),
    truncated files, other formats."""
    import cv2
    img = synth(8300, 64, 64)
    for params in ([cv2.IMWRITE_JPEG_SAMPLING_FACTOR, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_440], [cv2.IMWRITE_JPEG_SAMPLING_FACTOR, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_411],
                   [cv2.IMWRITE_JPEG_PROGRESSIVE, 1]):
        assert _emu_decode(emu, cv2.imencode(".jpg", img, params)[1].tobytes())[0] == 2, params
    assert _emu_decode(emu, cv2.imencode(".jpg", img[:, :4])[1].tobytes())[0] == 2
    small444 = cv2.imencode(".jpg", img[:3, :2], [cv2.IMWRITE_JPEG_SAMPLING_FACTOR, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_444])[1]
    rc, got = _emu_decode(emu, small444.tobytes())
    assert rc == 0 and np.array_equal(got, cv2.imdecode(small444, cv2.IMREAD_COLOR))   # no subsampling: any width
    whole = cv2.imencode(".jpg", img)[1].tobytes()
    assert _emu_decode(emu, whole[:len(whole) // 2])[0] == 3        # no EOI behind the scan
    assert _emu_decode(emu, whole[:100])[0] == 1                      # header cut short
    assert _emu_decode(emu, cv2.imencode(".png", img)[1].tobytes())[0] == 1
    assert _emu_decode(emu, b"not a jpeg")[0] == 1


def test_emu_jpeg_decoder_matches_golden(emu):
    """... and against the files and pixel hashes recorded in the build container (golden_jpegdec.json: OpenCV's decoder on
    4:2:0 / 4:2:2 / 4:4:4 / greyscale files, optimised tables, restart intervals), so the check does not rest on the local
    OpenCV build."""
    import base64
    import hashlib
    import json
    g = json.load(open(os.path.join(HERE, "golden", "golden_jpegdec.json")))
    assert len(g["cases"]) >= 9
    for c in g["cases"]:
        data = base64.b64decode(c["file_b64"])
        for mode in (0, 1):
            rc, got = _emu_decode(emu, data, mode)
            assert rc == 0 and list(got.shape) == c["shape"], (c["name"], mode, rc)
            assert hashlib.sha256(np.ascontiguousarray(got).tobytes()).hexdigest() == c["pixels_sha256"], (c["name"], mode)
