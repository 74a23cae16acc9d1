"""CPU tests: pin the oracle (oracle/celt_oracle.c) against the reference's golden vectors and
known-answer tests, and against the compiled reference when oracle/_ref exists."""
import ctypes as C
import os

import numpy as np
import pytest

from conftest import golden_names, load_golden
from oracle import oraclepy

L = oraclepy.lib()


# ---- known answers from opus/celt/tests/test_unit_mathops.c:89-140 -------------------------------------------
def test_bitexact_cos_known_answers():
    chk = 0
    mn, mx, last = 32767, 0, 32767
    min_d, max_d = 32767, 0
    for i in range(64, 16321):
        q = L.co_bitexact_cos(i)
        chk ^= q * i
        d = last - q
        min_d, max_d = min(min_d, d), max(max_d, d)
        mn, mx, last = min(mn, q), max(mx, q), q
    assert chk == 89408644 and max_d == 5 and min_d == 0    # test_unit_mathops.c:105
    assert L.co_bitexact_cos(64) == 32767
    assert L.co_bitexact_cos(16320) == 200
    assert L.co_bitexact_cos(8192) == 23171


def test_bitexact_log2tan_known_answers():
    chk = 0
    fail = False
    mn, mx, last = 15059, -15059, 15059
    min_d, max_d = 15059, 0
    for i in range(64, 8193):
        mid = L.co_bitexact_cos(i)
        side = L.co_bitexact_cos(16384 - i)
        q = L.co_bitexact_log2tan(mid, side)
        chk ^= q * i
        d = last - q
        if q != -L.co_bitexact_log2tan(side, mid):
            fail = True
        min_d, max_d = min(min_d, d), max(max_d, d)
        mn, mx, last = min(mn, q), max(mx, q), q
    assert not fail
    assert chk == 15821257 and max_d == 61 and min_d == -2   # test_unit_mathops.c:132
    assert L.co_bitexact_log2tan(32767, 200) == 15059
    assert L.co_bitexact_log2tan(30274, 12540) == 2611
    assert L.co_bitexact_log2tan(23171, 23171) == 0


def test_isqrt32():
    rng = np.random.default_rng(1)
    vals = np.concatenate([np.arange(1, 2000), rng.integers(1, 2**32 - 1, 20000, dtype=np.uint64), [2**32 - 1]])
    for v in vals:
        g = L.co_isqrt32(int(v))
        assert g * g <= v < (g + 1) * (g + 1)


# ---- PVQ codebook: cwrsi/icwrs identity over the (N,K) reachable in Opus modes (test_unit_cwrs32.c:74-161) ---
PN = [2, 3, 4, 6, 8, 9, 11, 12, 16, 18, 22, 24, 32, 36, 44, 48, 64, 72, 88, 96, 144, 176]
PKMAX = [128, 128, 128, 88, 36, 26, 18, 16, 12, 11, 9, 9, 7, 7, 6, 6, 5, 5, 5, 5, 4, 4]


def test_cwrs_roundtrip():
    rng = np.random.default_rng(7)
    for n, kmax in zip(PN, PKMAX):
        y = (C.c_int * n)()
        for k in range(1, kmax + 1):
            nc = L.co_pvq_v(n, k)
            assert nc > 0
            idxs = set([0, nc - 1, nc // 2]) | set(int(v) for v in rng.integers(0, nc, 24, dtype=np.uint64))
            for i in idxs:
                yy = L.co_cwrsi(n, k, i, y)
                ya = np.array(y[:n])
                assert np.abs(ya).sum() == k
                assert yy == int((ya.astype(np.int64) ** 2).sum())
                assert L.co_icwrs(n, y) == i


# ---- MDCT / FFT vs O(N^2) double-precision transforms, SNR >= 60 dB (test_unit_mdct.c:45-104, test_unit_dft.c) --
def _snr_db(ref, got):
    err = np.sum((ref - got) ** 2)
    return 10 * np.log10(np.sum(ref ** 2) / max(err, 1e-300))


@pytest.mark.parametrize("shift", [0, 1, 2, 3])
def test_fft_vs_dft(shift):
    n = 480 >> shift
    rng = np.random.default_rng(shift)
    x = (rng.integers(0, 32768, 2 * n) - 16384).astype(np.float32)
    data = x.copy()
    L.co_fft(data.ctypes.data_as(C.POINTER(C.c_float)), shift)
    ref = np.fft.fft(x[0::2].astype(np.float64) + 1j * x[1::2].astype(np.float64))
    got = data[0::2].astype(np.float64) + 1j * data[1::2]
    assert _snr_db(np.concatenate([ref.real, ref.imag]), np.concatenate([got.real, got.imag])) >= 60


@pytest.mark.parametrize("shift", [0, 1, 2, 3])
def test_imdct_vs_direct(shift):
    """check_inv of test_unit_mdct.c: out[i] == sum_k in[k] cos(2pi(i+.5+.25N)(k+.5)/N) away from the window."""
    n = 1920 >> shift
    n2 = n // 2
    rng = np.random.default_rng(10 + shift)
    x = (rng.integers(0, 32768, n2) - 16384).astype(np.float32)
    out = np.zeros(n2 + 120 + 8, np.float32)
    L.co_mdct_backward(x.ctypes.data_as(C.POINTER(C.c_float)), out.ctypes.data_as(C.POINTER(C.c_float)), shift, 1)
    # out[60 .. 60+n2) holds the raw inverse transform, except that out[60..120) was multiplied by the rising half of
    # the TDAC window (the previous-frame overlap memory is zero here): undo that and compare everything.
    w = np.sin(.5 * np.pi * np.sin(.5 * np.pi * (np.arange(120) + .5) / 120) ** 2)
    raw = out[60:60 + n2].astype(np.float64)
    raw[:60] /= w[60:]
    i = np.arange(60, 60 + n2)
    k = np.arange(n2)
    pos = i + n // 4 - 60       # position of out[i] inside the full N-sample inverse MDCT
    ref = (np.cos(2 * np.pi * (pos[:, None] + .5 + .25 * n) * (k[None, :] + .5) / n) * x[None, :].astype(np.float64)).sum(1)
    assert _snr_db(ref, raw) >= 60


# ---- golden vectors produced by the reference (tests/golden/make_golden.py) ----------------------------------
@pytest.mark.parametrize("name", golden_names())
def test_oracle_matches_golden(name):
    g = load_golden(name)
    ns = g["packets"].shape[0]
    for s in range(ns):
        pcm, rng, smp = oraclepy.decode_stream(g["packets"][s], g["lens"][s], g["frame_size"], g["dec_channels"])
        assert (smp == g["frame_size"]).all()
        assert (rng == g["dec_rng"][s]).all(), "final range mismatch vs reference decoder"
        if g["dec_channels"] == g["channels"]:
            assert (rng == g["enc_rng"][s]).all(), "final range mismatch vs reference encoder"
        if s < g["pcm"].shape[0]:
            assert np.abs(pcm - g["pcm"][s]).max() <= 2e-6   # float rounding only (spec tolerance is 1e-4)


def test_oracle_rejects_what_is_off_path():
    y = np.zeros(960 * 2, np.float32)
    for toc, exp in ((0x08, -5), (0x78, -5), (0xF9, -5), (0xFB, -5)):     # SILK, hybrid, code 1, code 3
        pk = np.array([[toc, 1, 2, 3, 4, 5, 6, 7]], np.uint8)
        _, _, smp = oraclepy.decode_stream(pk, np.array([8], np.int32), 960, 1)
        assert smp[0] == exp
    pk = np.array([[0xF8, 1, 2, 3, 4, 5, 6, 7]], np.uint8)
    _, _, smp = oraclepy.decode_stream(pk, np.array([8], np.int32), 480, 1)     # 20 ms packet, 10 ms buffer
    assert smp[0] == -2


def test_oracle_vs_live_reference_fuzz(have_ref):
    """Random (garbage) payloads behind valid CELT TOCs: oracle and reference must agree on final range and PCM."""
    if not have_ref:
        pytest.skip("oracle/_ref not built")
    from oracle import refpy
    rng = np.random.default_rng(99)
    for trial in range(60):
        cfg = 16 + int(rng.integers(0, 16))
        stereo = int(rng.integers(0, 2))
        toc = (cfg << 3) | (stereo << 2)
        fs = 120 << (cfg & 3)
        nf = 6
        ln = rng.integers(3, 200, nf).astype(np.int32)
        pk = rng.integers(0, 256, (nf, 200), dtype=np.uint8)
        pk[:, 0] = toc
        for dec_ch in (1, 2):
            ref, rr, rs = refpy.decode_stream(pk, ln, fs, dec_ch)
            o, orr, os_ = oraclepy.decode_stream(pk, ln, fs, dec_ch)
            assert (rs == os_).all() and (rr == orr).all()
            m = np.isfinite(ref)
            assert (m == np.isfinite(o)).all()
            scale = max(1.0, float(np.abs(ref[m]).max()))
            assert np.abs(o[m] - ref[m]).max() <= 1e-5 * scale
