"""CPU test: the product's per-thread / per-warp / per-block DEVICE code (opus_codec_b200/csrc/*.cuh), compiled by g++
with one emulated lane (tests/host_emul/emul.cpp), must agree with the oracle.  This validates the kernels' logic on the
GPU-less build box; the GPU tests then only have to catch synchronisation and hardware-specific problems."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, golden_names, load_golden, plc_golden_names, load_plc_golden, multiframe_stream, pathological_pcm
from oracle import oraclepy

EMU = os.path.join(ROOT, "tests", "host_emul")


os.environ["OB_EMUL_POISON"] = "1"      # the encoder's work area is filled with NaN patterns before every frame (see emul.cpp)


@pytest.fixture(scope="module")
def emul():
    from conftest import emul_lib
    return emul_lib()


def P(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def same_packets(a, al, b, bl):
    """Packets equal byte for byte over their lengths (what lies beyond a packet's length in its slot is not part of the packet: the
    reference leaves stale raw-bit bytes there when a VBR frame shrinks, the device code builds the payload in shared memory)."""
    if not (np.asarray(al) == np.asarray(bl)).all():
        return False
    m = np.arange(a.shape[1])[None, :] < np.maximum(np.asarray(al), 0)[:, None]
    return bool(((a == b) | ~m).all())


@pytest.mark.parametrize("name", golden_names())
def test_device_code_single_lane_matches_oracle(emul, name):
    g = load_golden(name)
    fs, dc = g["frame_size"], g["dec_channels"]
    for s in range(min(2, g["packets"].shape[0])):
        pk = np.ascontiguousarray(g["packets"][s]); ln = np.ascontiguousarray(g["lens"][s]); nf = pk.shape[0]
        opcm, orng, osmp = oraclepy.decode_stream(pk, ln, fs, dc)
        pcm = np.zeros((nf, fs * dc), np.float32); rng = np.zeros(nf, np.uint32); smp = np.zeros(nf, np.int32)
        emul.emul_decode_stream(P(pk, C.c_ubyte), P(ln, C.c_int), pk.shape[1], nf, fs, dc, P(pcm, C.c_float), P(rng, C.c_uint32), P(smp, C.c_int), None)
        assert (smp == osmp).all() and (rng == orng).all()
        assert np.abs(pcm - opcm).max() <= 1e-6


# ---- encoder: per-stream device code vs the reference encoder's pure-C build (oracle/_ref/libopus_ref_c.so) ---------------------------
ENC_CELT = [(1, 64000, 960, 159, 0, 10), (2, 96000, 960, 239, 0, 10), (2, 96000, 960, 1275, 1, 10), (1, 64000, 960, 1275, 2, 10),
            (2, 64000, 480, 79, 0, 10), (1, 48000, 240, 29, 0, 10), (2, 96000, 120, 29, 0, 10), (1, 24000, 960, 59, 0, 5),
            (2, 510000, 960, 1275, 0, 10), (2, 96000, 960, 239, 0, 0)]


@pytest.mark.parametrize("base", plc_golden_names())
def test_concealment_device_code_matches_reference_c_build(emul, base):
    """Lost packets, DTX payloads, the noise/pitch concealment switch and the recovery frames after a loss: the product's device
    code (one emulated lane) against PCM produced by the reference's pure-C build (fixtures: tests/golden/make_golden_plc.py)."""
    g, p = load_golden(base), load_plc_golden(base)
    fs, dc = g["frame_size"], g["dec_channels"]
    for s in range(p["lens"].shape[0]):
        nf = p["lens"].shape[1]
        pk = np.ascontiguousarray(g["packets"][s, :nf]); ln = np.ascontiguousarray(p["lens"][s])
        pcm = np.zeros((nf, fs * dc), np.float32); rng = np.zeros(nf, np.uint32); smp = np.zeros(nf, np.int32)
        emul.emul_decode_stream(P(pk, C.c_ubyte), P(ln, C.c_int), pk.shape[1], nf, fs, dc, P(pcm, C.c_float), P(rng, C.c_uint32), P(smp, C.c_int), None)
        assert (smp == p["samples"][s]).all() and (rng == p["ranges"][s]).all()
        assert (ln <= 2).sum() >= 10
        assert np.abs(pcm - p["pcm_c"][s]).max() <= 1e-6


@pytest.mark.parametrize("name", ["cfg2_mono_20ms_64k_cbr", "stereo_20ms_vbr_96k", "cfg4_stereo_5ms_96k", "cfg4_mono_2p5ms_64k"])
def test_multiframe_packets_device_code_matches_reference(emul, have_ref, name):
    """TOC codes 1, 2 and 3 (CBR/VBR, padding, a DTX frame inside) through the framing pass of the device code vs the reference."""
    if not have_ref:
        pytest.skip("oracle/_ref not built")
    from oracle import refpy
    g = load_golden(name)
    fs, dc = g["frame_size"], g["dec_channels"]
    pk, ln = multiframe_stream(g, 0, 14)
    ln[4] = 0                                              # and a lost packet in between: 3 frames' worth concealed in pieces
    nf, slot = pk.shape[0], 3 * fs
    ref, rr, rs = refpy.decode_stream(pk, ln, slot, dc, pure_c=True)
    pcm = np.zeros((nf, slot * dc), np.float32); rng = np.zeros(nf, np.uint32); smp = np.zeros(nf, np.int32)
    emul.emul_decode_stream(P(pk, C.c_ubyte), P(ln, C.c_int), pk.shape[1], nf, slot, dc, P(pcm, C.c_float), P(rng, C.c_uint32), P(smp, C.c_int), None)
    assert (smp == rs).all() and (rng == rr).all(), (smp, rs)
    assert set(int(b) & 3 for b in pk[:, 0]) == {0, 1, 2, 3} or name == "stereo_20ms_vbr_96k"
    for f in range(nf):
        assert np.abs(pcm[f, :rs[f] * dc] - ref[f, :rs[f] * dc]).max() <= 1e-6


def test_soft_clip_and_int16_rounding_match_reference(emul, have_ref):
    """The int16 back end (opus_pcm_soft_clip with its packet-to-packet gain + FLOAT2INT16) on signals that clip in every way the
    reference distinguishes: before the first zero crossing, across packet boundaries, beyond +-2, not at all."""
    if not have_ref:
        pytest.skip("oracle/_ref not built")
    from oracle import refpy
    rng = np.random.default_rng(5)
    for ch in (1, 2):
        mem_ref = np.zeros(ch, np.float32); mem = np.zeros(2, np.float32)
        for k in range(60):
            n = int(rng.choice([120, 240, 480, 960, 2880]))
            t = np.arange(n)[:, None]
            amp = rng.choice([0.3, 0.9, 1.05, 1.4, 2.6]) * (1 + 0.3 * rng.random((1, ch)))
            x = (amp * np.sin(2 * np.pi * (rng.integers(1, 40) / n) * t + rng.random((1, ch)) * 6.28) + 0.05 * rng.standard_normal((n, ch))).astype(np.float32)
            a = np.ascontiguousarray(x.reshape(-1)); b = a.copy()
            refpy.soft_clip(a, ch, mem_ref)
            want = np.clip(np.rint(np.clip(a * np.float32(32768), -32768, 32767)), -32768, 32767).astype(np.int16)
            out = np.zeros(n * ch, np.int16)
            emul.emul_packet_to_int16(P(b, C.c_float), P(out, C.c_int16), n, ch, P(mem, C.c_float))
            assert np.array_equal(a, b), (ch, k)
            assert np.array_equal(out, want) and np.array_equal(mem[:ch], mem_ref)


@pytest.mark.parametrize("ch,br,fs,vbr,cx", [(2, 96000, 960, 0, 10), (1, 32000, 480, 1, 10), (2, 64000, 240, 2, 5)])
def test_encoder_pathological_input_bit_identical_to_reference(emul, have_ref, ch, br, fs, vbr, cx):
    """Digital silence, DC, 6x full scale, 1e-7 noise, bursts, NaN / Inf samples, Nyquist energy, square waves."""
    if not have_ref:
        pytest.skip("oracle/_ref not built")
    from oracle import refpy
    L = refpy.lib_c()
    u8p, i32p, u32p, f32p = (C.POINTER(C.c_ubyte), C.POINTER(C.c_int), C.POINTER(C.c_uint32), C.POINTER(C.c_float))
    L.ref_encode_stream.argtypes = [f32p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, u8p, C.c_int, i32p, u32p]
    rng = np.random.default_rng(12)
    for s in range(8):
        pcm = pathological_pcm(s, 24000, ch, rng)
        nf = pcm.size // (fs * ch)
        a = np.zeros((nf, 1275), np.uint8); al = np.zeros(nf, np.int32); ar = np.zeros(nf, np.uint32)
        b = np.zeros((nf, 1275), np.uint8); bl = np.zeros(nf, np.int32); br_ = np.zeros(nf, np.uint32)
        assert L.ref_encode_stream(P(pcm, C.c_float), nf, fs, ch, 2051, br, vbr, cx, P(a, C.c_ubyte), 1275, P(al, C.c_int), P(ar, C.c_uint32)) == 0
        assert emul.emul_opus_encode_stream(P(pcm, C.c_float), nf, fs, ch, br, vbr, cx, P(b, C.c_ubyte), 1275, P(bl, C.c_int), P(br_, C.c_uint32)) == 0
        assert (ar == br_).all() and same_packets(a, al, b, bl), s


def _fuzz_streams(seed, trials, nf=10, with_loss=True):
    """Random (garbage) payloads behind valid CELT TOCs, optionally with lost packets and DTX payloads in between."""
    rng = np.random.default_rng(seed)
    for _ in range(trials):
        cfg = 16 + int(rng.integers(0, 16))
        toc = (cfg << 3) | (int(rng.integers(0, 2)) << 2)
        fs = 120 << (cfg & 3)
        ln = rng.integers(3, 200, nf).astype(np.int32)
        if with_loss:
            ln[rng.random(nf) < 0.2] = 0
            ln[rng.random(nf) < 0.05] = 1
        pk = rng.integers(0, 256, (nf, 200), dtype=np.uint8)
        pk[:, 0] = toc
        yield pk, ln, fs


def test_device_code_vs_live_reference_fuzz(emul, have_ref):
    """Garbage packets, losses and DTX through the product's device code (one emulated lane) and the reference's C build:
    samples and final range identical, PCM equal to rounding -- whatever the bytes are."""
    if not have_ref:
        pytest.skip("oracle/_ref not built")
    from oracle import refpy
    for pk, ln, fs in _fuzz_streams(2024, 40):
        nf = pk.shape[0]
        for dc in (1, 2):
            ref, rr, rs = refpy.decode_stream(pk, ln, fs, dc, pure_c=True)
            pcm = np.zeros((nf, fs * dc), np.float32); rng = np.zeros(nf, np.uint32); smp = np.zeros(nf, np.int32)
            emul.emul_decode_stream(P(pk, C.c_ubyte), P(ln, C.c_int), pk.shape[1], nf, fs, dc, P(pcm, C.c_float), P(rng, C.c_uint32), P(smp, C.c_int), None)
            assert (smp == rs).all() and (rng == rr).all()
            m = np.isfinite(ref)
            assert (m == np.isfinite(pcm)).all()
            scale = max(1.0, float(np.abs(ref[m]).max()))
            assert np.abs(pcm[m] - ref[m]).max() <= 1e-5 * scale


@pytest.mark.parametrize("ch,br,fs,nb,vbr,cx", ENC_CELT)
def test_encoder_device_code_bit_identical_to_reference_celt_encoder(emul, have_ref, ch, br, fs, nb, vbr, cx):
    """celt_encode_with_ec driven directly (no Opus-layer analysis): packets must be IDENTICAL, every byte of every frame."""
    if not have_ref:
        pytest.skip("oracle/_ref not built")
    from opus_codec_b200 import synth
    from oracle import refpy
    for s in range(3):
        pcm = synth.stream_pcm(s, 48000, ch)
        ref_pk, ref_ln, ref_rng = refpy.celt_encode_stream(pcm, fs, ch, br, nb, vbr=vbr, complexity=cx, pure_c=True)
        nf = pcm.size // (fs * ch)
        out = np.zeros((nf, 1275), np.uint8); lens = np.zeros(nf, np.int32); rng = np.zeros(nf, np.uint32)
        r = emul.emul_celt_encode_stream(P(np.ascontiguousarray(pcm), C.c_float), nf, fs, ch, br, vbr, cx, nb, P(out, C.c_ubyte), 1275, P(lens, C.c_int), P(rng, C.c_uint32))
        assert r == 0
        assert (rng == ref_rng).all() and same_packets(out, lens, ref_pk, ref_ln)


@pytest.mark.parametrize("ch,br,fs,vbr,cx", [(1, 64000, 960, 0, 5), (2, 96000, 960, 0, 6), (2, 96000, 960, 1, 5), (1, 24000, 480, 2, 6), (2, 24000, 960, 0, 4), (1, 12000, 960, 0, 6),
                                             (2, 96000, 960, 0, 10), (1, 64000, 960, 0, 10), (2, 96000, 960, 1, 10), (1, 24000, 480, 2, 9), (2, 64000, 240, 0, 8),
                                             (1, 48000, 120, 1, 7), (2, 24000, 960, 0, 10), (1, 12000, 960, 1, 10), (2, 510000, 960, 0, 10),
                                             # 40 / 60 / 80 / 120 ms packets (FrameSize::Ms40, Ms60, ...): 20 ms CELT frames + repacketizer
                                             (1, 64000, 1920, 0, 10), (2, 96000, 2880, 1, 10), (2, 64000, 1920, 0, 5), (1, 48000, 2880, 2, 6),
                                             (2, 128000, 3840, 0, 10), (1, 32000, 5760, 1, 9)])
def test_encoder_opus_layer_bit_identical_to_reference(emul, have_ref, ch, br, fs, vbr, cx):
    """opus_encode_float (RESTRICTED_LOWDELAY): TOC, byte budget, bandwidth / stereo decisions, dc_reject and -- at complexity >= 7 --
    the tonality analysis with its FFT, band statistics, bandwidth detector and GRU network: identical packets."""
    if not have_ref:
        pytest.skip("oracle/_ref not built")
    from opus_codec_b200 import synth
    from oracle import refpy
    L = refpy.lib_c()
    u8p, i32p, u32p, f32p = (C.POINTER(C.c_ubyte), C.POINTER(C.c_int), C.POINTER(C.c_uint32), C.POINTER(C.c_float))
    L.ref_encode_stream.argtypes = [f32p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, u8p, C.c_int, i32p, u32p]
    for s in range(3):
        pcm = np.ascontiguousarray(synth.stream_pcm(s, 48000, ch))
        nf = pcm.size // (fs * ch)
        a = np.zeros((nf, 1275), np.uint8); al = np.zeros(nf, np.int32); ar = np.zeros(nf, np.uint32)
        b = np.zeros((nf, 1275), np.uint8); bl = np.zeros(nf, np.int32); br_ = np.zeros(nf, np.uint32)
        assert L.ref_encode_stream(P(pcm, C.c_float), nf, fs, ch, 2051, br, vbr, cx, P(a, C.c_ubyte), 1275, P(al, C.c_int), P(ar, C.c_uint32)) == 0
        assert emul.emul_opus_encode_stream(P(pcm, C.c_float), nf, fs, ch, br, vbr, cx, P(b, C.c_ubyte), 1275, P(bl, C.c_int), P(br_, C.c_uint32)) == 0
        assert (ar == br_).all() and same_packets(a, al, b, bl)


@pytest.mark.parametrize("app,ch,br,fs,vbr,cx", [(2049, 2, 96000, 960, 0, 10), (2049, 1, 64000, 960, 1, 10), (2049, 2, 128000, 480, 0, 5), (2049, 1, 96000, 240, 2, 8),
                                                 (2049, 2, 96000, 2880, 0, 10), (2049, 1, 64000, 120, 0, 10), (2048, 2, 96000, 960, 0, 10), (2048, 2, 128000, 480, 0, 5),
                                                 (2048, 1, 96000, 240, 2, 8), (2048, 2, 96000, 1920, 1, 6), (2048, 1, 48000, 960, 0, 5),
                                                 # the reference leaves CELT at once, after one frame, or in mid-stream (frame 18 of stream 1)
                                                 (2049, 2, 48000, 960, 1, 10), (2049, 2, 32000, 480, 0, 7), (2048, 1, 64000, 960, 1, 10), (2049, 1, 24000, 960, 1, 5)])
def test_encoder_audio_and_voip_applications_bit_identical_while_celt_only(emul, have_ref, app, ch, br, fs, vbr, cx):
    """Application::Audio / ::Voip (src/types.rs): 4 ms delay compensation (delay_buffer), the VOIP high-pass (hp_cutoff), the stereo-width
    tracker and the SILK / CELT mode decision.  While the reference stays in MODE_CELT_ONLY the packets are identical; the first frame the
    reference gives to SILK / hybrid is OPUS_UNIMPLEMENTED here (and every frame before it is still identical)."""
    if not have_ref:
        pytest.skip("oracle/_ref not built")
    from opus_codec_b200 import synth
    from oracle import refpy
    L = refpy.lib_c()
    u8p, i32p, u32p, f32p = (C.POINTER(C.c_ubyte), C.POINTER(C.c_int), C.POINTER(C.c_uint32), C.POINTER(C.c_float))
    L.ref_encode_stream.argtypes = [f32p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, u8p, C.c_int, i32p, u32p]
    emul.emul_opus_encode_stream_app.argtypes = [f32p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, u8p, C.c_int, i32p, u32p]
    for s in range(4):
        pcm = np.ascontiguousarray(synth.stream_pcm(s, 48000, ch))
        nf = pcm.size // (fs * ch)
        a = np.zeros((nf, 1275), np.uint8); al = np.zeros(nf, np.int32); ar = np.zeros(nf, np.uint32)
        b = np.zeros((nf, 1275), np.uint8); bl = np.zeros(nf, np.int32); br_ = np.zeros(nf, np.uint32)
        L.ref_set_encoder_force_celt(0)                           # the encoder's own SILK / hybrid / CELT decision, as a user of the crate gets it
        try:
            assert L.ref_encode_stream(P(pcm, C.c_float), nf, fs, ch, app, br, vbr, cx, P(a, C.c_ubyte), 1275, P(al, C.c_int), P(ar, C.c_uint32)) == 0
        finally:
            L.ref_set_encoder_force_celt(1)
        rc = emul.emul_opus_encode_stream_app(P(pcm, C.c_float), nf, fs, ch, app, br, vbr, cx, P(b, C.c_ubyte), 1275, P(bl, C.c_int), P(br_, C.c_uint32))
        celt = (a[:, 0] & 0x80) != 0
        n = nf if celt.all() else int(np.argmin(celt))            # frames before the reference's first SILK / hybrid packet
        assert rc == (0 if n == nf else -5)
        assert (ar[:n] == br_[:n]).all() and same_packets(a[:n], al[:n], b[:n], bl[:n])
        if br >= 96000:
            assert n == nf                                        # these configurations never leave CELT


@pytest.mark.parametrize("app", [2051, 2049, 2048])
def test_encoder_too_small_budgets_emit_the_reference_plc_frames(emul, have_ref, app):
    """Budgets too small to code anything (opus_encoder.c:1202-1266): a TOC-only 'PLC frame' (code 0 / 1 / 3, padded in CBR), with the mode of
    the previous packet (MODE_HYBRID before the first one); 100 ms in one byte is OPUS_BUFFER_TOO_SMALL.  Final range 0."""
    if not have_ref:
        pytest.skip("oracle/_ref not built")
    from opus_codec_b200 import synth
    from oracle import refpy
    L = refpy.lib_c()
    u8p, i32p, u32p, f32p = (C.POINTER(C.c_ubyte), C.POINTER(C.c_int), C.POINTER(C.c_uint32), C.POINTER(C.c_float))
    L.ref_encode_stream.argtypes = [f32p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, u8p, C.c_int, i32p, u32p]
    emul.emul_opus_encode_stream_app.argtypes = [f32p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, u8p, C.c_int, i32p, u32p]
    for ch, br, fs, vbr, cx, mb in [(1, 6000, 120, 0, 10, 1275), (1, 1000, 960, 1, 5, 1275), (1, 2000, 2880, 0, 10, 1275), (2, 2000, 1920, 0, 5, 1275),
                                    (2, 2000, 1920, 1, 5, 1275), (1, 2000, 5760, 1, 9, 1275), (1, 64000, 960, 1, 9, 2), (2, 64000, 2880, 1, 9, 1),
                                    (1, 64000, 480, 0, 9, 2), (1, 500, 3840, 0, 5, 1275), (2, 96000, 4800, 1, 9, 1), (1, 8000, 240, 0, 5, 1275)]:
        pcm = np.ascontiguousarray(synth.stream_pcm(1, 48000, ch))
        nf = pcm.size // (fs * ch)
        a = np.zeros((nf, mb), np.uint8); al = np.zeros(nf, np.int32); ar = np.ones(nf, np.uint32)
        b = np.zeros((nf, mb), np.uint8); bl = np.zeros(nf, np.int32); br_ = np.ones(nf, np.uint32)
        L.ref_set_encoder_force_celt(0)
        try:
            r0 = L.ref_encode_stream(P(pcm, C.c_float), nf, fs, ch, app, br, vbr, cx, P(a, C.c_ubyte), mb, P(al, C.c_int), P(ar, C.c_uint32))
        finally:
            L.ref_set_encoder_force_celt(1)
        r1 = emul.emul_opus_encode_stream_app(P(pcm, C.c_float), nf, fs, ch, app, br, vbr, cx, P(b, C.c_ubyte), mb, P(bl, C.c_int), P(br_, C.c_uint32))
        assert r0 == r1 == (-2 if fs == 4800 else 0), (ch, br, fs, vbr, cx, mb)
        assert (ar == br_).all() and same_packets(a, al, b, bl), (ch, br, fs, vbr, cx, mb)


def _gappy_pcm(s, ch, n):
    """Programme material with a stretch of digital silence and a stretch of faint noise: what the generalised DTX reacts to."""
    from opus_codec_b200 import synth
    x = synth.stream_pcm(s, n, ch).reshape(-1, ch).copy()
    x[n // 4:n // 4 + n // 3] = 0
    x[3 * n // 4:3 * n // 4 + n // 5] = np.random.default_rng(s).normal(0, 2e-4, (n // 5, ch)).astype(np.float32)
    return np.ascontiguousarray(x.reshape(-1))


#                                  signal pred phase_inv dtx fec loss
@pytest.mark.parametrize("extras", [(0, 0, 0, 1, 0, 0), (3001, 0, 0, 0, 0, 0), (3002, 0, 0, 0, 0, 0), (0, 1, 0, 0, 0, 0), (0, 0, 1, 0, 0, 0), (0, 1, 1, 1, 0, 10),
                                    (0, 0, 0, 0, 1, 20), (3001, 0, 0, 1, 2, 30)])
def test_encoder_ctls_signal_prediction_phase_inversion_dtx_fec(emul, have_ref, extras):
    """OPUS_SET_SIGNAL / _PREDICTION_DISABLED / _PHASE_INVERSION_DISABLED / _DTX / _INBAND_FEC (+ loss): identical packets (DTX packets are the TOC
    byte alone, final range 0), identical OPUS_GET_IN_DTX, for the three applications; SILK / hybrid frames are OPUS_UNIMPLEMENTED as before."""
    if not have_ref:
        pytest.skip("oracle/_ref not built")
    from oracle import refpy
    L = refpy.lib_c()
    u8p, i32p, u32p, f32p = (C.POINTER(C.c_ubyte), C.POINTER(C.c_int), C.POINTER(C.c_uint32), C.POINTER(C.c_float))
    L.ref_encode_stream.argtypes = [f32p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, u8p, C.c_int, i32p, u32p]
    emul.emul_opus_encode_stream_app.argtypes = [f32p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, u8p, C.c_int, i32p, u32p]
    dtx_packets = 0
    L.ref_set_encoder_force_celt(0)
    L.ref_set_encoder_extras2(extras[0], extras[1], extras[2], extras[3], extras[4], 0, extras[5])
    emul.emul_set_encoder_extras(*extras)
    try:
        for app in (2051, 2049, 2048):
            for ch, br, fs, vbr, cx in [(2, 96000, 960, 0, 10), (1, 64000, 960, 1, 9), (2, 64000, 480, 0, 7), (1, 96000, 2880, 1, 10), (2, 128000, 240, 2, 5)]:
                pcm = _gappy_pcm(1, ch, 48000 * 2)
                nf = pcm.size // (fs * ch)
                a = np.zeros((nf, 1275), np.uint8); al = np.zeros(nf, np.int32); ar = np.ones(nf, np.uint32)
                b = np.zeros((nf, 1275), np.uint8); bl = np.zeros(nf, np.int32); br_ = np.ones(nf, np.uint32)
                assert L.ref_encode_stream(P(pcm, C.c_float), nf, fs, ch, app, br, vbr, cx, P(a, C.c_ubyte), 1275, P(al, C.c_int), P(ar, C.c_uint32)) == 0
                rc = emul.emul_opus_encode_stream_app(P(pcm, C.c_float), nf, fs, ch, app, br, vbr, cx, P(b, C.c_ubyte), 1275, P(bl, C.c_int), P(br_, C.c_uint32))
                celt = (a[:, 0] & 0x80) != 0
                n = nf if celt.all() else int(np.argmin(celt))
                assert rc == (0 if n == nf else -5), (app, ch, br, fs, vbr, cx)
                assert (ar[:n] == br_[:n]).all() and same_packets(a[:n], al[:n], b[:n], bl[:n]), (app, ch, br, fs, vbr, cx)
                if n == nf:
                    assert L.ref_last_in_dtx() == emul.emul_last_in_dtx()
                dtx_packets += int((bl[:n] == 1).sum())
    finally:
        L.ref_set_encoder_extras2(0, 0, 0, 0, 0, 0, 0)
        L.ref_set_encoder_force_celt(1)
        emul.emul_set_encoder_extras(0, 0, 0, 0, 0, 0)
    assert (dtx_packets > 20) == bool(extras[3])


def test_framing_pass_reserves_a_slot_for_every_packet(emul):
    """ob_frame_packets with as many frame slots as packets: a multi-frame packet that would eat the slots of later packets is
    OPUS_BUFFER_TOO_SMALL, and EVERY packet of the call gets at least one slot (so its samples / range are always written)."""
    from conftest import repacketize
    g = load_golden("cfg2_mono_20ms_64k_cbr")
    fs = g["frame_size"]
    pk0, ln0 = g["packets"][0], g["lens"][0]
    fr = [bytes(pk0[f, :ln0[f]]) for f in range(8)]
    pkts = [repacketize(fr[0:2], 1), repacketize(fr[2:4], 1), fr[4]]
    buf = np.frombuffer(b"".join(pkts), np.uint8).copy()
    lens = np.array([len(p) for p in pkts], np.int32)
    offs = np.concatenate([[0], np.cumsum(lens)[:-1]]).astype(np.int32)
    for cap, want in ((4, [(0, 0), (0, 0), (-2, 1), (0, 2)]), (3, [(-2, 0), (-2, 1), (0, 2)]), (5, [(0, 0), (0, 0), (0, 1), (0, 1), (0, 2)])):
        status = np.zeros(16, np.int32); pkt = np.zeros(16, np.int32)
        n = emul.emul_frame_packets(P(buf, C.c_ubyte), P(offs, C.c_int), P(lens, C.c_int), 3, 2 * fs, cap, P(status, C.c_int), P(pkt, C.c_int))
        assert [(int(status[i]), int(pkt[i])) for i in range(n)] == want, (cap, status[:n], pkt[:n])
        assert set(pkt[:n]) == {0, 1, 2}


@pytest.mark.parametrize("ch,br,fs,vbr,cx", [(2, 128000, 960, 0, 10), (2, 96000, 960, 0, 10), (1, 64000, 960, 1, 10), (2, 64000, 480, 2, 9), (1, 48000, 120, 1, 7),
                                             (2, 96000, 2880, 1, 10)])
def test_encoder_in_warp_summation_order_is_a_valid_encoder(emul, have_ref, ch, br, fs, vbr, cx):
    """The encoder source evaluated in the 32-lane warp's order (ObSoloW: strided partial sums + butterfly, chunk-per-lane scans, lane
    winners ranked by quotient) -- the packets the warp-per-stream GPU mapping must produce.  Not bit-identical to the reference's C build
    (nor is the reference's own SSE build), so the gates are north_star's: the REFERENCE decoder accepts every packet with the encoder's
    final range, and on BASELINE configs 1 / 3 the decoded audio passes the reference's opus_compare against the reference encoder's round trip."""
    if not have_ref:
        pytest.skip("oracle/_ref not built")
    from conftest import emul_warp_encode, opus_compare
    from opus_codec_b200 import synth
    from oracle import refpy
    L = refpy.lib_c()
    u8p, i32p, u32p, f32p = (C.POINTER(C.c_ubyte), C.POINTER(C.c_int), C.POINTER(C.c_uint32), C.POINTER(C.c_float))
    L.ref_encode_stream.argtypes = [f32p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, u8p, C.c_int, i32p, u32p]
    for s in range(3):
        pcm = np.ascontiguousarray(synth.stream_pcm(s, 48000 * 2, ch, base_seed=4242))
        out, lens, rng, rc = emul_warp_encode(pcm, fs, ch, br, vbr, cx)
        assert rc == 0 and (lens > 0).all()
        ours, dec_rng, smp = refpy.decode_stream(out, lens, fs, ch)
        assert (smp == fs).all() and (dec_rng == rng).all()
        if fs == 960 and ch == 2 and vbr == 0:
            nf = pcm.size // (fs * ch)
            a = np.zeros((nf, 1276), np.uint8); al = np.zeros(nf, np.int32); ar = np.zeros(nf, np.uint32)
            assert L.ref_encode_stream(P(pcm, C.c_float), nf, fs, ch, 2051, br, vbr, cx, P(a, C.c_ubyte), 1276, P(al, C.c_int), P(ar, C.c_uint32)) == 0
            theirs, _, _ = refpy.decode_stream(a, al, fs, ch)
            ok, err, text = opus_compare(theirs, ours, ch)
            assert ok, (s, text)


class _IrHdr(C.Structure):
    _fields_ = [("status", C.c_int32), ("final_range", C.c_uint32), ("n_leaves", C.c_uint16), ("pf_pitch", C.c_uint16),
                ("LM", C.c_uint8), ("C", C.c_uint8), ("end", C.c_uint8), ("flags", C.c_uint8),
                ("spread", C.c_uint8), ("pf_tapset", C.c_uint8), ("pf_qg", C.c_uint8), ("coded_bands", C.c_uint8),
                ("intensity", C.c_uint8), ("dual_stereo", C.c_uint8), ("skip_in", C.c_uint8), ("end_in", C.c_uint8),
                ("lcg_total", C.c_uint32), ("seed_in", C.c_uint32), ("loss_in", C.c_int32), ("lastfs_in", C.c_uint16), ("pad2", C.c_uint16),
                ("coarse_qi", C.c_int16 * 42), ("pulses", C.c_int16 * 21), ("fine_quant", C.c_uint8 * 21), ("fine_q2", C.c_uint8 * 42),
                ("final_bit", C.c_int8 * 42), ("collapse_masks", C.c_uint8 * 42), ("tf_change", C.c_int8 * 21), ("pad1", C.c_uint8 * 3)]


@pytest.mark.parametrize("name", golden_names())
def test_symbol_pass_integer_ir_matches_oracle_taps(emul, name):
    """The symbol kernel's device code (one emulated thread per frame, stateless) against the oracle's taps: energy indices, allocation,
    tf, masks, flags and every pulse vector, exactly.  (The same comparison runs on the GPU in tests/test_gpu_decode.py.)"""
    assert C.sizeof(_IrHdr) == emul.emul_hdr_size()
    g = load_golden(name)
    fs, dc = g["frame_size"], g["dec_channels"]
    buf = C.create_string_buffer(emul.emul_ir_size())
    iy_off = emul.emul_ir_size() - 2 * 1920
    for s in range(min(2, g["packets"].shape[0])):
        nf = min(40, g["packets"].shape[1])
        pk = np.ascontiguousarray(g["packets"][s, :nf]); ln = np.ascontiguousarray(g["lens"][s, :nf])
        _, orng, _, taps = oraclepy.decode_stream(pk, ln, fs, dc, want_taps=True)
        for f in range(nf):
            row = np.ascontiguousarray(pk[f])
            emul.emul_decode_symbols(P(row, C.c_ubyte), int(ln[f]), dc, 960, buf)
            h = _IrHdr.from_buffer_copy(buf.raw[:C.sizeof(_IrHdr)]); t = taps[f]
            C_, end = h.C, h.end
            assert h.status == fs and h.final_range == int(orng[f]) and (h.LM, C_, end) == (t.LM, t.C, t.end)
            assert [h.coarse_qi[c * 21 + i] for c in range(C_) for i in range(end)] == [t.coarse_qi[c * 21 + i] for c in range(C_) for i in range(end)]
            assert list(h.pulses)[:end] == list(t.pulses)[:end] and list(h.fine_quant)[:end] == list(t.fine_quant)[:end]
            assert list(h.tf_change)[:end] == list(t.tf_res)[:end] and list(h.collapse_masks)[:C_ * end] == list(t.collapse_masks)[:C_ * end]
            assert (h.spread, h.intensity, h.dual_stereo, h.coded_bands) == (t.spread, t.intensity, t.dual_stereo, t.coded_bands)
            iy = np.frombuffer(buf.raw, np.int16, 1920, iy_off)
            want = np.frombuffer(bytes(t.iy), np.int16); mask = np.frombuffer(bytes(t.iy_set), np.uint8) != 0
            assert np.array_equal(iy[mask], want[mask]), (name, s, f)


def test_rectangular_pvq_table_equals_the_reference_triangle():
    """OB_PVQ_U_RECT (one load per U(n,k)) holds exactly the 1272 numbers of CELT_PVQ_U_DATA behind CELT_PVQ_U_ROW (opus/celt/cwrs.c:213-428)."""
    from conftest import emul_lib
    assert emul_lib().emul_pvq_table_check() == 0

