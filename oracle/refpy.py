"""ctypes front-end to oracle/_ref/libopus_ref.so (the UNMODIFIED reference + oracle/ref_shim.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs.  Never imported by the product package.
"""
import ctypes as C
import os
import subprocess
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SO = os.path.join(HERE, "_ref", "libopus_ref.so")
REF_C_SO = os.path.join(HERE, "_ref", "libopus_ref_c.so")     # same sources, no x86 intrinsics (bit-reproducible float paths)
OPUS_COMPARE = os.path.join(HERE, "_ref", "opus_compare")

APP_VOIP, APP_AUDIO, APP_LOWDELAY = 2048, 2049, 2051
CBR, VBR, CVBR = 0, 1, 2

_lib = None
_lib_c = None


def build(quiet=True):
    """(Re)build oracle/_ref when the reference sources are present; otherwise keep the prebuilt files."""
    subprocess.run(["make", "-C", HERE, "ref", "-j8"] + (["-s"] if quiet else []), check=True)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(REF_SO):
            build()
        L = C.CDLL(REF_SO)
        u8p, i32p, u32p, f32p = (C.POINTER(C.c_ubyte), C.POINTER(C.c_int), C.POINTER(C.c_uint32), C.POINTER(C.c_float))
        L.ref_encode_stream.argtypes = [f32p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, u8p, C.c_int, i32p, u32p]
        L.ref_encode_stream.restype = C.c_int
        L.ref_decode_stream.argtypes = [u8p, i32p, C.c_int, C.c_int, C.c_int, C.c_int, f32p, u32p, i32p, C.c_void_p]
        L.ref_decode_stream.restype = C.c_int
        L.ref_decode_pool.argtypes = [u8p, i32p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, f32p, u32p]
        L.ref_decode_pool.restype = C.c_double
        L.ref_encode_pool.argtypes = [f32p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, u8p, C.c_int, i32p, u32p]
        L.ref_encode_pool.restype = C.c_double
        L.ref_tap_size.restype = C.c_int
        L.ref_version.restype = C.c_char_p
        _lib = L
    return _lib


def _declare_celt(L):
    u8p, i32p, u32p, f32p = (C.POINTER(C.c_ubyte), C.POINTER(C.c_int), C.POINTER(C.c_uint32), C.POINTER(C.c_float))
    L.ref_celt_encode_stream.argtypes = [f32p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, u8p, C.c_int, i32p, u32p]
    L.ref_celt_encode_stream.restype = C.c_int
    L.ref_celt_decode_stream.argtypes = [u8p, i32p, C.c_int, C.c_int, C.c_int, C.c_int, f32p, u32p]
    L.ref_celt_decode_stream.restype = C.c_int
    return L


def lib_c():
    """The no-intrinsics build of the reference (pure C float paths)."""
    global _lib_c
    if _lib_c is None:
        if not os.path.exists(REF_C_SO):
            build()
        _lib_c = _declare_celt(C.CDLL(REF_C_SO))
    return _lib_c


def celt_encode_stream(pcm, frame_size, channels, bitrate, nbytes, vbr=CBR, complexity=10, max_bytes=1275, pure_c=True):
    """CELT-level encode (no TOC, no Opus-layer analysis) -> (packets [nframes,max_bytes], lens, ranges)."""
    L = lib_c() if pure_c else _declare_celt(lib())
    pcm = np.ascontiguousarray(pcm, np.float32)
    nframes = pcm.size // (frame_size * channels)
    out = np.zeros((nframes, max_bytes), np.uint8)
    lens = np.zeros(nframes, np.int32)
    rng = np.zeros(nframes, np.uint32)
    r = L.ref_celt_encode_stream(_p(pcm, C.c_float), nframes, frame_size, channels, bitrate, vbr, complexity, nbytes,
                                 _p(out, C.c_ubyte), max_bytes, _p(lens, C.c_int), _p(rng, C.c_uint32))
    if r != 0:
        raise RuntimeError("ref_celt_encode_stream: opus error %d" % r)
    return out, lens, rng


def celt_decode_stream(pkts, lens, frame_size, channels, pure_c=False):
    L = lib_c() if pure_c else _declare_celt(lib())
    pkts = np.ascontiguousarray(pkts, np.uint8); lens = np.ascontiguousarray(lens, np.int32)
    nframes, stride = pkts.shape
    pcm = np.zeros((nframes, frame_size * channels), np.float32)
    rng = np.zeros(nframes, np.uint32)
    r = L.ref_celt_decode_stream(_p(pkts, C.c_ubyte), _p(lens, C.c_int), stride, nframes, frame_size, channels, _p(pcm, C.c_float), _p(rng, C.c_uint32))
    if r != 0:
        raise RuntimeError("ref_celt_decode_stream: opus error %d" % r)
    return pcm, rng


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t)) if a is not None else None


class Tap(C.Structure):
    _fields_ = [
        ("n_qab", C.c_int), ("LM", C.c_int), ("C", C.c_int), ("shortBlocks", C.c_int), ("spread", C.c_int),
        ("dual_stereo", C.c_int), ("intensity", C.c_int), ("codedBands", C.c_int),
        ("total_bits", C.c_int), ("balance", C.c_int),
        ("pulses", C.c_int * 21), ("tf_res", C.c_int * 21),
        ("collapse_masks", C.c_uint8 * 42), ("seed_in", C.c_uint32), ("seed_out", C.c_uint32),
        ("X", C.c_float * 1920),
        ("n_denorm", C.c_int), ("bandLogE", C.c_float * 42), ("freq", C.c_float * 1920),
        ("n_comb", C.c_int), ("presyn", C.c_float * 2160),
        ("comb_T0", C.c_int * 4), ("comb_T1", C.c_int * 4), ("comb_N", C.c_int * 4),
        ("comb_g0", C.c_float * 4), ("comb_g1", C.c_float * 4),
    ]


def encode_stream(pcm, frame_size, channels, bitrate, vbr=CBR, complexity=10, application=APP_LOWDELAY, max_bytes=1275,
                  bandwidth=0, force_channels=0):
    """pcm float32 [nframes*frame_size*channels] -> (packets u8 [nframes,max_bytes], lens i32, ranges u32).
    bandwidth: 0 or OPUS_BANDWIDTH_* 1101..1105; force_channels: 0, 1 or 2."""
    lib().ref_set_encoder_extras(bandwidth, force_channels)
    try:
        return _encode_stream(pcm, frame_size, channels, bitrate, vbr, complexity, application, max_bytes)
    finally:
        lib().ref_set_encoder_extras(0, 0)


def _encode_stream(pcm, frame_size, channels, bitrate, vbr, complexity, application, max_bytes):
    pcm = np.ascontiguousarray(pcm, np.float32)
    nframes = pcm.size // (frame_size * channels)
    out = np.zeros((nframes, max_bytes), np.uint8)
    lens = np.zeros(nframes, np.int32)
    rng = np.zeros(nframes, np.uint32)
    r = lib().ref_encode_stream(_p(pcm, C.c_float), nframes, frame_size, channels, application, bitrate, vbr,
                                complexity, _p(out, C.c_ubyte), max_bytes, _p(lens, C.c_int), _p(rng, C.c_uint32))
    if r != 0:
        raise RuntimeError("ref_encode_stream: opus error %d" % r)
    return out, lens, rng


def decode_stream(pkts, lens, frame_size, channels, want_taps=False, pure_c=False, gain_q8=0, phase_inv_disabled=False, fs=48000, fec_flags=None):
    """pkts u8 [nframes, stride] -> (pcm f32 [nframes, frame_size*channels], ranges u32, samples i32[, taps])."""
    pkts = np.ascontiguousarray(pkts, np.uint8)
    lens = np.ascontiguousarray(lens, np.int32)
    nframes, stride = pkts.shape
    pcm = np.zeros((nframes, frame_size * channels), np.float32)
    rng = np.zeros(nframes, np.uint32)
    smp = np.zeros(nframes, np.int32)
    taps = (Tap * nframes)() if want_taps else None
    assert not want_taps or C.sizeof(Tap) == lib().ref_tap_size()
    L = lib_c() if pure_c else lib()
    L.ref_set_decoder_extras(int(gain_q8), 1 if phase_inv_disabled else 0)
    L.ref_set_decoder_fs(int(fs))
    fec = np.ascontiguousarray(fec_flags, np.int32) if fec_flags is not None else None
    L.ref_set_decoder_fec_flags(_p(fec, C.c_int) if fec is not None else None)
    try:
        r = L.ref_decode_stream(_p(pkts, C.c_ubyte), _p(lens, C.c_int), stride, nframes, frame_size, channels,
                                _p(pcm, C.c_float), _p(rng, C.c_uint32), _p(smp, C.c_int),
                                C.cast(taps, C.c_void_p) if want_taps else None)
    finally:
        L.ref_set_decoder_extras(0, 0)
        L.ref_set_decoder_fs(48000)
        L.ref_set_decoder_fec_flags(None)
    if r != 0:
        raise RuntimeError("ref_decode_stream: opus error %d" % r)
    return (pcm, rng, smp, taps) if want_taps else (pcm, rng, smp)


def decode_stream_i16(pkts, lens, frame_size, channels, gain_q8=0, pure_c=False):
    """opus_decode (int16 API): (pcm i16 [nframes, frame_size*channels], ranges, samples)."""
    pkts = np.ascontiguousarray(pkts, np.uint8)
    lens = np.ascontiguousarray(lens, np.int32)
    nframes, stride = pkts.shape
    pcm = np.zeros((nframes, frame_size * channels), np.int16)
    rng = np.zeros(nframes, np.uint32)
    smp = np.zeros(nframes, np.int32)
    L = lib_c() if pure_c else lib()
    L.ref_set_decoder_extras(int(gain_q8), 0)
    try:
        r = L.ref_decode_stream_i16(_p(pkts, C.c_ubyte), _p(lens, C.c_int), stride, nframes, frame_size, channels,
                                    _p(pcm, C.c_int16), _p(rng, C.c_uint32), _p(smp, C.c_int))
    finally:
        L.ref_set_decoder_extras(0, 0)
    if r != 0:
        raise RuntimeError("ref_decode_stream_i16: opus error %d" % r)
    return pcm, rng, smp


def soft_clip(x, channels, mem):
    """opus_pcm_soft_clip on interleaved float32 x (in place); mem: float32[channels] carried state."""
    lib().opus_pcm_soft_clip(_p(x, C.c_float), x.size // channels, channels, _p(mem, C.c_float))


def decode_pool(pkts, lens, frame_size, channels, nthreads, want_pcm=False, want_ranges=False):
    """pkts u8 [nstreams, nframes, stride]; returns (seconds, pcm|None, ranges|None)."""
    pkts = np.ascontiguousarray(pkts, np.uint8)
    lens = np.ascontiguousarray(lens, np.int32)
    ns, nf, stride = pkts.shape
    pcm = np.zeros((ns, nf, frame_size * channels), np.float32) if want_pcm else None
    rng = np.zeros((ns, nf), np.uint32) if want_ranges else None
    t = lib().ref_decode_pool(_p(pkts, C.c_ubyte), _p(lens, C.c_int), stride, ns, nf, frame_size, channels, nthreads,
                              _p(pcm, C.c_float), _p(rng, C.c_uint32))
    if t < 0:
        raise RuntimeError("ref_decode_pool: opus error %d" % int(t))
    return t, pcm, rng


def encode_pool(pcm, frame_size, channels, bitrate, vbr, complexity, nthreads, application=APP_LOWDELAY, stride=1275, want_packets=True):
    """pcm f32 [nstreams, nframes*frame_size*channels]; returns (seconds, pkts, lens, ranges)."""
    pcm = np.ascontiguousarray(pcm, np.float32)
    ns = pcm.shape[0]
    nf = pcm.shape[1] // (frame_size * channels)
    out = np.zeros((ns, nf, stride), np.uint8) if want_packets else None
    lens = np.zeros((ns, nf), np.int32)
    rng = np.zeros((ns, nf), np.uint32)
    t = lib().ref_encode_pool(_p(pcm, C.c_float), ns, nf, frame_size, channels, application, bitrate, vbr, complexity,
                              nthreads, _p(out, C.c_ubyte), stride, _p(lens, C.c_int), _p(rng, C.c_uint32))
    if t < 0:
        raise RuntimeError("ref_encode_pool: opus error %d" % int(t))
    return t, out, lens, rng


def opus_compare(ref_pcm, test_pcm, channels, tmpdir):
    """Runs the reference's own opus_compare (opus/src/opus_compare.c) on two float PCM arrays (converted to s16).
    Returns (passed, quality_line)."""
    def to_s16(x):
        return np.clip(np.round(np.asarray(x, np.float64) * 32768.0), -32768, 32767).astype("<i2")
    a = os.path.join(tmpdir, "ref.sw")
    b = os.path.join(tmpdir, "tst.sw")
    r16 = to_s16(ref_pcm).reshape(-1)
    if channels == 1:            # opus_compare always reads file1 as stereo (opus_compare.c:231) and downmixes it
        r16 = np.repeat(r16, 2)
    r16.tofile(a)
    to_s16(test_pcm).tofile(b)
    args = [OPUS_COMPARE] + (["-s"] if channels == 2 else []) + ["-r", "48000", a, b]
    pr = subprocess.run(args, capture_output=True, text=True)
    txt = (pr.stdout + pr.stderr).strip()
    return pr.returncode == 0, txt
