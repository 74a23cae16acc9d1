"""ctypes front-end to oracle/_build/libcelt_oracle.so (our CPU restatement, oracle/celt_oracle.c).

TEST INFRASTRUCTURE ONLY -- see oracle/celt_oracle.h.
"""
import ctypes as C
import os
import subprocess
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "_build", "libcelt_oracle.so")
_lib = None


class Tap(C.Structure):
    _fields_ = [
        ("LM", C.c_int), ("C", C.c_int), ("end", C.c_int), ("silence", C.c_int), ("transient", C.c_int),
        ("intra", C.c_int), ("spread", C.c_int), ("trim", C.c_int), ("coded_bands", C.c_int),
        ("intensity", C.c_int), ("dual_stereo", C.c_int),
        ("pf_on", C.c_int), ("pf_pitch", C.c_int), ("pf_tapset", C.c_int), ("pf_qg", C.c_int),
        ("anti_collapse_on", C.c_int), ("anti_collapse_rsv", C.c_int),
        ("total_bits_q3", C.c_int), ("balance", C.c_int),
        ("tf_res", C.c_int * 21), ("pulses", C.c_int * 21), ("fine_quant", C.c_int * 21),
        ("fine_priority", C.c_int * 21), ("offsets", C.c_int * 21),
        ("coarse_qi", C.c_int * 42), ("collapse_masks", C.c_uint8 * 42),
        ("seed_in", C.c_uint32), ("seed_out", C.c_uint32), ("final_range", C.c_uint32),
        ("X", C.c_float * 1920), ("bandLogE", C.c_float * 42), ("freq", C.c_float * 1920),
        ("presyn", C.c_float * 2160), ("iy", C.c_int16 * 1920), ("iy_set", C.c_uint8 * 1920),
    ]


def build(quiet=True):
    subprocess.run(["make", "-C", HERE, "oracle"] + (["-s"] if quiet else []), check=True)


def lib():
    global _lib
    if _lib is None:
        src = os.path.join(HERE, "celt_oracle.c")
        if not os.path.exists(SO) or os.path.getmtime(SO) < os.path.getmtime(src):
            build()
        L = C.CDLL(SO)
        u8p, i32p, u32p, f32p = (C.POINTER(C.c_ubyte), C.POINTER(C.c_int), C.POINTER(C.c_uint32), C.POINTER(C.c_float))
        L.co_decode_stream.argtypes = [u8p, i32p, C.c_int, C.c_int, C.c_int, C.c_int, f32p, u32p, i32p, C.c_void_p]
        L.co_decode_stream.restype = C.c_int
        L.co_tap_size.restype = C.c_int
        L.co_bitexact_cos.argtypes = [C.c_int]; L.co_bitexact_cos.restype = C.c_int
        L.co_bitexact_log2tan.argtypes = [C.c_int, C.c_int]; L.co_bitexact_log2tan.restype = C.c_int
        L.co_isqrt32.argtypes = [C.c_uint32]; L.co_isqrt32.restype = C.c_uint
        L.co_pvq_v.argtypes = [C.c_int, C.c_int]; L.co_pvq_v.restype = C.c_uint32
        L.co_cwrsi.argtypes = [C.c_int, C.c_int, C.c_uint32, i32p]; L.co_cwrsi.restype = C.c_uint32
        L.co_icwrs.argtypes = [C.c_int, i32p]; L.co_icwrs.restype = C.c_uint32
        L.co_mdct_backward.argtypes = [f32p, f32p, C.c_int, C.c_int]
        L.co_fft.argtypes = [f32p, C.c_int]
        assert L.co_tap_size() == C.sizeof(Tap)
        _lib = L
    return _lib


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t)) if a is not None else None


def decode_stream(pkts, lens, frame_size, channels, want_taps=False):
    """pkts u8 [nframes, stride] (TOC included) -> (pcm [nframes, frame_size*channels], ranges, samples[, taps])."""
    pkts = np.ascontiguousarray(pkts, np.uint8)
    lens = np.ascontiguousarray(lens, np.int32)
    nframes, stride = pkts.shape
    pcm = np.zeros((nframes, frame_size * channels), np.float32)
    rng = np.zeros(nframes, np.uint32)
    smp = np.zeros(nframes, np.int32)
    taps = (Tap * nframes)() if want_taps else None
    r = lib().co_decode_stream(_p(pkts, C.c_ubyte), _p(lens, C.c_int), stride, nframes, frame_size, channels,
                               _p(pcm, C.c_float), _p(rng, C.c_uint32), _p(smp, C.c_int),
                               C.cast(taps, C.c_void_p) if want_taps else None)
    if r != 0:
        raise RuntimeError("co_decode_stream: error %d" % r)
    return (pcm, rng, smp, taps) if want_taps else (pcm, rng, smp)
