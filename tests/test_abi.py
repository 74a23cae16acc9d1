"""CPU tests of the drop-in boundary: the C-ABI library loads, exports every symbol include/opus_b200.h declares,
refuses to run without a GPU (no CPU fallback), and its TOC helpers agree with the reference's packet_* semantics."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import ROOT, load_golden


def _lib():
    from opus_codec_b200 import _lib
    _lib.build()
    return _lib.lib(), _lib


def test_library_exports_every_declared_symbol():
    L, mod = _lib()
    hdr = open(os.path.join(ROOT, "include", "opus_b200.h")).read()
    declared = set(re.findall(r"\b(ob_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    assert declared == set(mod.SYMBOLS)
    for name in declared:
        assert hasattr(L, name), name
    assert L.ob_version().decode().startswith("1.5.2-b200")
    assert L.ob_strerror(-4) == b"corrupted stream" and L.ob_strerror(0) == b"success"


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from opus_codec_b200.batch import BatchDecoder, OpusError, INTERNAL_ERROR
    with pytest.raises(OpusError) as e:
        BatchDecoder(4, 48000, 1)
    assert e.value.code == INTERNAL_ERROR


def test_create_argument_validation():
    L, _ = _lib()
    err = C.c_int32(0)
    for args, code in (((0, 48000, 1, 0, 1), -1), ((4, 48000, 3, 0, 1), -1), ((4, 48000, 1, 0, 0), -1), ((4, 44100, 1, 0, 1), -1)):
        assert not L.ob_decoder_create(*args, C.byref(err))
        assert err.value == code


def test_toc_helpers_match_reference_semantics():
    """All 64 TOC configs x stereo bit x frame-count codes (opus/tests/test_opus_decode.c walks the same space)."""
    L, _ = _lib()
    for toc in range(256):
        p = np.array([toc, 3, 0, 0], np.uint8)
        ptr = C.c_void_p(p.ctypes.data)
        assert L.ob_packet_get_nb_channels(ptr) == (2 if toc & 4 else 1)
        cfg = toc >> 3
        if cfg >= 16:
            spf = 120 << (cfg & 3)
            bw = [1101, 1103, 1104, 1105][(cfg >> 2) & 3]
        elif cfg >= 12:
            spf = 960 if cfg & 1 else 480
            bw = 1105 if cfg >= 14 else 1104
        else:
            spf = [480, 960, 1920, 2880][cfg & 3]
            bw = [1101, 1102, 1103][cfg >> 2]
        assert L.ob_packet_get_samples_per_frame(ptr, 48000) == spf, toc
        assert L.ob_packet_get_bandwidth(ptr) == bw, toc
        assert L.ob_packet_get_nb_frames(ptr, 4) == [1, 2, 2, 3][toc & 3]
    g = load_golden("cfg2_mono_20ms_64k_cbr")
    p = np.ascontiguousarray(g["packets"][0, 0])
    assert p[0] == 0xF8 and L.ob_packet_get_samples_per_frame(C.c_void_p(p.ctypes.data), 48000) == 960


def test_pack_packets_layout():
    from opus_codec_b200.batch import pack_packets
    buf, off, ln = pack_packets([[b"ab", b"cde"], [b"", b"f"]])
    assert ln.tolist() == [[2, 3], [0, 1]] and off.tolist() == [[0, 2], [5, 5]] and bytes(buf) == b"abcdef"
