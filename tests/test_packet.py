"""Packet byte work through the C ABI against the reference library's own functions (oracle/_ref): opus_packet_parse,
opus_repacketizer_*, opus_packet_pad / _unpad, with padding extensions.  Results must be identical: counts, sizes, offsets, error
codes and every output byte.  The host entry points need no GPU; the batched kernel is tested in test_gpu_decode.py."""
import ctypes as C

import numpy as np
import pytest

from conftest import golden_names, load_golden, multiframe_stream, repacketize


@pytest.fixture(scope="module")
def ref(have_ref):
    if not have_ref:
        pytest.skip("oracle/_ref not built")
    from oracle import refpy
    L = refpy.lib_c()
    vp, i32 = C.c_void_p, C.c_int32
    L.opus_packet_parse.argtypes = [vp, i32, vp, vp, vp, vp]; L.opus_packet_parse.restype = C.c_int
    L.opus_repacketizer_create.argtypes = []; L.opus_repacketizer_create.restype = vp
    L.opus_repacketizer_destroy.argtypes = [vp]; L.opus_repacketizer_destroy.restype = None
    L.opus_repacketizer_init.argtypes = [vp]; L.opus_repacketizer_init.restype = vp
    L.opus_repacketizer_cat.argtypes = [vp, vp, i32]; L.opus_repacketizer_cat.restype = C.c_int
    L.opus_repacketizer_get_nb_frames.argtypes = [vp]; L.opus_repacketizer_get_nb_frames.restype = C.c_int
    L.opus_repacketizer_out_range.argtypes = [vp, C.c_int, C.c_int, vp, i32]; L.opus_repacketizer_out_range.restype = i32
    L.opus_repacketizer_out.argtypes = [vp, vp, i32]; L.opus_repacketizer_out.restype = i32
    L.opus_packet_pad.argtypes = [vp, i32, i32]; L.opus_packet_pad.restype = C.c_int
    L.opus_packet_unpad.argtypes = [vp, i32]; L.opus_packet_unpad.restype = i32
    return L


def _buf(b, n=None):
    n = max(1, len(b) if n is None else n)
    a = (C.c_uint8 * n)()
    C.memmove(a, bytes(b), len(b))
    return a


def ref_parse(L, pkt):
    d = _buf(pkt)
    toc = C.c_uint8(0); po = C.c_int(0)
    frames = (C.c_void_p * 48)(); sizes = (C.c_int16 * 48)()
    n = L.opus_packet_parse(d, len(pkt), C.byref(toc), frames, sizes, C.byref(po))
    if n < 0:
        return n, 0, 0, [], []
    base = C.addressof(d)
    return n, toc.value, po.value, [frames[i] - base for i in range(n)], list(sizes[:n])


def ref_merge(L, pkts, begin=None, end=None, maxlen=8000):
    """-> bytes, or the error code of the first failing cat / of out."""
    rp = L.opus_repacketizer_create()
    keep = []
    try:
        for p in pkts:
            b = _buf(p); keep.append(b)
            r = L.opus_repacketizer_cat(rp, b, len(p))
            if r != 0:
                return r
        out = (C.c_uint8 * maxlen)()
        n = L.opus_repacketizer_out(rp, out, maxlen) if begin is None else L.opus_repacketizer_out_range(rp, begin, end, out, maxlen)
        return bytes(out[:n]) if n >= 0 else n
    finally:
        L.opus_repacketizer_destroy(rp)


def our_merge(pkts, begin=None, end=None, maxlen=8000):
    from opus_codec_b200.packet import Repacketizer
    from opus_codec_b200.batch import OpusError
    with Repacketizer() as rp:
        try:
            for p in pkts:
                rp.push(p)
            return rp.out(maxlen) if begin is None else rp.out_range(begin, end, maxlen)
        except OpusError as e:
            return e.code


def sample_packets():
    """Code 0-3 packets (padding, DTX frame inside) of every golden configuration."""
    out = []
    for name in golden_names():
        g = load_golden(name)
        pk, ln = g["packets"][0], g["lens"][0]
        out += [bytes(pk[f, :ln[f]]) for f in range(min(6, pk.shape[0]))]
        if pk.shape[0] >= 24:
            out += [bytes(p) for p in multiframe_stream(g, 0, 7)]
    return out


def ext_padding():
    """Padding bytes that carry extensions (opus/src/extensions.c): short ids with 0/1 byte, a frame separator, a long extension with an
    explicit length, 0x01 filler, and a last long extension without length."""
    return bytes([0x01, 0x01, (2 << 1) | 1, ord("a"), 0x02, (33 << 1) | 1, 5]) + b"HELLO" + bytes([(5 << 1) | 0, 0x03, 2, (40 << 1) | 0]) + b"tail-bytes"


def with_padding(frames, padding):
    """A code-3 VBR packet whose padding is `padding` (may hold extensions)."""
    toc = frames[0][0] & 0xFC
    body = [bytes(f[1:]) for f in frames]
    out = bytes([toc | 3, 0x80 | 0x40 | len(body)])
    p = len(padding)
    while p > 254:
        out += bytes([255]); p -= 254
    out += bytes([p])
    for b in body[:-1]:
        n = len(b)
        out += bytes([n]) if n < 252 else bytes([252 + (n & 3), (n - (252 + (n & 3))) >> 2])
    return out + b"".join(body) + padding


def test_packet_parse_matches_reference(ref):
    from opus_codec_b200.packet import packet_parse_raw
    rng = np.random.default_rng(5)
    pkts = sample_packets()
    assert len(pkts) > 40
    for p in list(pkts):                                      # truncations and bit flips of valid packets
        for _ in range(6):
            q = bytearray(p[:int(rng.integers(1, len(p) + 1))])
            for _ in range(int(rng.integers(0, 3))):
                q[int(rng.integers(0, min(len(q), 6)))] ^= 1 << int(rng.integers(0, 8))
            pkts.append(bytes(q))
    pkts += [bytes(rng.integers(0, 256, int(rng.integers(1, 40)), dtype=np.uint8)) for _ in range(3000)]
    pkts += [b"\xfb" + bytes([0x80 | 0x40 | 3]) + b"\xff\xff\x10" + bytes(700), b"\xf8" * 1277, b"\xf8" + bytes(1275), b"\xfb\x00", b"\xfb\x31" + bytes(49)]
    seen = set()
    for p in pkts:
        a = ref_parse(ref, p)
        b = packet_parse_raw(p)
        assert a[0] == b[0], (p[:8].hex(), a[0], b[0])
        seen.add(a[0] if a[0] < 0 else "ok%d" % (p[0] & 3))
        if a[0] > 0:
            assert a == b, p[:8].hex()
    assert {"ok0", "ok1", "ok2", "ok3", -4} <= seen


def test_packet_parse_wrapper_follows_the_crate():
    from opus_codec_b200.packet import packet_parse
    from opus_codec_b200.batch import OpusError
    g = load_golden(golden_names()[0])
    fr = [bytes(g["packets"][0][f, :g["lens"][0][f]]) for f in range(3)]
    toc, po, frames = packet_parse(repacketize(fr, 3, pad=9))
    assert toc == ((fr[0][0] & 0xFC) | 3) and frames == [f[1:] for f in fr] and po > 2
    with pytest.raises(OpusError) as e:
        packet_parse(b"")
    assert e.value.code == -1
    with pytest.raises(OpusError) as e:
        packet_parse(b"\xf9\x00\x00\x00")                      # code 1 with an odd payload
    assert e.value.code == -4


def test_repacketizer_matches_reference(ref):
    rng = np.random.default_rng(8)
    checked = 0
    for name in golden_names():
        g = load_golden(name)
        pk, ln = g["packets"][0], g["lens"][0]
        fr = [bytes(pk[f, :ln[f]]) for f in range(pk.shape[0])]
        per120 = 5760 // g["frame_size"]
        for trial in range(12):
            m = int(rng.integers(1, min(per120, 8, len(fr)) + 1))
            k0 = int(rng.integers(0, len(fr) - m + 1))
            grp = fr[k0:k0 + m]
            if trial % 3 == 1 and m >= 2:                     # feed multi-frame packets (and padding) back in
                grp = [repacketize(grp[:2], 3, pad=int(rng.integers(0, 400)))] + grp[2:]
            if trial % 4 == 2:
                grp[-1] = grp[-1][:1]                          # a DTX frame (TOC only)
            a, b = ref_merge(ref, grp), our_merge(grp)
            assert a == b, (name, trial, m)
            checked += 1
            if isinstance(a, bytes) and m >= 2:
                b0, e0 = sorted(rng.choice(m + 1, 2, replace=False).tolist())
                assert ref_merge(ref, grp, b0, e0) == our_merge(grp, b0, e0)
                assert ref_merge(ref, grp, maxlen=len(a) - 1) == our_merge(grp, maxlen=len(a) - 1) == -2      # OPUS_BUFFER_TOO_SMALL
    assert checked > 50
    # error paths: mismatched configuration, more than 120 ms, bad ranges
    g1, g2 = load_golden(golden_names()[0]), load_golden(golden_names()[-1])
    p1, p2 = bytes(g1["packets"][0][0, :g1["lens"][0][0]]), bytes(g2["packets"][0][0, :g2["lens"][0][0]])
    if (p1[0] & 0xFC) != (p2[0] & 0xFC):
        assert ref_merge(ref, [p1, p2]) == our_merge([p1, p2]) == -4
    n = 5760 // g1["frame_size"] + 1
    assert ref_merge(ref, [p1] * n) == our_merge([p1] * n) == -4
    assert ref_merge(ref, [p1, p1], 1, 1) == our_merge([p1, p1], 1, 1) == -1
    assert ref_merge(ref, [p1, p1], 0, 3) == our_merge([p1, p1], 0, 3) == -1


def test_repacketizer_object_reset_and_frames():
    from opus_codec_b200.packet import Repacketizer
    g = load_golden(golden_names()[0])
    fr = [bytes(g["packets"][0][f, :g["lens"][0][f]]) for f in range(4)]
    with Repacketizer() as rp:
        assert rp.frames() == 0
        rp.push(fr[0]); rp.push(repacketize(fr[1:3], 3))
        assert rp.frames() == 3
        whole = rp.out()
        assert rp.out_range(1, 2) == bytes([fr[1][0] & 0xFC]) + fr[1][1:]
        rp.reset()
        assert rp.frames() == 0
        rp.push(fr[3])
        assert rp.out() == bytes([fr[3][0] & 0xFC]) + fr[3][1:]
    from opus_codec_b200.packet import packet_parse
    assert packet_parse(whole)[2] == [f[1:] for f in fr[:3]]


def test_pad_unpad_and_extensions_match_reference(ref):
    from opus_codec_b200 import packet as pkt_mod
    from opus_codec_b200.packet import packet_pad
    from opus_codec_b200.batch import OpusError

    def packet_unpad(p):
        try:
            return pkt_mod.packet_unpad(p)
        except OpusError as e:
            return e.code

    def ref_unpad(p):
        d = _buf(p)
        n = ref.opus_packet_unpad(d, len(p))
        return bytes(d[:n]) if n >= 0 else n
    rng = np.random.default_rng(21)
    pkts = sample_packets()
    g = load_golden(golden_names()[0])
    fr = [bytes(g["packets"][0][f, :g["lens"][0][f]]) for f in range(6)]
    ext = [with_padding(fr[:1], ext_padding()), with_padding(fr[:3], ext_padding()), with_padding(fr[:2], bytes([0x01] * 5) + ext_padding()),
           with_padding(fr[:2], bytes([(3 << 1) | 1, 7, 0x02, 0x02, (100 << 1) | 1, 255, 45]) + bytes(range(256)) + bytes(44) + bytes([(2 << 1)]))]
    for p in pkts + ext:
        for new_len in (len(p), len(p) + 1, len(p) + 2, len(p) + int(rng.integers(3, 600)), len(p) + 255, len(p) + 256, len(p) + 257):
            d = _buf(p, new_len)
            r = ref.opus_packet_pad(d, len(p), new_len)
            try:
                ours = packet_pad(p, new_len)
                assert r == 0 and ours == bytes(d[:new_len]), (p[:4].hex(), new_len)
                assert packet_unpad(ours) == ref_unpad(ours)
            except OpusError as e:
                assert e.code == r, (p[:4].hex(), new_len)
        assert packet_unpad(p) == ref_unpad(p)
    # extensions travel through the repacketizer: merged packets renumber the frames they belong to
    for grp in ([ext[0], fr[3]], [fr[3], ext[0]], [ext[1], ext[0]], [ext[2], ext[2]], [ext[0]], [ext[3], ext[1]]):
        a, b = ref_merge(ref, grp), our_merge(grp)
        assert isinstance(a, bytes) and a == b
    assert ref_merge(ref, [ext[1], ext[0]], 1, 4) == our_merge([ext[1], ext[0]], 1, 4)
    with pytest.raises(OpusError) as e:
        packet_pad(fr[0], len(fr[0]) - 1)
    assert e.value.code == -1


def test_self_delimited_framing_matches_reference_internals(ref):
    """opus_packet_parse_impl / opus_repacketizer_out_range_impl with self_delimited = 1 (the framing of all but the last stream of a
    multistream packet): the restatement in csrc/repacketizer.cuh, compiled by g++, against the reference's internal functions."""
    import os, subprocess, tempfile
    from conftest import ROOT
    so = os.path.join(tempfile.gettempdir(), "ob_emul_pkt_%d.so" % os.getpid())
    subprocess.run(["g++", "-O1", "-std=c++17", "-shared", "-fPIC", "-ffp-contract=off", "-Wno-unknown-pragmas", "-o", so,
                    os.path.join(ROOT, "tests", "host_emul", "emul.cpp")], check=True)
    E = C.CDLL(so)
    vp, i32 = C.c_void_p, C.c_int32
    ref.opus_packet_parse_impl.argtypes = [vp, i32, C.c_int, vp, vp, vp, vp, vp, vp, vp]; ref.opus_packet_parse_impl.restype = C.c_int
    ref.opus_repacketizer_out_range_impl.argtypes = [vp, C.c_int, C.c_int, vp, i32, C.c_int, C.c_int, vp, C.c_int]
    ref.opus_repacketizer_out_range_impl.restype = i32
    rng = np.random.default_rng(17)
    pkts = [p for p in sample_packets() if len(p) < 4000]
    g = load_golden(golden_names()[0])
    fr = [bytes(g["packets"][0][f, :g["lens"][0][f]]) for f in range(6)]
    pkts += [with_padding(fr[:2], ext_padding())]
    made = []
    for p in pkts:                                            # plain packet -> self-delimited packet, both implementations
        for pad in (0, 1):
            maxlen = len(p) + 3 + (40 if pad else 0)
            rp = ref.opus_repacketizer_create()
            b = _buf(p)
            if ref.opus_repacketizer_cat(rp, b, len(p)) != 0:
                ref.opus_repacketizer_destroy(rp); continue
            nfr = ref.opus_repacketizer_get_nb_frames(rp)
            want = (C.c_uint8 * maxlen)()
            n0 = ref.opus_repacketizer_out_range_impl(rp, 0, nfr, want, maxlen, 1, pad, None, 0)
            ref.opus_repacketizer_destroy(rp)
            got = (C.c_uint8 * maxlen)()
            off = (C.c_int * 1)(0); ln = (C.c_int * 1)(len(p))
            n1 = E.emul_repacketize_self_delimited(b, off, ln, 1, 0, nfr, got, maxlen, pad)
            assert n0 == n1 and (n0 < 0 or bytes(want[:n0]) == bytes(got[:n0])), (p[:4].hex(), pad, n0, n1)
            if n0 > 0:
                made.append(bytes(want[:n0]))
    assert len(made) > 60
    # parse them back (self-delimited), followed by trailing bytes of another packet, plus damaged variants
    cases = []
    for m in made:
        cases.append(m + fr[5])
        q = bytearray(m + fr[5]); q[int(rng.integers(0, min(len(q), 8)))] ^= 1 << int(rng.integers(0, 8)); cases.append(bytes(q))
        cases.append(m[:int(rng.integers(1, len(m)))])
    seen = set()
    for c in cases:
        b = _buf(c)
        toc0 = C.c_uint8(0); po0 = C.c_int(0); pko0 = C.c_int32(0); pad0 = C.c_int32(0); padp = C.c_void_p(0)
        frames = (C.c_void_p * 48)(); sz0 = (C.c_int16 * 48)()
        n0 = ref.opus_packet_parse_impl(b, len(c), 1, C.byref(toc0), frames, sz0, C.byref(po0), C.byref(pko0), C.byref(padp), C.byref(pad0))
        toc1 = C.c_uint8(0); po1 = C.c_int(0); pko1 = C.c_int(0); pad1 = C.c_int(0)
        offs = (C.c_int * 48)(); sz1 = (C.c_int16 * 48)()
        n1 = E.emul_parse_packet(b, len(c), 1, C.byref(toc1), offs, sz1, C.byref(po1), C.byref(pko1), C.byref(pad1))
        assert n0 == n1, (c[:6].hex(), n0, n1)
        seen.add(n0 if n0 < 0 else "ok")
        if n0 > 0:
            base = C.addressof(b)
            assert toc0.value == toc1.value and po0.value == po1.value and pko0.value == pko1.value and pad0.value == pad1.value
            assert list(sz0[:n0]) == list(sz1[:n0]) and [frames[i] - base for i in range(n0)] == list(offs[:n0])
    assert {"ok", -4} <= seen
    os.remove(so)


def test_multistream_packet_pad_unpad_match_reference(ref):
    """multistream_packet_pad / _unpad (src/packet.rs:253-290): 2- and 3-stream packets built with the reference's self-delimited output."""
    from opus_codec_b200.packet import multistream_packet_pad, multistream_packet_unpad
    from opus_codec_b200.batch import OpusError
    vp, i32 = C.c_void_p, C.c_int32
    ref.opus_repacketizer_out_range_impl.argtypes = [vp, C.c_int, C.c_int, vp, i32, C.c_int, C.c_int, vp, C.c_int]
    ref.opus_repacketizer_out_range_impl.restype = i32
    ref.opus_multistream_packet_pad.argtypes = [vp, i32, i32, C.c_int]; ref.opus_multistream_packet_pad.restype = C.c_int
    ref.opus_multistream_packet_unpad.argtypes = [vp, i32, C.c_int]; ref.opus_multistream_packet_unpad.restype = i32

    def self_delimited(p):
        rp = ref.opus_repacketizer_create()
        b = _buf(p)
        assert ref.opus_repacketizer_cat(rp, b, len(p)) == 0
        out = (C.c_uint8 * (len(p) + 4))()
        n = ref.opus_repacketizer_out_range_impl(rp, 0, ref.opus_repacketizer_get_nb_frames(rp), out, len(p) + 4, 1, 0, None, 0)
        ref.opus_repacketizer_destroy(rp)
        assert n > 0
        return bytes(out[:n])

    g1, g2 = load_golden("cfg3_stereo_20ms_96k_cbr"), load_golden("cfg2_mono_20ms_64k_cbr")
    a = [bytes(g1["packets"][0][f, :g1["lens"][0][f]]) for f in range(4)]
    m = [bytes(g2["packets"][0][f, :g2["lens"][0][f]]) for f in range(4)]
    streams = [([a[0], m[0]], 2), ([repacketize(a[:2], 3, pad=30), m[1]], 2), ([a[1], repacketize(m[:3], 3, pad=5), m[3]], 3),
               ([with_padding(a[2:3], ext_padding()), a[3]], 2)]
    for parts, nb in streams:
        ms = b"".join(self_delimited(p) for p in parts[:-1]) + parts[-1]
        for new_len in (len(ms), len(ms) + 1, len(ms) + 2, len(ms) + 77, len(ms) + 300):
            d = _buf(ms, new_len)
            r = ref.opus_multistream_packet_pad(d, len(ms), new_len, nb)
            assert r == 0
            padded = multistream_packet_pad(ms, new_len, nb)
            assert padded == bytes(d[:new_len])
            d2 = _buf(padded)
            n = ref.opus_multistream_packet_unpad(d2, len(padded), nb)
            assert n > 0 and multistream_packet_unpad(padded, nb) == bytes(d2[:n])
        d = _buf(ms)
        n = ref.opus_multistream_packet_unpad(d, len(ms), nb)
        assert multistream_packet_unpad(ms, nb) == bytes(d[:n])
        # wrong stream count: same error as the reference
        d = _buf(ms)
        n = ref.opus_multistream_packet_unpad(d, len(ms), nb + 2)
        try:
            got = multistream_packet_unpad(ms, nb + 2)
            assert n > 0 and got == bytes(d[:n])
        except OpusError as e:
            assert e.code == n


def test_packet_nb_samples_and_has_lbrr_match_reference(ref):
    """packet_nb_samples / packet_has_lbrr (src/packet.rs:72-120) on CELT, SILK and hybrid packets (the latter two from the reference's VOIP
    encoder with in-band FEC on) and on random TOCs."""
    from opus_codec_b200 import _lib
    from oracle import refpy
    L = _lib.lib()
    ref.opus_packet_get_nb_samples.argtypes = [C.c_void_p, C.c_int32, C.c_int32]; ref.opus_packet_get_nb_samples.restype = C.c_int
    ref.opus_packet_has_lbrr.argtypes = [C.c_void_p, C.c_int32]; ref.opus_packet_has_lbrr.restype = C.c_int
    pkts = [p for p in sample_packets() if len(p) < 4000]
    # SILK / hybrid packets with LBRR: VOIP, 24 kb/s, FEC on, 20 % loss
    u8p, i32p, u32p, f32p = C.POINTER(C.c_ubyte), C.POINTER(C.c_int), C.POINTER(C.c_uint32), C.POINTER(C.c_float)
    ref.ref_encode_stream.argtypes = [f32p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, u8p, C.c_int, i32p, u32p]
    from opus_codec_b200 import synth
    for ch, fs in ((1, 960), (2, 960), (1, 2880)):
        pcm = np.ascontiguousarray(synth.stream_pcm(2, 48000, ch))
        nf = pcm.size // (fs * ch)
        out = np.zeros((nf, 600), np.uint8); lens = np.zeros(nf, np.int32); rg = np.zeros(nf, np.uint32)
        ref.ref_set_encoder_force_celt(0); ref.ref_set_encoder_extras2(3001, 0, 0, 0, 1, 0, 20)
        try:
            assert ref.ref_encode_stream(pcm.ctypes.data_as(f32p), nf, fs, ch, 2048, 24000, 1, 5, out.ctypes.data_as(u8p), 600, lens.ctypes.data_as(i32p), rg.ctypes.data_as(u32p)) == 0
        finally:
            ref.ref_set_encoder_force_celt(1); ref.ref_set_encoder_extras2(0, 0, 0, 0, 0, 0, 0)
        pkts += [bytes(out[f, :lens[f]]) for f in range(nf)]
    rng = np.random.default_rng(9)
    pkts += [bytes(rng.integers(0, 256, int(rng.integers(2, 30)), dtype=np.uint8)) for _ in range(2000)]
    lbrr_seen = set()
    for p in pkts:
        b = _buf(p, len(p) + 1)                                 # one zero byte behind: the reference peeks at frames[0][0] even when that frame is empty
        for fs in (48000, 16000, 8000):
            assert ref.opus_packet_get_nb_samples(b, len(p), fs) == L.ob_packet_get_nb_samples(bytes(p), len(p), fs)
        a = ref.opus_packet_has_lbrr(b, len(p))
        assert a == L.ob_packet_has_lbrr(bytes(p), len(p)), p[:6].hex()
        lbrr_seen.add(a)
    assert {0, 1} <= lbrr_seen
