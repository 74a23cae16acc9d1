"""The reference crate's own integration tests (tests/encoder_ctls.rs, decoder_ctls.rs, opus_tests.rs), restated against the batched API:
same objects, same calls, same assertions -- for every stream of a small batch.  Where the reference test drives libopus into SILK / hybrid
(Application::Voip on silence at the default bitrate) the batched path must say Unimplemented for exactly those packets."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

AUDIO, VOIP = 2049, 2048
S = 3


def _ref_encode(pcm, fs, ch, app, float_api=True):
    """The reference with the crate's defaults (bitrate AUTO, VBR, complexity 9), its own mode decision."""
    from oracle import refpy
    L = refpy.lib_c()
    u8p, i32p, u32p = C.POINTER(C.c_ubyte), C.POINTER(C.c_int), C.POINTER(C.c_uint32)
    nf = pcm.size // (fs * ch)
    out = np.zeros((nf, 500), np.uint8); lens = np.zeros(nf, np.int32); rng = np.zeros(nf, np.uint32)
    fn = L.ref_encode_stream if float_api else L.ref_encode_stream_i16
    ptr = C.POINTER(C.c_float) if float_api else C.POINTER(C.c_int16)
    fn.argtypes = [ptr, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, u8p, C.c_int, i32p, u32p]
    L.ref_set_encoder_force_celt(0)
    try:
        assert fn(pcm.ctypes.data_as(ptr), nf, fs, ch, app, -1000, 2, 9, out.ctypes.data_as(u8p), 500, lens.ctypes.data_as(i32p), rng.ctypes.data_as(u32p)) == 0
    finally:
        L.ref_set_encoder_force_celt(1)
    return [bytes(out[f, :lens[f]]) for f in range(nf)]


def test_encoder_control_roundtrip():
    """tests/encoder_ctls.rs::encoder_control_roundtrip"""
    from opus_codec_b200.batch import BatchEncoder
    with BatchEncoder(S, 48000, 2, application=AUDIO, device=0) as encoder:
        encoder.set_bitrate(96_000)
        assert encoder.bitrate() == 96_000
        encoder.set_complexity(4)
        assert encoder.complexity() == 4
        encoder.set_vbr(False)
        assert not encoder.vbr()
        encoder.set_vbr_constraint(True)
        assert encoder.vbr_constraint()
        encoder.set_inband_fec(True)
        assert encoder.inband_fec()
        encoder.set_packet_loss_perc(15)
        assert encoder.packet_loss_perc() == 15
        encoder.set_signal(3002)                                 # Signal::Music
        assert encoder.signal() == 3002
        encoder.set_max_bandwidth(1103)                          # Bandwidth::Wideband
        assert encoder.max_bandwidth() == 1103
        encoder.set_force_channels(1)                            # Some(Channels::Mono)
        assert encoder.force_channels() == 1
        encoder.set_force_channels(-1000)                        # None
        assert encoder.force_channels() == -1000


def test_decoder_control_roundtrip():
    """tests/decoder_ctls.rs::decoder_control_roundtrip"""
    from opus_codec_b200.batch import BatchDecoder
    with BatchDecoder(S, 48000, 2, device=0) as decoder:
        decoder.set_gain(256)
        assert decoder.gain() == 256
        decoder.set_phase_inversion_disabled(True)
        assert decoder.phase_inversion_disabled()
        decoder.set_phase_inversion_disabled(False)
        assert not decoder.phase_inversion_disabled()
        assert decoder.sample_rate == 48000 and decoder._L.ob_decoder_sample_rate(decoder._h) == 48000      # get_sample_rate
        assert (decoder.last_packet_duration() == 0).all()       # before any decode


def test_packet_analysis(have_ref):
    """tests/opus_tests.rs::test_packet_analysis -- and the packet is the reference's packet."""
    from opus_codec_b200 import _lib
    from opus_codec_b200.batch import BatchEncoder
    from opus_codec_b200.packet import packet_parse
    L = _lib.lib()
    pcm = np.zeros((S, 1, 960 * 2), np.int16)                    # 20 ms stereo, silent
    with BatchEncoder(S, 48000, 2, application=AUDIO, device=0) as encoder:
        out, lens, _ = encoder.encode_multi(pcm, 960, max_bytes=100)
    for s in range(S):
        packet = bytes(out[s, 0, :lens[s, 0]])
        assert L.ob_packet_get_nb_frames(packet, len(packet)) > 0
        assert L.ob_packet_get_nb_frames(packet, len(packet)) * L.ob_packet_get_samples_per_frame(packet, 48000) == 960
        assert L.ob_packet_get_nb_channels(packet) == 2
        assert L.ob_packet_get_bandwidth(packet) != 1101         # != Narrowband
        _toc, _offset, frames = packet_parse(packet)
        assert len(frames) > 0 or len(packet) == 1
        if have_ref:
            assert packet == _ref_encode(np.zeros(960 * 2, np.int16), 960, 2, AUDIO, float_api=False)[0]


def test_float_api(have_ref):
    """tests/opus_tests.rs::test_float_api"""
    from opus_codec_b200.batch import BatchDecoder, BatchEncoder, pack_packets
    frame_size = 480                                             # 10 ms
    pcm_in = np.zeros((S, 1, frame_size * 2), np.float32)
    with BatchEncoder(S, 48000, 2, application=AUDIO, device=0) as encoder, BatchDecoder(S, 48000, 2, device=0) as decoder:
        out, lens, _ = encoder.encode_float_multi(pcm_in, frame_size, max_bytes=500)
        assert (lens > 0).all()
        pk = [[bytes(out[s, 0, :lens[s, 0]])] for s in range(S)]
        b, o, l = pack_packets(pk)
        pcm_out, decoded_len, _ = decoder.decode_float_multi(b, o, l, frame_size)
        assert (decoded_len == frame_size).all()
    if have_ref:
        assert pk[0][0] == _ref_encode(np.zeros(frame_size * 2, np.float32), frame_size, 2, AUDIO)[0]


def test_encode_decode_voip_default_rate_is_silk_territory(have_ref):
    """tests/opus_tests.rs::test_encode_decode uses Application::Voip, mono, silence, default bitrate: libopus codes that with SILK / hybrid.
    The batched path reports Unimplemented for those packets (CELT-only scope); at a bitrate where libopus stays CELT-only the test's
    assertions hold."""
    from opus_codec_b200.batch import BatchDecoder, BatchEncoder, pack_packets
    frame_size = 960
    pcm_in = np.zeros((S, 1, frame_size), np.int16)
    if have_ref:
        ref = _ref_encode(np.zeros(frame_size, np.int16), frame_size, 1, VOIP, float_api=False)[0]
        assert not ref[0] & 0x80                                  # the reference's packet is not CELT-only
    with BatchEncoder(S, 48000, 1, application=VOIP, device=0) as encoder:
        _, lens, _ = encoder.encode_multi(pcm_in, frame_size, max_bytes=500)
        assert (lens == -5).all()
    with BatchEncoder(S, 48000, 1, application=VOIP, device=0) as encoder, BatchDecoder(S, 48000, 1, device=0) as decoder:
        encoder.set_bitrate(96_000)
        out, lens, _ = encoder.encode_multi(pcm_in, frame_size, max_bytes=500)
        assert (lens > 0).all()
        b, o, l = pack_packets([[bytes(out[s, 0, :lens[s, 0]])] for s in range(S)])
        pcm_out, decoded_len, _ = decoder.decode_multi(b, o, l, frame_size)
        assert (decoded_len == frame_size).all()


def test_repacketizer(have_ref):
    """tests/opus_tests.rs::test_repacketizer (its packets come from a Voip encoder: SILK / hybrid packets are fine for the repacketizer)."""
    from opus_codec_b200 import _lib
    from opus_codec_b200.packet import Repacketizer
    if not have_ref:
        pytest.skip("oracle/_ref not built")
    packet1, packet2 = _ref_encode(np.zeros(960 * 2, np.int16), 960, 1, VOIP, float_api=False)
    rp = Repacketizer()
    rp.push(packet1)
    rp.push(packet2)
    assert rp.frames() == 2
    merged = rp.out(500)
    assert len(merged) > 0
    assert _lib.lib().ob_packet_get_nb_frames(merged, len(merged)) == 2


def test_buffer_empty():
    """tests/opus_tests.rs::test_buffer_empty: an empty output buffer is BadArg."""
    from opus_codec_b200.batch import BatchEncoder, OpusError
    with BatchEncoder(S, 48000, 1, application=VOIP, device=0) as encoder:
        with pytest.raises(OpusError) as e:
            encoder.encode_multi(np.zeros((S, 1, 960), np.int16), 960, max_bytes=0)
        assert e.value.code == -1


def test_soft_clip_validations():
    """tests/opus_tests.rs::test_soft_clip_validations: argument checks of soft_clip, then a clipped signal comes back within [-1, 1]."""
    from opus_codec_b200.batch import OpusError
    from opus_codec_b200.packet import soft_clip_batch
    mem = np.zeros((S, 2), np.float32)
    for bad in (lambda: soft_clip_batch(np.zeros((S, 7), np.float32), 2, mem),                 # not a whole number of stereo samples
                lambda: soft_clip_batch(np.zeros((S, 8), np.float32), 2, np.zeros((S, 1), np.float32)),   # soft-clip memory too short
                lambda: soft_clip_batch(np.zeros((S, 8), np.float64), 2, mem)):
        with pytest.raises(OpusError) as e:
            bad()
        assert e.value.code == -1
    x = np.tile(np.array([1.7, -1.9], np.float32), (S, 240))
    soft_clip_batch(x, 2, mem)
    assert np.abs(x).max() <= 1.0


def _pcm_frame():
    """tests/multhithread.rs::pcm_frame: a sawtooth-like 20 ms stereo frame."""
    i = np.arange(960 * 2, dtype=np.int64)
    return (((i * 17) & 0xFFFF).astype(np.uint16).view(np.int16) - np.int16(16000)).astype(np.int16)


def test_multithread_smoke():
    """tests/multhithread.rs::{encoder,decoder}_multithread_smoke: 4 threads, each with its own batch encoder / decoder (objects are Send, not
    shared), 16 iterations; every thread must produce what a single-threaded run produces."""
    import threading
    from opus_codec_b200.batch import BatchDecoder, BatchEncoder, pack_packets
    THREADS, ITERATIONS = 4, 16
    frame = np.tile(_pcm_frame(), (S, 1, 1))

    def encode_run(result, k):
        with BatchEncoder(S, 48000, 2, application=AUDIO, device=0) as encoder:
            got = []
            for _ in range(ITERATIONS):
                out, lens, rng = encoder.encode_multi(frame, 960, max_bytes=4096)
                assert (lens > 0).all()
                got.append((out[:, 0, :int(lens.max())].copy(), lens.copy(), rng.copy()))
            result[k] = got

    res = [None] * (THREADS + 1)
    encode_run(res, THREADS)                                      # single-threaded yardstick
    th = [threading.Thread(target=encode_run, args=(res, k)) for k in range(THREADS)]
    [t.start() for t in th]; [t.join() for t in th]
    for k in range(THREADS):
        assert res[k] is not None
        for a, b in zip(res[k], res[THREADS]):
            assert all(np.array_equal(x, y) for x, y in zip(a, b))
    out, lens, _ = res[THREADS][0]
    packet = [[bytes(out[s, :lens[s, 0]])] for s in range(S)]

    def decode_run(result, k):
        with BatchDecoder(S, 48000, 2, device=0) as decoder:
            got = []
            b, o, l = pack_packets(packet)
            for _ in range(ITERATIONS):
                pcm, smp, rng = decoder.decode_multi(b, o, l, 960)
                assert (smp == 960).all()
                got.append((pcm.copy(), rng.copy()))
            result[k] = got

    res = [None] * (THREADS + 1)
    decode_run(res, THREADS)
    th = [threading.Thread(target=decode_run, args=(res, k)) for k in range(THREADS)]
    [t.start() for t in th]; [t.join() for t in th]
    for k in range(THREADS):
        assert res[k] is not None
        for a, b in zip(res[k], res[THREADS]):
            assert all(np.array_equal(x, y) for x, y in zip(a, b))
