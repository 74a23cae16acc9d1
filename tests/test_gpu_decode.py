"""GPU parity tests (run on the B200 box with -m gpu): the CUDA path, called through the C ABI
(include/opus_b200.h via opus_codec_b200.batch), against the golden vectors the reference produced,
against the oracle on the same inputs, and -- at full batch size -- through tiling properties.

Bars (BASELINE.json north_star): OPUS_GET_FINAL_RANGE bit-exact; float PCM max |err| <= 1e-4 of full scale
and opus_compare pass."""
import os
import tempfile

import numpy as np
import pytest

from conftest import plc_golden_names, load_plc_golden, multiframe_stream, golden_names, load_golden

pytestmark = pytest.mark.gpu
PCM_TOL = 1e-4


def _offsets(S, F, stride):
    return (np.arange(S * F, dtype=np.int32) * stride).reshape(S, F)


def _decode_all(g, max_frames=None, streams=None):
    from opus_codec_b200.batch import BatchDecoder
    pk = g["packets"] if streams is None else g["packets"][streams]
    ln = g["lens"] if streams is None else g["lens"][streams]
    S, F, stride = pk.shape
    with BatchDecoder(S, 48000, g["dec_channels"], device=0, max_frames=F) as dec:
        return dec.decode_float_multi(pk.reshape(-1), _offsets(S, F, stride), ln, g["frame_size"])


@pytest.mark.parametrize("name", golden_names())
def test_golden_final_range_and_pcm(name, have_ref):
    from oracle import oraclepy
    g = load_golden(name)
    pcm, samples, ranges = _decode_all(g)
    assert (samples == g["frame_size"]).all()
    assert (ranges == g["dec_rng"]).all(), "OPUS_GET_FINAL_RANGE differs from the reference decoder"
    npcm = g["pcm"].shape[0]
    assert np.abs(pcm[:npcm] - g["pcm"]).max() <= PCM_TOL
    for s in range(g["packets"].shape[0]):          # every stream against the oracle on the same packets
        opcm, orng, _ = oraclepy.decode_stream(g["packets"][s], g["lens"][s], g["frame_size"], g["dec_channels"])
        assert (ranges[s] == orng).all()
        assert np.abs(pcm[s] - opcm).max() <= PCM_TOL
    if have_ref:                                     # ... and EVERY stream against the unmodified reference decoder run here on the same packets
        from oracle import refpy                     # (the stored fixture keeps the reference's PCM of one stream per configuration)
        for s in range(g["packets"].shape[0]):
            rpcm, rrng, rsmp = refpy.decode_stream(g["packets"][s], g["lens"][s], g["frame_size"], g["dec_channels"])
            assert (rsmp == samples[s]).all() and (rrng == ranges[s]).all()
            assert np.abs(pcm[s] - rpcm).max() <= PCM_TOL, (name, s)


def test_streaming_one_frame_per_call_matches_multi():
    """State must persist across calls exactly like Decoder::decode_float called once per packet."""
    from opus_codec_b200.batch import BatchDecoder
    g = load_golden("cfg1_stereo_20ms_128k_cbr")
    S, F, stride = g["packets"].shape
    F = 20
    with BatchDecoder(S, 48000, 2, device=0, max_frames=1) as dec:
        assert (dec.final_range() == 0).all() and (dec.last_packet_duration() == 0).all()
        for f in range(F):
            pcm, smp = dec.decode_float([bytes(g["packets"][s, f, :g["lens"][s, f]]) for s in range(S)], 960)
            assert (smp == 960).all()
            assert (dec.final_range() == g["dec_rng"][:, f]).all()
            assert np.abs(pcm[0] - g["pcm"][0, f]).max() <= PCM_TOL
        assert (dec.last_packet_duration() == 960).all()


def test_mixed_frame_sizes_in_one_batch():
    """BASELINE config 4: 2.5/5/10/20 ms streams (incl. transient class) decoded by ONE launch."""
    from opus_codec_b200.batch import BatchDecoder
    names = ["cfg4_stereo_2p5ms_96k", "cfg4_stereo_5ms_96k", "cfg4_stereo_10ms_96k", "cfg3_stereo_20ms_96k_cbr"]
    gs = [load_golden(n) for n in names]
    F = 50
    stride = max(g["packets"].shape[2] for g in gs)
    pk = np.zeros((len(gs) * 3, F, stride), np.uint8)
    ln = np.zeros((len(gs) * 3, F), np.int32)
    for k, g in enumerate(gs):
        pk[3 * k:3 * k + 3, :, :g["packets"].shape[2]] = g["packets"][:3, :F]
        ln[3 * k:3 * k + 3] = g["lens"][:3, :F]
    S = pk.shape[0]
    with BatchDecoder(S, 48000, 2, device=0, max_frames=F) as dec:
        pcm, samples, ranges = dec.decode_float_multi(pk.reshape(-1), _offsets(S, F, stride), ln, 960)
    for k, g in enumerate(gs):
        fs = g["frame_size"]
        assert (samples[3 * k:3 * k + 3] == fs).all()
        assert (ranges[3 * k:3 * k + 3] == g["dec_rng"][:3, :F]).all()
        assert np.abs(pcm[3 * k, :, :fs * 2] - g["pcm"][0, :F]).max() <= PCM_TOL


def test_per_stream_errors_do_not_disturb_neighbours():
    from opus_codec_b200.batch import BatchDecoder, UNIMPLEMENTED, BUFFER_TOO_SMALL, INVALID_PACKET
    g = load_golden("cfg2_mono_20ms_64k_cbr")
    S, F = 6, 10
    pk = g["packets"][:S, :F].copy()
    ln = g["lens"][:S, :F].copy()
    stride = pk.shape[2]
    ln[1, 3] = 0                      # lost packet -> concealed: 960 samples, final range 0
    pk[2, 4, 0] = 0x08                # SILK TOC
    pk[3, 5, 0] = 0xF9                # code-1 packet whose payload is not two equal halves (159 bytes): invalid (opus.c:240-242)
    ln[4, 6] = 2                      # 1-byte payload -> DTX, concealed for the TOC's duration
    with BatchDecoder(S, 48000, 1, device=0, max_frames=F) as dec:
        pcm, samples, ranges = dec.decode_float_multi(pk.reshape(-1), _offsets(S, F, stride), ln, 960)
    assert samples[2, 4] == UNIMPLEMENTED and samples[3, 5] == INVALID_PACKET
    assert samples[1, 3] == 960 and samples[4, 6] == 960 and ranges[1, 3] == 0 and ranges[4, 6] == 0
    good = np.ones((S, F), bool)
    for s, f in ((2, 4), (3, 5)):
        good[s, f] = False
    assert (samples[good] == 960).all()
    assert (ranges[0] == g["dec_rng"][0, :F]).all() and (ranges[5] == g["dec_rng"][5, :F]).all()
    assert np.abs(pcm[0] - g["pcm"][0, :F]).max() <= PCM_TOL
    # frames before the damaged one are untouched
    assert (ranges[1, :3] == g["dec_rng"][1, :3]).all()
    # 20 ms packets into a 10 ms slot
    with BatchDecoder(S, 48000, 1, device=0, max_frames=F) as dec:
        _, samples, _ = dec.decode_float_multi(np.ascontiguousarray(g["packets"][:S, :F]).reshape(-1), _offsets(S, F, stride), g["lens"][:S, :F], 480)
    assert (samples == BUFFER_TOO_SMALL).all()


@pytest.mark.parametrize("base", plc_golden_names())
def test_packet_loss_concealment_matches_reference(base):
    """Lost packets (len 0), DTX payloads (<= 1 byte), noise- and pitch-based concealment, recovery after a loss: samples and
    final range bit-exact, PCM within 1e-4 of the reference's pure-C build (the reference's own SSE build differs from its C build
    by up to 2e-3 in concealed frames -- stored per frame in the fixture as sse_diff -- so the C build is the bar)."""
    from opus_codec_b200.batch import BatchDecoder
    g, p = load_golden(base), load_plc_golden(base)
    fs, dc = g["frame_size"], g["dec_channels"]
    S, F = p["lens"].shape
    pk = np.ascontiguousarray(g["packets"][:S, :F]); stride = pk.shape[2]
    with BatchDecoder(S, 48000, dc, device=0, max_frames=F) as dec:
        pcm, samples, ranges = dec.decode_float_multi(pk.reshape(-1), _offsets(S, F, stride), p["lens"], fs)
        assert (samples == p["samples"]).all() and (ranges == p["ranges"]).all()
        err = np.abs(pcm.reshape(S, F, -1) - p["pcm_c"]).max(axis=2)
        # Frames on which the reference's two builds agree (sse_diff <= 1e-6: everything but concealed frames and the few
        # recovery frames after them) must meet the 1e-4 bar.  Concealment re-derives an order-24 LPC filter from the decoded
        # history; that analysis amplifies last-bit differences of the history by ~1e3 (the reference's SSE build is up to 2e-3
        # away from its C build there), so those frames are held to the reference's own spread instead.  The exact arithmetic
        # of the concealment is pinned to 1e-6 by tests/test_host_emul.py.
        calm = p["sse_diff"] <= 1e-6
        assert err[calm].max() <= 1e-4, (float(err[calm].max()), np.argwhere(calm & (err > 1e-4))[:5].tolist())
        assert err.max() <= 3e-3, (float(err.max()), np.argwhere(err > 3e-3)[:5].tolist())
        # the same packets one call per frame (live streaming): identical output
        dec.reset()
        for f in range(F):
            one, smp1, rng1 = dec.decode_float_multi(pk.reshape(-1), _offsets(S, F, stride)[:, f:f + 1], p["lens"][:, f:f + 1], fs)
            assert (smp1[:, 0] == p["samples"][:, f]).all() and (rng1[:, 0] == p["ranges"][:, f]).all()
            assert np.array_equal(one.reshape(S, -1), pcm.reshape(S, F, -1)[:, f])


def test_lost_packet_slot_larger_than_20ms_is_concealed_in_pieces(have_ref):
    """A lost packet conceals the caller's whole slot; slots that are not a CELT frame size are covered by several concealment
    frames (opus_decoder.c:313-335): 1440 = 960 + 480, 720 = 480 + 240."""
    from oracle import refpy
    if not have_ref or not os.path.exists(os.path.join(os.path.dirname(refpy.OPUS_COMPARE), "libopus_ref_c.so")):
        pytest.skip("oracle/_ref not built")
    from opus_codec_b200.batch import BatchDecoder
    g = load_golden("cfg2_mono_20ms_64k_cbr")
    S, F = 2, 8
    pk = np.ascontiguousarray(g["packets"][:S, :F]); stride = pk.shape[2]
    for slot in (1440, 720):
        ln = g["lens"][:S, :F].copy()
        ln[:, 4] = 0; ln[1, 5] = 0
        with BatchDecoder(S, 48000, 1, device=0, max_frames=F) as dec:
            pcm, samples, ranges = dec.decode_float_multi(pk.reshape(-1), _offsets(S, F, stride), ln, slot)
        for s in range(S):
            rp, rr, rs = refpy.decode_stream(pk[s], ln[s], slot, 1, pure_c=True)
            assert (samples[s] == rs).all() and (ranges[s] == rr).all()
            for f in range(F):
                assert np.abs(pcm.reshape(S, F, -1)[s, f, :rs[f]] - rp[f, :rs[f]]).max() <= (1e-4 if f < 4 else 3e-3)


def test_decode_gain_and_phase_inversion_ctls(have_ref):
    """OPUS_SET_GAIN and OPUS_SET_PHASE_INVERSION_DISABLED (Decoder::set_gain / set_phase_inversion_disabled) against the
    reference decoder with the same CTLs; the final range never depends on either."""
    from opus_codec_b200.batch import BatchDecoder, OpusError, BAD_ARG
    from oracle import refpy
    g = load_golden("cfg3_stereo_20ms_96k_cbr")            # these packets carry the stereo inversion flag (checked against the reference)
    S, F, stride = g["packets"].shape
    S, F = 3, 25
    pk = np.ascontiguousarray(g["packets"][:S, :F]); ln = g["lens"][:S, :F]
    with BatchDecoder(S, 48000, 2, device=0, max_frames=F) as dec:
        base, _, rng0 = dec.decode_float_multi(pk.reshape(-1), _offsets(S, F, stride), ln, 960)
        base = base.copy()
        assert dec.gain() == 0 and dec.phase_inversion_disabled() is False
        with pytest.raises(OpusError) as e:
            dec.set_gain(40000)
        assert e.value.code == BAD_ARG
        for q8 in (-1536, 777):
            dec.reset(); dec.set_gain(q8)
            assert dec.gain() == q8
            out, _, rng = dec.decode_float_multi(pk.reshape(-1), _offsets(S, F, stride), ln, 960)
            assert (rng == rng0).all()
            lin = np.float32(np.exp(0.6931471805599453094 * np.float32(np.float32(6.48814081e-4) * np.float32(q8))))
            assert np.abs(out - base * lin).max() <= 1e-6
            if have_ref:
                ref = refpy.decode_stream(pk[0], ln[0], 960, 2, gain_q8=q8)[0]
                assert np.abs(out.reshape(S, F, -1)[0] - ref).max() <= PCM_TOL * max(1.0, float(lin))
        dec.reset(); dec.set_gain(0); dec.set_phase_inversion_disabled(True)
        assert dec.phase_inversion_disabled() is True
        out, _, rng = dec.decode_float_multi(pk.reshape(-1), _offsets(S, F, stride), ln, 960)
        assert (rng == rng0).all()
        assert np.abs(out - base).max() > 1e-3               # the flag does change the stereo image of these packets
        if have_ref:
            for s in range(S):
                ref = refpy.decode_stream(pk[s], ln[s], 960, 2, phase_inv_disabled=True)[0]
                assert np.abs(out.reshape(S, F, -1)[s] - ref).max() <= PCM_TOL


@pytest.mark.timeout(300)
@pytest.mark.parametrize("with_loss", [False, True])
def test_fuzz_garbage_packets_match_reference(have_ref, with_loss):
    """Random payloads behind valid CELT TOCs (every frame size / bandwidth / channel count mixed in one batch), optionally with
    lost packets and DTX payloads: the GPU must agree with the reference on samples and final range for every frame, and on
    PCM to 1e-4 of the signal's scale up to a stream's first concealed frame.  (Concealing GARBAGE is chaotic: the LPC synthesis
    filter fitted to noise-like history is near-unstable and amplifies last-bit differences exponentially within one frame, so
    PCM after a loss is compared on real signals only -- test_packet_loss_concealment_matches_reference -- and bit-exactly in
    tests/test_host_emul.py::test_device_code_vs_live_reference_fuzz.)"""
    if not have_ref:
        pytest.skip("oracle/_ref not built")
    from oracle import refpy
    from opus_codec_b200.batch import BatchDecoder
    rng = np.random.default_rng(31337 + with_loss)
    S, F, stride = 96, 8, 200
    pk = rng.integers(0, 256, (S, F, stride), dtype=np.uint8)
    ln = rng.integers(3, stride, (S, F)).astype(np.int32)
    cfg = 16 + rng.integers(0, 16, S)
    toc = (cfg << 3) | (rng.integers(0, 2, S) << 2)
    pk[:, :, 0] = toc[:, None]
    if with_loss:
        ln[rng.random((S, F)) < 0.2] = 0
        ln[rng.random((S, F)) < 0.05] = 2
    for dc in (1, 2):
        with BatchDecoder(S, 48000, dc, device=0, max_frames=F) as dec:
            pcm, samples, ranges = dec.decode_float_multi(pk.reshape(-1), _offsets(S, F, stride), ln, 960)
        pcm = pcm.reshape(S, F, 960 * dc)
        for s in range(S):
            fs = 120 << int(cfg[s] & 3)
            # a lost packet conceals the caller's whole 960-sample slot, a coded or DTX packet its own frame size
            ref, rr, rs = refpy.decode_stream(pk[s], ln[s], 960, dc, pure_c=True)
            assert (samples[s] == rs).all() and (ranges[s] == rr).all(), (s, samples[s], rs)
            for f in range(F):
                if ln[s, f] <= 2:
                    break
                a, b = pcm[s, f, :rs[f] * dc], ref[f, :rs[f] * dc]
                m = np.isfinite(b)
                assert (m == np.isfinite(a)).all()
                scale = max(1.0, float(np.abs(b[m]).max())) if m.any() else 1.0
                assert np.abs(a[m] - b[m]).max() <= 1e-4 * scale, (s, f, fs)


def test_pipelined_async_calls_match_blocking_calls():
    """ob_decode_float_multi_async with two calls in flight produces exactly what consecutive blocking calls produce."""
    from opus_codec_b200.batch import BatchDecoder
    g = load_golden("cfg3_stereo_20ms_96k_cbr")
    S, Ftot, stride = g["packets"].shape
    F, ncalls = 5, 8
    offs = _offsets(S, Ftot, stride)
    with BatchDecoder(S, 48000, 2, device=0, max_frames=F) as dec:
        want = [dec.decode_float_multi(g["packets"].reshape(-1), offs[:, k * F:(k + 1) * F], g["lens"][:, k * F:(k + 1) * F], 960) for k in range(ncalls)]
        want = [(a.copy(), b.copy(), c.copy()) for a, b, c in want]
        dec.reset()
        pk = np.ascontiguousarray(g["packets"].reshape(-1))
        bufs = [(np.zeros((S, F, 1920), np.float32), np.zeros((S, F), np.int32), np.zeros((S, F), np.uint32)) for _ in range(2)]
        ins = [(np.ascontiguousarray(offs[:, k * F:(k + 1) * F]), np.ascontiguousarray(g["lens"][:, k * F:(k + 1) * F])) for k in range(ncalls)]
        for k in range(ncalls):
            pcm, smp, rng = bufs[k & 1]
            dec.decode_float_multi_async(pk, ins[k][0], ins[k][1], 960, pcm, smp, rng)
            if k:
                dec.wait(1)
                p2, s2, r2 = bufs[(k - 1) & 1]
                assert np.array_equal(p2, want[k - 1][0]) and np.array_equal(s2, want[k - 1][1]) and np.array_equal(r2, want[k - 1][2])
        dec.wait(0)
        p2, s2, r2 = bufs[(ncalls - 1) & 1]
        assert np.array_equal(p2, want[-1][0]) and np.array_equal(r2, want[-1][2])
        assert (dec.final_range() == want[-1][2][:, -1]).all()


@pytest.mark.parametrize("name", ["cfg2_mono_20ms_64k_cbr", "stereo_20ms_vbr_96k", "cfg4_stereo_5ms_96k", "cfg4_mono_2p5ms_64k"])
def test_multiframe_packets(have_ref, name):
    """TOC codes 1, 2, 3 (CBR / VBR sizes, padding, a DTX frame inside a packet) and a lost 3-frame slot.  Property: a multi-frame
    packet decodes to the concatenation of its frames decoded as code-0 packets, final range = the last frame's; and, when the
    reference is on the box, everything equals the reference's decode of the same packets."""
    from opus_codec_b200.batch import BatchDecoder
    g = load_golden(name)
    fs, dc = g["frame_size"], g["dec_channels"]
    S = 3
    built = [multiframe_stream(g, s, 14) for s in range(S)]
    F = min(b[0].shape[0] for b in built)
    stride = max(b[0].shape[1] for b in built)
    pk = np.zeros((S, F, stride), np.uint8); ln = np.zeros((S, F), np.int32)
    for s, (p, l) in enumerate(built):
        pk[s, :, :p.shape[1]] = p[:F]; ln[s] = l[:F]
    ln[1, 4] = 0
    slot = 3 * fs
    with BatchDecoder(S, 48000, dc, device=0, max_frames=3 * F) as dec:
        pcm, samples, ranges = dec.decode_float_multi(pk.reshape(-1), _offsets(S, F, stride), ln, slot)
    pcm = pcm.reshape(S, F, slot * dc)
    # stream 0 holds no loss and no DTX up to packet 5: compare with the golden's frame-by-frame decode
    k = 0
    for f in range(6):
        n = samples[0, f] // fs
        assert samples[0, f] > 0 and samples[0, f] % fs == 0
        want = g["pcm"][0, k:k + n].reshape(-1)
        assert np.abs(pcm[0, f, :samples[0, f] * dc] - want).max() <= PCM_TOL
        assert ranges[0, f] == g["dec_rng"][0, k + n - 1]
        k += n
    if have_ref:
        from oracle import refpy
        for s in range(S):
            ref, rr, rs = refpy.decode_stream(pk[s], ln[s], slot, dc, pure_c=True)
            assert (samples[s] == rs).all() and (ranges[s] == rr).all()
            for f in range(F):
                tol = PCM_TOL if (s != 1 or f < 4) and f < 6 else 3e-3      # after a concealment: the reference's own spread (see the PLC test)
                assert np.abs(pcm[s, f, :rs[f] * dc] - ref[f, :rs[f] * dc]).max() <= tol, (s, f)


def test_multiframe_packets_with_too_few_frame_slots_report_buffer_too_small():
    """A decoder created with max_frames == the number of packets, fed multi-frame packets: every packet must still get a status -- the
    ones that fit decode, the ones whose frames would eat the slots of later packets are OPUS_BUFFER_TOO_SMALL (include/opus_b200.h), and
    nothing is left unwritten (ob_frame_packets keeps one slot in reserve for every packet still to come)."""
    from conftest import repacketize
    from opus_codec_b200.batch import BatchDecoder
    g = load_golden("cfg2_mono_20ms_64k_cbr")
    fs, dc = g["frame_size"], g["dec_channels"]
    pk0, ln0 = g["packets"][0], g["lens"][0]
    fr = [bytes(pk0[f, :ln0[f]]) for f in range(8)]
    for maxf, want in ((4, (2 * fs, -2, fs)), (3, (-2, -2, fs)), (5, (2 * fs, 2 * fs, fs))):
        pkts = [repacketize(fr[0:2], 1), repacketize(fr[2:4], 1), fr[4]]
        stride = max(len(p) for p in pkts)
        buf = np.zeros((1, 3, stride), np.uint8); ln = np.zeros((1, 3), np.int32)
        for i, p in enumerate(pkts):
            buf[0, i, :len(p)] = np.frombuffer(p, np.uint8); ln[0, i] = len(p)
        with BatchDecoder(1, 48000, dc, device=0, max_frames=maxf) as dec:
            pcm, samples, ranges = dec.decode_float_multi(buf.reshape(-1), _offsets(1, 3, stride), ln, 2 * fs)
        assert tuple(int(v) for v in samples[0]) == want, (maxf, samples)


def test_int16_api_soft_clip_matches_reference(have_ref):
    """ob_decode_multi (Decoder::decode): with +12 dB of decode gain the signal clips, so the soft clipper and its packet-to-packet
    state are exercised; int16 PCM must equal the reference's up to rounding flips: the float paths differ by up to ~0.2 LSB after
    the gain, so a +-1 on under 2 % of the samples is tolerated (the clipper and the rounding themselves are bit-exact against the
    reference in tests/test_host_emul.py)."""
    if not have_ref:
        pytest.skip("oracle/_ref not built")
    from oracle import refpy
    from opus_codec_b200.batch import BatchDecoder
    for name, q8 in (("cfg1_stereo_20ms_128k_cbr", 3072), ("cfg2_mono_20ms_64k_cbr", 0), ("cfg4_stereo_5ms_96k", 3328)):
        g = load_golden(name)
        fs, dc = g["frame_size"], g["dec_channels"]
        S, F, stride = g["packets"].shape
        S, F = min(S, 3), min(F, 40)
        pk = np.ascontiguousarray(g["packets"][:S, :F]); ln = g["lens"][:S, :F]
        with BatchDecoder(S, 48000, dc, device=0, max_frames=F) as dec:
            dec.set_gain(q8)
            pcm, samples, ranges = dec.decode_multi(pk.reshape(-1), _offsets(S, F, stride), ln, fs)
            assert pcm.dtype == np.int16
            clipped = 0
            for s in range(S):
                ref, rr, rs = refpy.decode_stream_i16(pk[s], ln[s], fs, dc, gain_q8=q8)
                assert (samples[s] == rs).all() and (ranges[s] == rr).all()
                d = np.abs(pcm[s].astype(np.int32) - ref.astype(np.int32))
                assert d.max() <= 1 and (d != 0).mean() < 2e-2, (name, int(d.max()), float((d != 0).mean()))
                clipped += int((np.abs(ref.astype(np.int32)) >= 32000).sum())
            assert q8 == 0 or clipped > 100          # the gain really drove the signal into the clipper
            # and the float API afterwards forgets the clipper state (opus_decoder.c:806)
            f32, _, _ = dec.decode_float_multi(pk.reshape(-1), _offsets(S, F, stride), ln, fs)
            assert np.isfinite(f32).all()


def test_int16_api_lost_packets_after_clipping_match_reference(have_ref):
    """A lost packet right after a clipped one, int16 API: libopus leaves opus_decode_native through its len == 0 branch (opus_decoder.c:714-729),
    i.e. BEFORE the soft clipper -- the concealment is only saturated and the clipper's carried state is not touched.  Concealed frames differ
    between the reference's own builds by ~2e-3 of full scale (see the PLC test), so they get that tolerance; the packets after the loss must
    again agree to +-1 LSB, which they only do if the clipper state was carried across the loss exactly as the reference carries it."""
    if not have_ref:
        pytest.skip("oracle/_ref not built")
    from oracle import refpy
    from opus_codec_b200.batch import BatchDecoder
    g = load_golden("cfg1_stereo_20ms_128k_cbr")
    fs, dc = g["frame_size"], g["dec_channels"]
    S, F, stride = g["packets"].shape
    S, F = min(S, 3), min(F, 40)
    pk = np.ascontiguousarray(g["packets"][:S, :F]); ln = g["lens"][:S, :F].copy()
    lost = (7, 8, 20, 31)
    ln[:, lost] = 0
    with BatchDecoder(S, 48000, dc, device=0, max_frames=F) as dec:
        dec.set_gain(3072)
        pcm, samples, ranges = dec.decode_multi(pk.reshape(-1), _offsets(S, F, stride), ln, fs)
    for s in range(S):
        ref, rr, rs = refpy.decode_stream_i16(pk[s], ln[s], fs, dc, gain_q8=3072)
        assert (samples[s] == rs).all() and (ranges[s] == rr).all()
        d = np.abs(pcm[s].astype(np.int32) - ref.astype(np.int32)).reshape(F, -1)
        for f in range(F):
            near_loss = any(0 <= f - l <= 2 for l in lost)
            assert d[f].max() <= (int(3e-3 * 32768 * 4) if near_loss else 1), (s, f, int(d[f].max()))


@pytest.mark.parametrize("fs_out", [24000, 16000, 12000, 8000])
def test_output_sample_rates_below_48k(have_ref, fs_out):
    """Decoder::new(SampleRate::Hz8000..Hz24000): the spectrum is cut at the new Nyquist and every ds-th de-emphasised sample is
    kept (celt_decoder.c:326-373, bands.c:206-208); frame sizes, sample counts and lost-packet slots count OUTPUT samples."""
    if not have_ref:
        pytest.skip("oracle/_ref not built")
    from oracle import refpy
    from opus_codec_b200.batch import BatchDecoder
    ds = 48000 // fs_out
    for name in ("cfg1_stereo_20ms_128k_cbr", "cfg4_mono_5ms_48k"):
        g = load_golden(name)
        fs, dc = g["frame_size"] // ds, g["dec_channels"]
        S, F, stride = g["packets"].shape
        S, F = min(S, 3), min(F, 30)
        pk = np.ascontiguousarray(g["packets"][:S, :F]); ln = g["lens"][:S, :F].copy()
        ln[1, 7] = 0; ln[2, 3] = 1
        with BatchDecoder(S, fs_out, dc, device=0, max_frames=F) as dec:
            pcm, samples, ranges = dec.decode_float_multi(pk.reshape(-1), _offsets(S, F, stride), ln, fs)
        pcm = pcm.reshape(S, F, fs * dc)
        for s in range(S):
            ref, rr, rs = refpy.decode_stream(pk[s], ln[s], fs, dc, pure_c=True, fs=fs_out)
            assert (samples[s] == rs).all() and (ranges[s] == rr).all(), (samples[s], rs)
            lossy = s != 0
            for f in range(F):
                tol = 3e-3 if lossy and f >= 3 else PCM_TOL
                assert np.abs(pcm[s, f] - ref[f]).max() <= tol, (name, s, f)


def test_reset_restarts_streams():
    from opus_codec_b200.batch import BatchDecoder
    g = load_golden("cfg2_mono_20ms_64k_cbr")
    S, F, stride = g["packets"].shape
    F = 12
    with BatchDecoder(S, 48000, 1, device=0, max_frames=F) as dec:
        a, _, _ = dec.decode_float_multi(g["packets"].reshape(-1), _offsets(S, g["packets"].shape[1], stride)[:, :F], g["lens"][:, :F], 960)
        a = a.copy()
        dec.reset([0, 2])
        b, _, rb = dec.decode_float_multi(g["packets"].reshape(-1), _offsets(S, g["packets"].shape[1], stride)[:, :F], g["lens"][:, :F], 960)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[2], b[2])      # reset streams replay identically
    assert not np.array_equal(a[1], b[1])                                   # the others carried their state on
    assert (rb == g["dec_rng"][:, :F]).all()                               # final range never depends on state


def test_opus_compare_passes():
    """The reference's own conformance metric (opus/src/opus_compare.c) on GPU output vs reference decoder output."""
    from oracle import refpy
    if not os.path.exists(refpy.OPUS_COMPARE):
        pytest.skip("oracle/_ref/opus_compare not built")
    for name, ch in (("cfg2_mono_20ms_64k_cbr", 1), ("cfg1_stereo_20ms_128k_cbr", 2), ("cfg4_stereo_5ms_96k", 2)):
        g = load_golden(name)
        pcm, _, _ = _decode_all(g, streams=[0])
        with tempfile.TemporaryDirectory() as td:
            ok, txt = refpy.opus_compare(g["pcm"][0].reshape(-1), pcm[0].reshape(-1), ch, td)
        assert ok, txt


def test_full_size_batch_4096_streams_tiling_property():
    """BASELINE config 2 at full size: 4096 mono 20 ms 64 kb/s streams.  The batch tiles the 9 golden streams, so every
    tile must reproduce the golden final ranges bit-exactly and identical PCM across tiles (independence of streams)."""
    from opus_codec_b200.batch import BatchDecoder
    g = load_golden("cfg2_mono_20ms_64k_cbr")
    P, F, stride = g["packets"].shape
    F = 25
    S = 4096
    idx = np.arange(S) % P
    pk = np.ascontiguousarray(g["packets"][idx, :F])
    ln = np.ascontiguousarray(g["lens"][idx, :F])
    with BatchDecoder(S, 48000, 1, device=0, max_frames=F) as dec:
        pcm, samples, ranges = dec.decode_float_multi(pk.reshape(-1), _offsets(S, F, stride), ln, 960)
        assert dec.launches() >= 3
    assert (samples == 960).all()
    assert (ranges == g["dec_rng"][idx, :F]).all()
    assert np.abs(pcm[0] - g["pcm"][0, :F]).max() <= PCM_TOL
    for s in range(P, S):
        assert np.array_equal(pcm[s], pcm[s % P])


def test_decode_fec_flag_conceals_celt_packets_like_the_reference(have_ref):
    """Decoder::decode(.., fec = true): a CELT-only packet carries no FEC, so libopus conceals the frame as if the packet were lost -- after
    parsing it (a malformed packet is still an error).  Checked both ways: the reference gives the same PCM for `fec` and for a lost packet,
    and so does the GPU; the GPU's PCM equals the reference's within the concealment tolerance."""
    if not have_ref:
        pytest.skip("oracle/_ref not built")
    from oracle import refpy
    from opus_codec_b200.batch import BatchDecoder, pack_packets
    g = load_golden("cfg1_stereo_20ms_128k_cbr")
    ch, fs, F, S = g["channels"], g["frame_size"], 30, 3
    fec_at = {7, 8, 15, 22}
    pk = [[bytes(g["packets"][s][f, :g["lens"][s][f]]) for f in range(F)] for s in range(S)]
    stride = g["packets"].shape[2]
    for s in range(S):
        flags = np.array([int(f in fec_at) for f in range(F)], np.int32)
        lens_lost = g["lens"][s][:F].copy(); lens_lost[list(fec_at)] = 0
        a, ra, sa = refpy.decode_stream(g["packets"][s][:F], g["lens"][s][:F], fs, ch, pure_c=True, fec_flags=flags)
        b, rb, sb = refpy.decode_stream(g["packets"][s][:F], lens_lost, fs, ch, pure_c=True)
        assert np.array_equal(a, b) and (ra == rb).all() and (sa == sb).all()
    out = np.zeros((S, F, fs * ch), np.float32); rng = np.zeros((S, F), np.uint32)
    out2 = np.zeros_like(out)
    with BatchDecoder(S, 48000, ch, device=0, max_frames=1) as dec, BatchDecoder(S, 48000, ch, device=0, max_frames=1) as dec2:
        for f in range(F):
            dec.set_decode_fec(f in fec_at)
            bb, oo, ll = pack_packets([[pk[s][f]] for s in range(S)])
            p, smp, r = dec.decode_float_multi(bb, oo, ll, fs)
            assert (smp == fs).all()
            out[:, f] = p[:, 0]; rng[:, f] = r[:, 0]
            bb, oo, ll = pack_packets([[b"" if f in fec_at else pk[s][f]] for s in range(S)])
            out2[:, f] = dec2.decode_float_multi(bb, oo, ll, fs)[0][:, 0]
        # errors: a malformed packet is reported even with fec on; the frame size must be a multiple of 2.5 ms
        dec.set_decode_fec(True)
        bad = bytes([pk[0][0][0] | 1]) + pk[0][0][1:4]                       # code 1 with an odd payload
        bb, oo, ll = pack_packets([[bad], [pk[1][0]], [pk[2][0]]])
        p, smp, r = dec.decode_float_multi(bb, oo, ll, fs)
        assert smp[0, 0] == -4 and (smp[1:, 0] == fs).all() and (r[1:, 0] == 0).all()
        p, smp, r = dec.decode_float_multi(bb, oo, ll, 1000)
        assert (smp[:, 0] == -1).all()
    assert np.array_equal(out, out2)
    for s in range(S):
        flags = np.array([int(f in fec_at) for f in range(F)], np.int32)
        a, ra, _ = refpy.decode_stream(g["packets"][s][:F], g["lens"][s][:F], fs, ch, pure_c=True, fec_flags=flags)
        assert (ra == rng[s]).all()
        assert np.abs(a - out[s]).max() <= 3e-3
        first = min(fec_at)
        assert np.abs(a[:first] - out[s][:first]).max() <= 1e-4


def test_get_pitch_matches_reference(have_ref):
    """OPUS_GET_PITCH after every packet (post-filter period of the last decoded frame; unchanged by a concealed one; 0 on a fresh decoder)."""
    if not have_ref:
        pytest.skip("oracle/_ref not built")
    import ctypes as C
    from oracle import refpy
    from opus_codec_b200.batch import BatchDecoder, pack_packets
    g = load_golden("cfg2_mono_20ms_64k_cbr")
    ch, fs, F, S = g["channels"], g["frame_size"], 40, 4
    L = refpy.lib()
    want = np.zeros((S, F), np.int32)
    for s in range(S):
        lens = g["lens"][s][:F].copy(); lens[[9, 10, 25]] = 0
        buf = np.zeros(F, np.int32)
        L.ref_set_decoder_pitch_out(buf.ctypes.data_as(C.POINTER(C.c_int)))
        try:
            refpy.decode_stream(g["packets"][s][:F], lens, fs, ch)
        finally:
            L.ref_set_decoder_pitch_out(None)
        want[s] = buf
    assert (want > 0).any()
    with BatchDecoder(S, 48000, ch, device=0, max_frames=1) as dec:
        assert (dec.pitch() == 0).all()
        for f in range(F):
            bb, oo, ll = pack_packets([[b"" if f in (9, 10, 25) else bytes(g["packets"][s][f, :g["lens"][s][f]])] for s in range(S)])
            dec.decode_float_multi(bb, oo, ll, fs)
            assert (dec.pitch() == want[:, f]).all(), f


def test_mono_decoder_with_mono_and_stereo_packets_across_windows(have_ref):
    """A mono decoder fed streams that switch between mono and stereo packets (legal: the decoder down-mixes), in a call large enough to run
    in several frame windows: some windows are all-mono (mono-sized band / synthesis kernels), others hold stereo frames (straggler pass +
    the two-channel synthesis kernel), and the stream state must carry across the variants.  Checked against the reference on a sample."""
    if not have_ref:
        pytest.skip("oracle/_ref not built")
    from oracle import refpy
    from opus_codec_b200.batch import BatchDecoder
    gm, gs = load_golden("cfg2_mono_20ms_64k_cbr"), load_golden("cfg3_stereo_20ms_96k_cbr")
    S, F = 1400, 48                                             # 67 200 frames: three windows of 16 frames
    stride = max(gm["packets"].shape[2], gs["packets"].shape[2])
    pk = np.zeros((S, F, stride), np.uint8); ln = np.zeros((S, F), np.int32)
    rng = np.random.default_rng(3)
    for s in range(S):
        a, b = s % gm["packets"].shape[0], s % gs["packets"].shape[0]
        kind = s % 4                                            # 0: all mono, 1: stereo in the middle window only, 2: random blocks, 3: stereo with a lost packet
        for f in range(F):
            stereo = (kind == 1 and 16 <= f < 32) or (kind == 2 and rng.random() < 0.3) or kind == 3
            g, i = (gs, b) if stereo else (gm, a)
            n = g["lens"][i][f]
            pk[s, f, :n] = g["packets"][i][f, :n]; ln[s, f] = n
        if kind == 3:
            ln[s, 20] = 0
    with BatchDecoder(S, 48000, 1, device=0, max_frames=F) as dec:
        pcm, smp, ranges = dec.decode_float_multi(pk.reshape(-1), _offsets(S, F, stride), ln, 960)
    assert (smp == 960).all()
    for s in list(range(8)) + [S - 4, S - 3, S - 2, S - 1]:
        ref_pcm, ref_rng, ref_smp = refpy.decode_stream(pk[s], ln[s], 960, 1, pure_c=True)
        assert (ref_rng == ranges[s]).all(), s
        tol = 3e-3 if s % 4 == 3 else PCM_TOL                   # concealment tolerance for the stream with a lost packet
        assert np.abs(ref_pcm - pcm[s]).max() <= tol, s


def test_single_frame_call_split_over_two_compute_streams_with_stereo_frames():
    """One frame per stream, 20 000 mono-decoder streams: the call is split by stream ranges over two CUDA streams, each with its own
    stereo-frame counter and straggler list.  Every stream must decode exactly as it does in a small batch."""
    from opus_codec_b200.batch import BatchDecoder
    gm, gs = load_golden("cfg2_mono_20ms_64k_cbr"), load_golden("cfg3_stereo_20ms_96k_cbr")
    S, F = 20000, 3
    stride = max(gm["packets"].shape[2], gs["packets"].shape[2])
    pk = np.zeros((S, F, stride), np.uint8); ln = np.zeros((S, F), np.int32)
    for s in range(S):
        for f in range(F):
            g = gs if (s * 7 + f) % 5 == 0 else gm
            i = s % g["packets"].shape[0]
            n = g["lens"][i][f]
            pk[s, f, :n] = g["packets"][i][f, :n]; ln[s, f] = n
    out = np.zeros((S, F, 960), np.float32); rngs = np.zeros((S, F), np.uint32)
    with BatchDecoder(S, 48000, 1, device=0, max_frames=1) as dec:
        for f in range(F):                                      # one frame per call: the stream-range split
            p, smp, r = dec.decode_float_multi(np.ascontiguousarray(pk[:, f]).reshape(-1), _offsets(S, 1, stride), ln[:, f:f + 1], 960)
            assert (smp == 960).all()
            out[:, f] = p[:, 0]; rngs[:, f] = r[:, 0]
    sub = np.r_[0:40, 9990:10030, S - 40:S]
    with BatchDecoder(len(sub), 48000, 1, device=0, max_frames=F) as dec:
        p, smp, r = dec.decode_float_multi(np.ascontiguousarray(pk[sub]).reshape(-1), _offsets(len(sub), F, stride), ln[sub], 960)
    assert np.array_equal(p, out[sub]) and (r == rngs[sub]).all()


@pytest.mark.parametrize("channels,F", [(2, 8), (1, 2)])
def test_chunking_of_a_host_call_never_changes_its_output(channels, F):
    """How ob_decode_float_multi cuts a call into chunks (frame windows with 2-D copies, stream ranges with contiguous copies, up to 16 stream
    ranges for live calls; opus_b200.cu ob_decode_submit) is a pipelining decision: PCM, sample counts and final ranges of a large batch must
    equal, bit for bit, what the same streams give in a small batch that is not chunked at all -- over two consecutive calls (carried state)."""
    from opus_codec_b200.batch import BatchDecoder
    g = load_golden("cfg3_stereo_20ms_96k_cbr" if channels == 2 else "cfg2_mono_20ms_64k_cbr")
    S = 17000 if channels == 2 else 70000           # 136 000 stereo frames per call: stream-range chunks (rows of a frame window would be 15 KB); 140 000 mono frames at F = 2: 16 stream ranges
    P, stride, N = g["packets"].shape[0], g["packets"].shape[2], g["frame_size"]
    idx = (np.arange(S) * 5) % P
    calls = [(np.ascontiguousarray(g["packets"][idx, c * F:(c + 1) * F]), np.ascontiguousarray(g["lens"][idx, c * F:(c + 1) * F]).astype(np.int32)) for c in range(2)]
    sub = np.r_[0:24, S // 2 - 12:S // 2 + 12, S - 24:S]
    big, small = [], []
    with BatchDecoder(S, 48000, channels, device=0, max_frames=F) as dec:
        for pk, ln in calls:
            p, smp, r = dec.decode_float_multi(pk.reshape(-1), _offsets(S, F, stride), ln, N)
            assert (smp == N).all()
            big.append((p[sub].copy(), r[sub].copy()))
    with BatchDecoder(len(sub), 48000, channels, device=0, max_frames=F) as dec:
        for pk, ln in calls:
            p, smp, r = dec.decode_float_multi(np.ascontiguousarray(pk[sub]).reshape(-1), _offsets(len(sub), F, stride), ln[sub], N)
            small.append((p, r))
    for (pb, rb), (ps, rs) in zip(big, small):
        assert (rb == rs).all() and np.array_equal(pb, ps)
    assert (small[1][1] == g["dec_rng"][idx[sub], F:2 * F]).all()              # and the reference's final ranges


class _IrHdr(__import__("ctypes").Structure):
    import ctypes as _C
    _fields_ = [("status", _C.c_int32), ("final_range", _C.c_uint32), ("n_leaves", _C.c_uint16), ("pf_pitch", _C.c_uint16),
                ("LM", _C.c_uint8), ("C", _C.c_uint8), ("end", _C.c_uint8), ("flags", _C.c_uint8),
                ("spread", _C.c_uint8), ("pf_tapset", _C.c_uint8), ("pf_qg", _C.c_uint8), ("coded_bands", _C.c_uint8),
                ("intensity", _C.c_uint8), ("dual_stereo", _C.c_uint8), ("skip_in", _C.c_uint8), ("end_in", _C.c_uint8),
                ("lcg_total", _C.c_uint32), ("seed_in", _C.c_uint32), ("loss_in", _C.c_int32), ("lastfs_in", _C.c_uint16), ("pad2", _C.c_uint16),
                ("coarse_qi", _C.c_int16 * 42), ("pulses", _C.c_int16 * 21), ("fine_quant", _C.c_uint8 * 21), ("fine_q2", _C.c_uint8 * 42),
                ("final_bit", _C.c_int8 * 42), ("collapse_masks", _C.c_uint8 * 42), ("tf_change", _C.c_int8 * 21), ("pad1", _C.c_uint8 * 3)]


@pytest.mark.parametrize("name", golden_names())
def test_integer_ir_on_the_device_matches_oracle_and_reference_taps(have_ref, name):
    """BASELINE north_star: 'decoded energy indices and pulse vectors must match exactly'.  The symbol kernel's integer record of every frame
    (csrc/ob_ir.h) is read back from the device (ob_decoder_debug_read_ir) and compared field by field with the oracle's taps -- coarse energy
    indices, tf, PVQ bit allocation, fine-energy bits, collapse masks, spread / intensity / dual stereo / coded bands, post-filter parameters,
    and every decoded pulse vector iy[] at its position -- and, where the compiled reference is present, with the values the REFERENCE
    itself passed to quant_all_bands (the --wrap taps of oracle/ref_shim.c)."""
    import ctypes as C
    from opus_codec_b200 import _lib
    from opus_codec_b200.batch import BatchDecoder
    from oracle import oraclepy
    L = _lib.lib()
    lay = (C.c_int32 * 8)()
    assert L.ob_debug_ir_layout(lay, 8) == 0
    ir_size, hdr_size, _, _, iy_off = lay[0], lay[1], lay[2], lay[3], lay[4]
    assert hdr_size == C.sizeof(_IrHdr)
    g = load_golden(name)
    fs, dc = g["frame_size"], g["dec_channels"]
    S, F, stride = g["packets"].shape
    S, F = min(S, 2), min(F, 30)
    pk = np.ascontiguousarray(g["packets"][:S, :F]); ln = np.ascontiguousarray(g["lens"][:S, :F])
    buf = C.create_string_buffer(ir_size)
    with BatchDecoder(S, 48000, dc, device=0, max_frames=F) as dec:
        dec.decode_float_multi(pk.reshape(-1), _offsets(S, F, stride), ln, fs)
        for s in range(S):
            _, orng, _, taps = oraclepy.decode_stream(pk[s], ln[s], fs, dc, want_taps=True)
            rtaps = None
            if have_ref:
                from oracle import refpy
                rtaps = refpy.decode_stream(pk[s], ln[s], fs, dc, want_taps=True)[3]
            for f in range(F):
                assert L.ob_decoder_debug_read_ir(dec._h, s, f, buf, ir_size) == 0
                h = _IrHdr.from_buffer_copy(buf.raw[:hdr_size]); t = taps[f]
                C_, end = h.C, h.end
                assert h.status == fs and h.final_range == int(orng[f]) and h.LM == t.LM and C_ == t.C and end == t.end
                assert [h.coarse_qi[c * 21 + i] for c in range(C_) for i in range(end)] == [t.coarse_qi[c * 21 + i] for c in range(C_) for i in range(end)]
                assert list(h.pulses)[:end] == list(t.pulses)[:end] and list(h.fine_quant)[:end] == list(t.fine_quant)[:end]
                assert list(h.tf_change)[:end] == list(t.tf_res)[:end]
                assert list(h.collapse_masks)[:C_ * end] == list(t.collapse_masks)[:C_ * end]
                assert (h.spread, h.intensity, h.dual_stereo, h.coded_bands) == (t.spread, t.intensity, t.dual_stereo, t.coded_bands)
                assert bool(h.flags & 8) == bool(t.pf_on) and (not t.pf_on or (h.pf_pitch, h.pf_tapset, h.pf_qg) == (t.pf_pitch, t.pf_tapset, t.pf_qg))
                assert bool(h.flags & 2) == bool(t.transient) and bool(h.flags & 1) == bool(t.silence)
                iy = np.frombuffer(buf.raw, np.int16, 1920, iy_off)
                want = np.frombuffer(bytes(t.iy), np.int16); mask = np.frombuffer(bytes(t.iy_set), np.uint8) != 0
                assert mask.any() or t.silence or ln[s, f] < 12
                assert np.array_equal(iy[mask], want[mask]), (name, s, f)
                if rtaps is not None and rtaps[f].n_qab == 1:
                    r = rtaps[f]
                    assert list(h.pulses)[:end] == list(r.pulses)[:end] and list(h.tf_change)[:end] == list(r.tf_res)[:end]
                    assert list(h.collapse_masks)[:C_ * end] == list(r.collapse_masks)[:C_ * end]
                    assert (h.spread, h.intensity, h.dual_stereo, h.coded_bands, h.seed_in) == (r.spread, r.intensity, r.dual_stereo, r.codedBands, r.seed_in)
