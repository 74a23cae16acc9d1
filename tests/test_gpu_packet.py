"""The batched repacketizer kernel (ob_repacketize_batch) against the reference library's opus_repacketizer_* / opus_packet_pad, and
the merged packets through the batched decoder: the PCM of a merged packet equals the PCM of its parts, bit for bit."""
import numpy as np
import pytest

from conftest import load_golden
from test_packet import ext_padding, ref, ref_merge, with_padding, _buf      # noqa: F401  (ref is a fixture)

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name,group", [("cfg2_mono_20ms_64k_cbr", 3), ("stereo_20ms_vbr_96k", 2), ("cfg4_stereo_5ms_96k", 24), ("cfg4_mono_2p5ms_64k", 48),
                                        ("stereo_20ms_510k_highrate", 6), ("cfg4_stereo_10ms_96k", 5)])
def test_batch_merge_matches_reference_and_decodes_to_the_same_pcm(ref, name, group):
    from opus_codec_b200.batch import BatchDecoder, pack_packets
    from opus_codec_b200.packet import repacketize_batch
    g = load_golden(name)
    S = min(4, g["packets"].shape[0])
    F = (min(g["packets"].shape[1], 4 * group + 1))                                   # the last group is a short one
    pk = [[bytes(g["packets"][s][f, :g["lens"][s][f]]) for f in range(F)] for s in range(S)]
    out, lens = repacketize_batch(pk, group)
    G = (F + group - 1) // group
    assert out.shape[:2] == (S, G)
    merged = []
    for s in range(S):
        row = []
        for k in range(G):
            want = ref_merge(ref, pk[s][k * group:(k + 1) * group])
            assert isinstance(want, bytes) and lens[s, k] == len(want) and bytes(out[s, k, :lens[s, k]]) == want, (s, k)
            assert not out[s, k, lens[s, k]:].any()
            row.append(want)
        merged.append(row)
    fs, ch = g["frame_size"], g["channels"]
    with BatchDecoder(S, 48000, ch, device=0, max_frames=F) as dec:
        b, o, l = pack_packets(pk)
        pcm0, smp0, rng0 = dec.decode_float_multi(b, o, l, fs)
    with BatchDecoder(S, 48000, ch, device=0, max_frames=G * group) as dec:
        b, o, l = pack_packets(merged)
        pcm1, smp1, rng1 = dec.decode_float_multi(b, o, l, fs * group)
    for s in range(S):
        flat = np.concatenate([pcm1[s, k, :smp1[s, k] * ch] for k in range(G)])
        assert smp1[s].sum() == F * fs and np.array_equal(flat, pcm0[s].reshape(-1))
        assert (rng1[s] == rng0[s][[min(F, (k + 1) * group) - 1 for k in range(G)]]).all()


def test_batch_padding_extensions_and_errors(ref):
    from opus_codec_b200.packet import repacketize_batch
    g = load_golden("cfg3_stereo_20ms_96k_cbr")
    g2 = load_golden("cfg2_mono_20ms_64k_cbr")
    fr = [bytes(g["packets"][0][f, :g["lens"][0][f]]) for f in range(8)]
    other = bytes(g2["packets"][0][0, :g2["lens"][0][0]])
    ext1, ext3 = with_padding(fr[:1], ext_padding()), with_padding(fr[1:4], ext_padding())
    rows = [[fr[0], fr[1]], [ext1, fr[2]], [fr[3], ext1], [ext3, ext1], [fr[0], other], [fr[0], b""], [ext3, ext3], [fr[4][:1], fr[5]]]
    out, lens = repacketize_batch(rows, 2, max_bytes=2000)
    for s, row in enumerate(rows):
        want = ref_merge(ref, row) if all(len(p) for p in row) else -4
        if isinstance(want, bytes):
            assert lens[s, 0] == len(want) and bytes(out[s, 0, :len(want)]) == want, s
        else:
            assert lens[s, 0] == want, s
    assert lens[4, 0] == -4 and lens[5, 0] == -4 and lens[6, 0] > 0             # configuration change, lost packet; 3 + 3 frames of 20 ms = 120 ms is legal
    four = with_padding(fr[:4], b"")
    assert repacketize_batch([[four, ext3]], 2, max_bytes=4000)[1][0, 0] == ref_merge(ref, [four, ext3]) == -4      # 140 ms
    # pad_to: opus_packet_pad of the merged packet (CBR transport); too small a target is OPUS_BUFFER_TOO_SMALL
    out, lens = repacketize_batch(rows[:4], 2, pad_to=1500)
    for s, row in enumerate(rows[:4]):
        m = ref_merge(ref, row)
        d = _buf(m, 1500)
        assert ref.opus_packet_pad(d, len(m), 1500) == 0
        assert lens[s, 0] == 1500 and bytes(out[s, 0]) == bytes(d), s
    out, lens = repacketize_batch(rows[:1], 2, pad_to=100)
    assert lens[0, 0] == -2


def test_reference_tool_round_trips_through_the_bit_container(have_ref):
    """oracle/_ref/opus_demo (the reference's own tool) encodes -> .bit -> GPU int16 decode == the tool's own decode; GPU encode -> .bit ->
    the tool decodes it and its per-packet final-range check passes."""
    import os, subprocess, tempfile
    from conftest import ROOT
    from opus_codec_b200 import containers, synth
    from opus_codec_b200.batch import BatchDecoder, BatchEncoder, pack_packets
    demo = os.path.join(ROOT, "oracle", "_ref", "opus_demo")
    if not os.path.exists(demo):
        pytest.skip("oracle/_ref/opus_demo not built")
    ch, fs, F = 2, 960, 50
    pcm = synth.stream_pcm(5, fs * F, ch)
    pcm16 = np.clip(np.rint(pcm * 32768), -32768, 32767).astype("<i2")
    with tempfile.TemporaryDirectory() as d:
        pcm16.tofile(os.path.join(d, "in.pcm"))
        r = subprocess.run([demo, "-e", "restricted-lowdelay", "48000", "2", "96000", "-cbr", os.path.join(d, "in.pcm"), os.path.join(d, "ref.bit")], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        r = subprocess.run([demo, "-d", "48000", "2", os.path.join(d, "ref.bit"), os.path.join(d, "ref.pcm")], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        want = np.fromfile(os.path.join(d, "ref.pcm"), "<i2")
        pk, rng = containers.read_bit(open(os.path.join(d, "ref.bit"), "rb").read())
        with BatchDecoder(1, 48000, ch, device=0, max_frames=len(pk)) as dec:
            b, o, l = pack_packets([pk])
            got, smp, drng = dec.decode_multi(b, o, l, fs)
        assert (smp == fs).all() and (drng[0] == np.array(rng, np.uint32)).all()
        got = got.reshape(-1)
        assert got.size == want.size
        diff = np.abs(got.astype(np.int32) - want.astype(np.int32))
        assert diff.max() <= 1 and (diff != 0).mean() < 0.02
        # GPU encode -> the tool decodes and checks every final range
        with BatchEncoder(1, 48000, ch, device=0, max_frames=F) as enc:
            enc.set_bitrate(96000); enc.set_vbr(False); enc.set_complexity(10)
            out, lens, erng = enc.encode_multi(pcm16.reshape(1, F, fs * ch), fs)
        ours = [bytes(out[0, f, :lens[0, f]]) for f in range(F)]
        open(os.path.join(d, "gpu.bit"), "wb").write(containers.write_bit(ours, erng[0]))
        r = subprocess.run([demo, "-d", "48000", "2", os.path.join(d, "gpu.bit"), os.path.join(d, "gpu.pcm")], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        assert ours == pk[:F]                                   # and the GPU's packets are the tool's packets


def test_ogg_and_rtp_feed_the_batch_decoder():
    """An .opus file written from golden packets with an output gain, and the same packets as RTP with two of them lost: header gain ->
    set_gain, lost packets -> concealment; results equal the plain packet-list path."""
    from opus_codec_b200 import containers as c
    from opus_codec_b200.batch import BatchDecoder, pack_packets
    g = load_golden("cfg2_mono_20ms_64k_cbr")
    pk = [bytes(g["packets"][0][f, :g["lens"][0][f]]) for f in range(40)]
    o = c.ogg_read(c.ogg_write(pk, 1, output_gain_q8=-512))
    lossy, _ = c.rtp_to_packets([c.rtp_pack(p, i, 960 * i, 9) for i, p in enumerate(pk) if i not in (11, 12)])
    assert lossy == [b"" if i in (11, 12) else p for i, p in enumerate(pk)]
    res = []
    for packets, gain in ((o["packets"], o["output_gain_q8"]), (pk, -512), (lossy, 0), ([b"" if i in (11, 12) else p for i, p in enumerate(pk)], 0)):
        with BatchDecoder(1, 48000, 1, device=0, max_frames=40) as dec:
            dec.set_gain(gain)
            b, of, l = pack_packets([packets])
            res.append(dec.decode_float_multi(b, of, l, 960))
    assert np.array_equal(res[0][0], res[1][0]) and (res[0][2] == res[1][2]).all()
    assert np.array_equal(res[2][0], res[3][0]) and (res[2][1] == 960).all()
    assert np.abs(res[0][0]).max() > 0 and not np.array_equal(res[0][0], res[2][0])


def test_soft_clip_batch_matches_reference(have_ref):
    """ob_pcm_soft_clip_batch == opus_pcm_soft_clip of the reference's C build, sample for sample, with the carried state across calls."""
    if not have_ref:
        pytest.skip("oracle/_ref not built")
    import ctypes as C
    from oracle import refpy
    from opus_codec_b200.packet import soft_clip_batch
    L = refpy.lib_c()
    rng = np.random.default_rng(77)
    for ch in (1, 2):
        S, n = 9, 480
        mem = np.zeros((S, ch), np.float32); rmem = mem.copy()
        for call in range(4):
            t = np.arange(n * ch, dtype=np.float32).reshape(1, -1)
            amp = rng.uniform(0.2, 3.5, (S, 1)).astype(np.float32)
            x = (amp * np.sin(t * rng.uniform(0.01, 0.3, (S, 1)) + call) + rng.normal(0, 0.2, (S, n * ch))).astype(np.float32)
            x[0] *= 0.1                                           # never clips
            want = x.copy()
            for s in range(S):
                L.opus_pcm_soft_clip(want[s].ctypes.data_as(C.POINTER(C.c_float)), n, ch, rmem[s].ctypes.data_as(C.POINTER(C.c_float)))
            got = x.copy()
            soft_clip_batch(got, ch, mem)
            assert np.array_equal(got, want) and np.array_equal(mem, rmem), (ch, call)
            assert np.abs(got).max() <= 1.0 and np.abs(x).max() > 1.5
