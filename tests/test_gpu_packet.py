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
