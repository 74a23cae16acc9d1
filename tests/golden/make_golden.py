"""Generates tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref/libopus_ref.so).

Run in the build container (needs oracle/_ref, which is compiled from /root/reference/opus):
    python tests/golden/make_golden.py
Each file holds, for one configuration of BASELINE.json / SURVEY.md section 8d:
    packets u8 [streams, frames, stride]   reference-ENCODED packets (TOC included)
    lens    i32 [streams, frames]
    enc_rng u32 [streams, frames]          OPUS_GET_FINAL_RANGE of the reference encoder
    dec_rng u32 [streams, frames]          OPUS_GET_FINAL_RANGE of the reference decoder
    pcm     f32 [npcm, frames, fs*ch]      reference-decoded PCM of the first npcm streams
    meta    (channels, frame_size, bitrate, vbr, dec_channels)
"""
import os
import sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from opus_codec_b200 import synth          # noqa: E402
from oracle import refpy                   # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

# name: (channels, frame_size, bitrate, vbr, streams, frames, npcm, extras)
CONFIGS = {
    "cfg2_mono_20ms_64k_cbr":     (1, 960, 64000, 0, 9, 50, 1, {}),
    "cfg1_stereo_20ms_128k_cbr":  (2, 960, 128000, 0, 6, 50, 1, {}),
    "cfg3_stereo_20ms_96k_cbr":   (2, 960, 96000, 0, 6, 50, 1, {}),
    "stereo_20ms_vbr_96k":        (2, 960, 96000, 1, 3, 50, 1, {}),
    "cfg4_stereo_10ms_96k":       (2, 480, 96000, 0, 3, 100, 1, {}),
    "cfg4_stereo_5ms_96k":        (2, 240, 96000, 0, 3, 200, 1, {}),
    "cfg4_stereo_2p5ms_96k":      (2, 120, 96000, 0, 3, 400, 1, {}),
    "cfg4_mono_10ms_48k":         (1, 480, 48000, 0, 3, 100, 1, {}),
    "cfg4_mono_5ms_48k":          (1, 240, 48000, 0, 3, 200, 1, {}),
    "cfg4_mono_2p5ms_64k":        (1, 120, 64000, 0, 3, 400, 1, {}),
    "mono_20ms_12k_lowrate":      (1, 960, 12000, 0, 3, 50, 1, {}),
    "stereo_20ms_24k_lowrate":    (2, 960, 24000, 0, 3, 50, 1, {}),
    "mono_20ms_256k_highrate":    (1, 960, 256000, 0, 3, 50, 1, {}),
    "stereo_20ms_510k_highrate":  (2, 960, 510000, 0, 3, 50, 1, {}),
    "mono_20ms_wb":               (1, 960, 32000, 0, 3, 50, 1, {"bandwidth": 1103}),
    "stereo_10ms_swb":            (2, 480, 64000, 0, 3, 50, 1, {"bandwidth": 1104}),
    "mono_20ms_nb":               (1, 960, 24000, 0, 3, 50, 1, {"bandwidth": 1101}),
    "stereo_dec_of_mono_pkts":    (1, 960, 64000, 0, 3, 50, 1, {"dec_channels": 2}),
    "mono_dec_of_stereo_pkts":    (2, 960, 96000, 0, 3, 50, 1, {"dec_channels": 1}),
}
# Packet pools for bench.py ("throughput sets reuse a pool of 256 distinct streams tiled to the batch size").
# No PCM is stored; final ranges are, so the bench can verify every decoded frame.
POOLS = {
    "pool_cfg2_mono_20ms_64k_cbr": (1, 960, 64000, 0, 256, 50, 0, {}),
}


def make(name):
    ch, fs, br, vbr, ns, nf, npcm, ex = (CONFIGS.get(name) or POOLS[name])
    dec_ch = ex.get("dec_channels", ch)
    pk_all, ln_all, er_all, dr_all, pcm_all = [], [], [], [], []
    stride = 0
    for s in range(ns):
        pcm = synth.stream_pcm(s, nf * fs, ch, base_seed=4242)
        pk, ln, er = refpy.encode_stream(pcm, fs, ch, br, vbr=vbr, bandwidth=ex.get("bandwidth", 0))
        out, dr, smp = refpy.decode_stream(pk, ln, fs, dec_ch)
        assert (smp == fs).all()
        if dec_ch == ch:
            assert (er == dr).all(), "reference enc/dec final range mismatch?!"
        pk_all.append(pk); ln_all.append(ln); er_all.append(er); dr_all.append(dr)
        if s < npcm:
            pcm_all.append(out)
        stride = max(stride, int(ln.max()))
    packets = np.stack([p[:, :stride] for p in pk_all])
    np.savez_compressed(os.path.join(HERE, name + ".npz"), packets=packets, lens=np.stack(ln_all),
                        enc_rng=np.stack(er_all), dec_rng=np.stack(dr_all),
                        pcm=np.stack(pcm_all) if pcm_all else np.zeros((0, nf, fs * dec_ch), np.float32),
                        meta=np.array([ch, fs, br, vbr, dec_ch], np.int32))
    return packets.shape


if __name__ == "__main__":
    for n in (sys.argv[1:] or list(CONFIGS) + list(POOLS)):
        print(n, make(n))
