"""Generates tests/golden/plc_*.npz: reference decoder output under packet loss and DTX payloads.

Run in the build container (needs oracle/_ref, compiled from /root/reference/opus):
    python tests/golden/make_golden_plc.py
Each file reuses the packets of the base golden it is named after and holds
    lens     i32 [streams, frames]   the base lens with lost packets (0) and DTX payloads (1 or 2 bytes) marked
    pcm_c    f32 [streams, frames, fs*ch]  output of the reference's PURE-C build (oracle/_ref/libopus_ref_c.so)
    samples  i32, ranges u32 [streams, frames]  return value / OPUS_GET_FINAL_RANGE after every call
    sse_diff f32 [streams, frames]   max |pcm(SSE build) - pcm(C build)|: the reference's own spread between its two builds,
                                     which reaches 2e-3 in concealed frames (the order-24 LPC analysis amplifies the different
                                     summation order of xcorr_kernel_sse); the parity bar for concealment is therefore the C build.
"""
import os
import sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import refpy                   # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
BASES = ["cfg2_mono_20ms_64k_cbr", "cfg1_stereo_20ms_128k_cbr", "cfg4_stereo_5ms_96k", "cfg4_mono_2p5ms_64k",
         "stereo_dec_of_mono_pkts", "mono_20ms_wb", "cfg4_stereo_10ms_96k"]
NSTREAMS, NFRAMES20 = 2, 32                     # 32 x 20 ms of audio per stream


def pattern(nf, fs, s):
    """Deterministic loss script: losses before any packet, isolated losses, short bursts, a burst long enough to switch to
    noise-based concealment (>= 100 ms), DTX payloads, and a random tail."""
    ln = np.ones(nf, np.int32)                       # 1 = keep
    u = max(1, 960 // fs)                            # frames per 20 ms
    def lose(a, b): ln[a:b] = 0
    if s == 0:
        lose(0, 2)
    lose(5 * u, 5 * u + 1)
    lose(8 * u, 8 * u + 3)
    lose(14 * u, 14 * u + 7 * u)                     # 140 ms
    ln[24 * u] = -1; ln[24 * u + 1] = -2             # DTX: TOC only / TOC + 1 byte
    r = np.random.default_rng(99 + s + fs)
    tail = np.arange(28 * u, nf)
    ln[tail[r.random(len(tail)) < 0.12]] = 0
    return ln


def make(base):
    g = np.load(os.path.join(HERE, base + ".npz"))
    ch, fs, br, vbr, dec_ch = [int(x) for x in g["meta"]]
    nf = min(NFRAMES20 * max(1, 960 // fs), g["packets"].shape[1])
    lens_all, pcm_all, smp_all, rng_all, diff_all = [], [], [], [], []
    for s in range(NSTREAMS):
        pk = np.ascontiguousarray(g["packets"][s, :nf])
        ln = np.ascontiguousarray(g["lens"][s, :nf]).copy()
        pat = pattern(nf, fs, s)
        ln[pat == 0] = 0
        ln[pat == -1] = 1
        ln[pat == -2] = 2
        pc, rc, sc = refpy.decode_stream(pk, ln, fs, dec_ch, pure_c=True)
        ps, rs, ss = refpy.decode_stream(pk, ln, fs, dec_ch)
        assert (rc == rs).all() and (sc == ss).all()
        lens_all.append(ln); pcm_all.append(pc); smp_all.append(sc); rng_all.append(rc)
        diff_all.append(np.abs(pc - ps).max(axis=1))
    np.savez_compressed(os.path.join(HERE, "plc_" + base + ".npz"), lens=np.stack(lens_all), pcm_c=np.stack(pcm_all),
                        samples=np.stack(smp_all), ranges=np.stack(rng_all), sse_diff=np.stack(diff_all).astype(np.float32))
    return nf, float(np.stack(diff_all).max())


if __name__ == "__main__":
    for b in (sys.argv[1:] or BASES):
        print(b, make(b))
