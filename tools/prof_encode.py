"""Profiling aid: run only the encoder kernel (config 3 shape: stereo, complexity 10, 96 kb/s CBR) on S streams x F frames.
usage: python tools/prof_encode.py [S] [F] [reps] [mapping: 0 auto, 1 warp per stream, 2 lane per stream]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from opus_codec_b200 import synth
from opus_codec_b200.batch import BatchEncoder

S = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
F = int(sys.argv[2]) if len(sys.argv) > 2 else 4
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
mapping = int(sys.argv[4]) if len(sys.argv) > 4 else 0
pool = np.stack([synth.stream_pcm(s, 960 * F, 2, base_seed=4242) for s in range(min(S, 96))])
pcm = np.ascontiguousarray(pool[np.arange(S) % pool.shape[0]].reshape(S, F, 1920))
with BatchEncoder(S, 48000, 2, device=0, max_frames=F) as enc:
    enc.set_mapping(mapping); enc.set_bitrate(96000); enc.set_complexity(10); enc.set_vbr(False)
    for r in range(reps):
        t = time.time()
        out, lens, rng = enc.encode_float_multi(pcm, 960)
        print("rep", r, "wall %.1f ms" % ((time.time() - t) * 1e3), "kernel ms", enc.kernel_ms(), "bad", int((lens <= 0).sum()))
