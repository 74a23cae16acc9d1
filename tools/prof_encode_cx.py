"""Profiling aid: warp-per-stream encoder time per call against complexity / channels (which stages cost what).
usage: python tools/prof_encode_cx.py [S] [F] [mapping]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from opus_codec_b200 import synth
from opus_codec_b200.batch import BatchEncoder
S = int(sys.argv[1]) if len(sys.argv) > 1 else 2368
F = int(sys.argv[2]) if len(sys.argv) > 2 else 4
mapping = int(sys.argv[3]) if len(sys.argv) > 3 else 1
for ch, br in ((2, 96000), (1, 64000)):
    pool = np.stack([synth.stream_pcm(s, 960 * F, ch, base_seed=4242) for s in range(96)])
    pcm = np.ascontiguousarray(pool[np.arange(S) % 96].reshape(S, F, 960 * ch))
    for cx in (10, 8, 7, 5, 4, 2, 0):
        with BatchEncoder(S, 48000, ch, device=0, max_frames=F) as enc:
            enc.set_mapping(mapping); enc.set_bitrate(br); enc.set_complexity(cx); enc.set_vbr(False)
            for r in range(2):
                out, lens, rng = enc.encode_float_multi(pcm, 960)
            print("ch %d complexity %2d: kernel %.2f ms per call of %d frames x %d streams" % (ch, cx, enc.kernel_ms(), F, S))
