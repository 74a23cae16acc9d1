#!/usr/bin/env python
"""The reference crate's examples/basic_example.rs on the batched API: the same sine frame, for 1024 streams at once.
(Application::Audio here: the example's Application::Voip on a mono sine at the default bitrate is SILK territory, which stays on the
crate's own Encoder.)  Needs a B200: there is no CPU fallback."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from opus_codec_b200.batch import BatchDecoder, BatchEncoder, pack_packets  # noqa: E402

N_STREAMS = 1024
print("Opus Codec Basic Example, batched")
print("=================================")
with BatchEncoder(N_STREAMS, 48000, 1, application=2049, device=0) as encoder, BatchDecoder(N_STREAMS, 48000, 1, device=0) as decoder:
    print("created %d encoders and %d decoders: 48000 Hz, 1 channel, Audio application" % (N_STREAMS, N_STREAMS))
    encoder.set_bitrate(96000)      # at the default (51 kb/s mono) libopus gives tonal, speech-like frames to SILK: those come back as Unimplemented (-5)
    num_samples = 960                                                       # one 20 ms frame
    t = np.arange(num_samples, dtype=np.float32) / 48000.0
    freqs = 220.0 * 2.0 ** (np.arange(N_STREAMS, dtype=np.float32)[:, None] / 256.0)     # a different note per stream
    input_pcm = (np.sin(2 * np.pi * freqs * t) * 32767 * 0.1).astype(np.int16)
    out, lens, _ = encoder.encode_multi(input_pcm.reshape(N_STREAMS, 1, num_samples), num_samples, max_bytes=4000)
    print("encoded: %d..%d bytes per stream (compression ratio %.2f)" % (lens.min(), lens.max(), num_samples * 2.0 / lens.mean()))
    buf, offsets, plens = pack_packets([[bytes(out[s, 0, :lens[s, 0]])] for s in range(N_STREAMS)])
    decoded_pcm, decoded_samples, _ = decoder.decode_multi(buf, offsets, plens, num_samples)
    print("decoded %d samples per stream" % decoded_samples[0, 0])
    err = input_pcm.astype(np.float32) - decoded_pcm[:, 0].astype(np.float32)
    print("RMS reconstruction error of the first frame (codec delay included): %.2f" % np.sqrt((err ** 2).mean()))
    b = encoder.bitrate()
    print("bitrate:", "Auto" if b == -1000 else "Max" if b == -1 else "%d bps" % b)
    print("encoder complexity:", encoder.complexity())
    print("VBR:", encoder.vbr(), " lookahead:", encoder.lookahead(), "samples")
