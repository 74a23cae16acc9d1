"""Profiling aid: the stereo decode shape of bench.py's transcode leg (16384 stereo 48 kHz CELT 20 ms @96 kb/s streams, F frames per call,
device-pointer path), timed with CUDA events on the decoder's stream.
usage: python tools/prof_decode_stereo.py [S] [F] [steps] [warmup]      (warmup 0 under ncu: every launch is then a measured one)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from opus_codec_b200 import _lib
from opus_codec_b200.batch import BatchDecoder

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
S = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
F = int(sys.argv[2]) if len(sys.argv) > 2 else 10
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
warm = int(sys.argv[4]) if len(sys.argv) > 4 else 2
L = _lib.lib()
z = np.load(os.path.join(ROOT, "tests", "golden", "cfg3_stereo_20ms_96k_cbr.npz"))
pk0, ln0, rg0 = z["packets"], z["lens"], z["dec_rng"]
idx = np.arange(S) % pk0.shape[0]
pk = np.ascontiguousarray(pk0[idx, :F]); ln = np.ascontiguousarray(ln0[idx, :F]).astype(np.int32)
stride = pk.shape[2]
offsets = (np.arange(S * F, dtype=np.int32) * stride).reshape(S, F)
dec = BatchDecoder(S, 48000, 2, device=0, max_frames=F)
ext = torch.cuda.ExternalStream(L.ob_decoder_cuda_stream(dec.handle), device=0)
dev = torch.device("cuda", 0)
d_pk = torch.from_numpy(pk.reshape(-1)).to(dev); d_off = torch.from_numpy(offsets.reshape(-1)).to(dev); d_len = torch.from_numpy(ln.reshape(-1)).to(dev)
d_pcm = torch.empty(S * F * 960 * 2, dtype=torch.float32, device=dev)
d_smp = torch.empty(S * F, dtype=torch.int32, device=dev); d_rng = torch.empty(S * F, dtype=torch.int32, device=dev)


def step():
    r = L.ob_decode_float_device(dec.handle, F, d_pk.data_ptr(), d_off.data_ptr(), d_len.data_ptr(), d_pcm.data_ptr(), 960, d_smp.data_ptr(), d_rng.data_ptr(), 0)
    assert r == 0, r


for _ in range(warm): step()
torch.cuda.synchronize()
first_ok = None
if warm == 0:
    step(); torch.cuda.synchronize()
    first_ok = bool((d_rng.cpu().numpy().view(np.uint32).reshape(S, F) == rg0[idx, :F].astype(np.uint32)).all())
    steps = max(0, steps - 1)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(ext)
for _ in range(steps): step()
e1.record(ext); torch.cuda.synchronize()
if steps:
    ms = e0.elapsed_time(e1) / steps
    print("kernel ms (symbols, bands, synth) of the last launch set:", [round(v, 3) for v in dec.kernel_ms()])
    print("stereo device path: %.3f ms/step  %.0f audio-s/s" % (ms, S * F * 0.02 / ms * 1e3))
if first_ok is not None:
    print("final ranges of the first call equal the reference's:", first_ok)
