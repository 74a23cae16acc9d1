"""Profiling aid: time the decoder's device-pointer path (bench config 2 shape) and, optionally, the host-pointer path.
usage: python tools/prof_decode.py [S] [F] [steps] [host]     (env OB_DEC_GROUPS / OB_DEC_CHUNKS select the overlap variants)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from opus_codec_b200 import _lib
from opus_codec_b200.batch import BatchDecoder

S = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
F = int(sys.argv[2]) if len(sys.argv) > 2 else 50
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
host = len(sys.argv) > 4
L = _lib.lib()
pk, ln, rng_expect = bench.load_pool(S, F)
stride = pk.shape[2]
offsets = (np.arange(S * F, dtype=np.int32) * stride).reshape(S, F)
dec = BatchDecoder(S, 48000, 1, device=0, max_frames=F)
ext = torch.cuda.ExternalStream(L.ob_decoder_cuda_stream(dec.handle), device=0)
dev = torch.device("cuda", 0)
d_pk = torch.from_numpy(pk.reshape(-1)).to(dev); d_off = torch.from_numpy(offsets.reshape(-1)).to(dev); d_len = torch.from_numpy(ln.reshape(-1)).to(dev)
d_pcm = torch.empty(S * F * 960, dtype=torch.float32, device=dev); d_smp = torch.empty(S * F, dtype=torch.int32, device=dev); d_rng = torch.empty(S * F, dtype=torch.int32, device=dev)
def step():
    r = L.ob_decode_float_device(dec.handle, F, d_pk.data_ptr(), d_off.data_ptr(), d_len.data_ptr(), d_pcm.data_ptr(), 960, d_smp.data_ptr(), d_rng.data_ptr(), 0)
    assert r == 0, r
for _ in range(3): step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(ext)
for _ in range(steps): step()
e1.record(ext); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
ok = bool((d_rng.cpu().numpy().view(np.uint32).reshape(S, F) == rng_expect).all())
try: print("kernel ms (symbols, bands, synth) of the last timed launch set:", [round(v, 3) for v in dec.kernel_ms()])
except Exception as ex: print("kernel ms unavailable", ex)
print("device path: %.3f ms/step  %.0f audio-s/s  ranges_ok=%s groups=%s" % (ms, S * F * 0.02 / ms * 1e3, ok, os.environ.get("OB_DEC_GROUPS", "1")))
if host:
    h_pk = torch.from_numpy(pk.reshape(-1).copy()).pin_memory(); h_off = torch.from_numpy(offsets.reshape(-1).copy()).pin_memory(); h_len = torch.from_numpy(ln.reshape(-1).copy()).pin_memory()
    h_pcm = torch.empty(S * F * 960, dtype=torch.float32).pin_memory(); h_smp = torch.empty(S * F, dtype=torch.int32).pin_memory(); h_rng = torch.empty(S * F, dtype=torch.int32).pin_memory()
    def hstep():
        r = L.ob_decode_float_multi(dec.handle, F, h_pk.data_ptr(), h_off.data_ptr(), h_len.data_ptr(), h_pcm.data_ptr(), 960, h_smp.data_ptr(), h_rng.data_ptr())
        assert r == 0, r
    hstep(); hstep()
    t = time.perf_counter()
    for _ in range(steps): hstep()
    ms = (time.perf_counter() - t) * 1e3 / steps
    ok = bool((h_rng.numpy().view(np.uint32).reshape(S, F) == rng_expect).all())
    print("host path:   %.3f ms/step  %.0f audio-s/s  ranges_ok=%s chunks=%s" % (ms, S * F * 0.02 / ms * 1e3, ok, os.environ.get("OB_DEC_CHUNKS", "default")))
