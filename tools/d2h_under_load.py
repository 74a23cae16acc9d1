#!/usr/bin/env python
"""Device->host copy rate of one GPU while its decoder kernels run (bench shape, device-pointer path) against the same copies on an idle GPU:
tells whether the float end-to-end leg (kernels of call n+1 overlap the PCM copies of call n) can expect the idle-GPU `host_ceiling_GBps`.
  python tools/d2h_under_load.py [S] [F] [copy MB] [host buffer MB]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from opus_codec_b200 import _lib
from opus_codec_b200.batch import BatchDecoder

S = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
F = int(sys.argv[2]) if len(sys.argv) > 2 else 200
MB = int(sys.argv[3]) if len(sys.argv) > 3 else 512
HMB = int(sys.argv[4]) if len(sys.argv) > 4 else 4096
L = _lib.lib()
pk, ln, _ = bench.load_pool(S, F)
offsets = (np.arange(S * F, dtype=np.int32) * pk.shape[2]).reshape(S, F)
dec = BatchDecoder(S, 48000, 1, device=0, max_frames=F)
dev = torch.device("cuda", 0)
d_pk = torch.from_numpy(pk.reshape(-1)).to(dev); d_off = torch.from_numpy(offsets.reshape(-1)).to(dev); d_len = torch.from_numpy(ln.reshape(-1)).to(dev)
d_pcm = torch.empty(S * F * 960, dtype=torch.float32, device=dev); d_smp = torch.empty(S * F, dtype=torch.int32, device=dev); d_rng = torch.empty(S * F, dtype=torch.int32, device=dev)
def step():
    assert L.ob_decode_float_device(dec.handle, F, d_pk.data_ptr(), d_off.data_ptr(), d_len.data_ptr(), d_pcm.data_ptr(), 960, d_smp.data_ptr(), d_rng.data_ptr(), 0) == 0
nbytes = MB << 20
src = torch.empty(nbytes, dtype=torch.uint8, device=dev)
host = torch.empty(HMB << 20, dtype=torch.uint8).pin_memory()      # larger than any host cache: the copies walk through it
nslot = (HMB << 20) // nbytes
cs = torch.cuda.Stream()
def copies(n):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(cs):
        e0.record(cs)
        for i in range(n):
            k = i % nslot
            host[k * nbytes:(k + 1) * nbytes].copy_(src, non_blocking=True)
        e1.record(cs)
    return e0, e1
for _ in range(3): step()
torch.cuda.synchronize()
n = max(4, int(2.0 * 50e9 / nbytes))
e0, e1 = copies(n); torch.cuda.synchronize()
idle = n * nbytes / (e0.elapsed_time(e1) / 1e3) / 1e9
k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ext = torch.cuda.ExternalStream(L.ob_decoder_cuda_stream(dec.handle), device=0)
k0.record(ext)
for _ in range(40): step()                       # ~2 s of kernels queued on the decoder's stream
k1.record(ext)
e0, e1 = copies(n); torch.cuda.synchronize()
busy = n * nbytes / (e0.elapsed_time(e1) / 1e3) / 1e9
print("d2h GB/s, %d MB copies into a %d MB pinned buffer: idle GPU %.1f, while the decoder kernels run %.1f (kernel step %.1f ms with copies running)" %
      (MB, HMB, idle, busy, k0.elapsed_time(k1) / 40))
