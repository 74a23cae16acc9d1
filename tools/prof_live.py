"""Profiling aid: blocking latency of one live decode call (F = 1, host buffers) for S mono streams; OB_DEC_CHUNKS selects the D2H pipeline depth.
usage: python tools/prof_live.py S [S ...]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from opus_codec_b200 import _lib
from opus_codec_b200.batch import BatchDecoder
L = _lib.lib()
zp = np.load(bench.POOL)
pk0, ln0, rg0 = zp["packets"], zp["lens"], zp["dec_rng"]
stride = pk0.shape[2]
for S in [int(a) for a in sys.argv[1:]] or [131072]:
    idx = np.arange(S) % pk0.shape[0]
    nfr = 9
    h_pk = [torch.from_numpy(np.ascontiguousarray(pk0[idx, f]).reshape(-1)).pin_memory() for f in range(nfr)]
    h_ln = [torch.from_numpy(np.ascontiguousarray(ln0[idx, f]).astype(np.int32)).pin_memory() for f in range(nfr)]
    h_off = torch.from_numpy(np.arange(S, dtype=np.int32) * stride).pin_memory()
    h_pcm = [torch.empty(S * 960, dtype=torch.float32).pin_memory() for _ in range(2)]
    h_smp = [torch.empty(S, dtype=torch.int32).pin_memory() for _ in range(2)]
    h_rng = [torch.empty(S, dtype=torch.int32).pin_memory() for _ in range(2)]
    dec = BatchDecoder(S, 48000, 1, device=0, max_frames=1)
    lat = []
    for n in range(nfr):
        p = n & 1
        t0 = time.perf_counter()
        assert L.ob_decode_float_multi_async(dec.handle, 1, h_pk[n].data_ptr(), h_off.data_ptr(), h_ln[n].data_ptr(), h_pcm[p].data_ptr(), 960, h_smp[p].data_ptr(), h_rng[p].data_ptr()) == 0
        assert L.ob_decoder_wait(dec.handle, 0) == 0
        if n >= 3: lat.append((time.perf_counter() - t0) * 1e3)
        assert (h_rng[p].numpy().view(np.uint32) == rg0[idx, n]).all()
    dec.close()
    print("S %d chunks %s: median %.2f ms  min %.2f" % (S, os.environ.get("OB_DEC_CHUNKS", "default"), float(np.median(lat)), min(lat)))
