import sys, ctypes as C, os
ROOT=os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0,ROOT); sys.path.insert(0,os.path.join(ROOT,'tests'))
import numpy as np
from conftest import golden_names, load_golden
from oracle import oraclepy
L=C.CDLL(os.path.join(ROOT,'tests/host_emul/libemul.so'))
class Hdr(C.Structure):
    _fields_=[('status',C.c_int32),('final_range',C.c_uint32),('n_leaves',C.c_uint16),('pf_pitch',C.c_uint16),
      ('LM',C.c_uint8),('C',C.c_uint8),('end',C.c_uint8),('flags',C.c_uint8),
      ('spread',C.c_uint8),('pf_tapset',C.c_uint8),('pf_qg',C.c_uint8),('coded_bands',C.c_uint8),
      ('intensity',C.c_uint8),('dual_stereo',C.c_uint8),('skip_in',C.c_uint8),('end_in',C.c_uint8),('lcg_total',C.c_uint32),('seed_in',C.c_uint32),('loss_in',C.c_int32),('lastfs_in',C.c_uint16),('pad2',C.c_uint16),
      ('coarse_qi',C.c_int16*42),('pulses',C.c_int16*21),('fine_quant',C.c_uint8*21),('fine_q2',C.c_uint8*42),
      ('final_bit',C.c_int8*42),('collapse_masks',C.c_uint8*42),('tf_change',C.c_int8*21),('pad1',C.c_uint8*3)]
assert C.sizeof(Hdr)==L.emul_hdr_size(), (C.sizeof(Hdr), L.emul_hdr_size())
buf=C.create_string_buffer(L.emul_ir_size())
tot=0;bad=0;maxleaves=0
for name in golden_names():
    g=load_golden(name)
    for s in range(g['packets'].shape[0]):
        pcm,rng,smp,taps=oraclepy.decode_stream(g['packets'][s],g['lens'][s],g['frame_size'],g['dec_channels'],want_taps=True)
        for f in range(g['packets'].shape[1]):
            pk=np.ascontiguousarray(g['packets'][s,f]); ln=int(g['lens'][s,f])
            L.emul_decode_symbols(pk.ctypes.data_as(C.POINTER(C.c_ubyte)),ln,g['dec_channels'],960,buf)
            h=Hdr.from_buffer(buf); t=taps[f]
            tot+=1; maxleaves=max(maxleaves,h.n_leaves)
            ok = h.final_range==int(rng[f]) and h.status==g['frame_size']
            ok = ok and list(h.coarse_qi)==list(t.coarse_qi) and list(h.pulses)==list(t.pulses) and list(h.tf_change)==list(t.tf_res)
            ok = ok and list(h.collapse_masks)==list(t.collapse_masks) and list(h.fine_quant)==list(t.fine_quant)
            ok = ok and h.coded_bands==t.coded_bands and h.intensity==t.intensity and h.dual_stereo==t.dual_stereo and h.spread==t.spread
            if not ok:
                bad+=1
                if bad<5:
                    print('MISMATCH',name,s,f,hex(h.final_range),hex(int(rng[f])),h.status, h.n_leaves)
                    print(list(h.collapse_masks)); print(list(t.collapse_masks))
                    print(list(h.pulses)); print(list(t.pulses))
print('frames',tot,'bad',bad,'max leaves',maxleaves)
