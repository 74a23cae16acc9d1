import sys, os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),'tests'))
import numpy as np
from opus_codec_b200 import synth
from opus_codec_b200.batch import BatchEncoder
from test_gpu_encode import _ref_c_encode
ch,br,fs,vbr,cx=2,96000,960,0,6
S=int(sys.argv[2]) if len(sys.argv)>2 else 3; nfr=int(sys.argv[1]) if len(sys.argv)>1 else 4
pcm=np.stack([synth.stream_pcm(s,960*nfr,ch,base_seed=777) for s in range(S)]); F=nfr
with BatchEncoder(S,48000,ch,device=0,max_frames=F) as enc:
    enc.set_bitrate(br); enc.set_complexity(cx); enc.set_vbr(vbr!=0)
    out,lens,rng=enc.encode_float_multi(pcm.reshape(S,F,fs*ch),fs)
    print('lens zero count',int((lens==0).sum()),'of',lens.size, 'neg',int((lens<0).sum()), 'zero positions', np.argwhere(lens==0)[:10].tolist())
for s in range(S):
    ro,rl,rr=_ref_c_encode(pcm[s],fs,ch,br,vbr,cx)
    bad=[f for f in range(F) if not np.array_equal(ro[f,:rl[f]],out[s,f,:rl[f]])]
    print('stream',s,'bad frames',len(bad),bad[:8],'tail garbage frames',int(sum((out[s,f,rl[f]:]!=0).any() for f in range(F))))
