import sys, os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),'tests'))
import numpy as np
from opus_codec_b200 import synth
from opus_codec_b200.batch import BatchEncoder
from test_gpu_encode import _ref_c_encode
S,F,fs=8,6,960
for ch in (1,2):
  br=64000*ch
  pcm=np.stack([synth.stream_pcm(s,960*F,ch,base_seed=777) for s in range(S)])
  for cx in (0,1,2,3,4,5,6):
    with BatchEncoder(S,48000,ch,device=0,max_frames=F) as enc:
        enc.set_bitrate(br); enc.set_complexity(cx); enc.set_vbr(False)
        out,lens,rng=enc.encode_float_multi(pcm.reshape(S,F,fs*ch),fs)
    res=[]
    for s in range(S):
        ro,rl,rr=_ref_c_encode(pcm[s],fs,ch,br,0,cx)
        res.append(sum(1 for f in range(F) if lens[s,f]!=rl[f] or not np.array_equal(ro[f,:rl[f]],out[s,f,:rl[f]])))
    print('ch',ch,'cx',cx,'bad frames per stream',res,'zero lens',int((lens==0).sum()))
