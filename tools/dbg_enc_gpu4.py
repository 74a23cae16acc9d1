import sys, os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),'tests'))
import numpy as np
from opus_codec_b200 import synth
from opus_codec_b200.batch import BatchEncoder
from test_gpu_encode import _ref_c_encode
S,F,fs,ch,br,cx=8,6,960,2,128000,0
def run(pcm, tag, force=0):
    with BatchEncoder(S,48000,ch,device=0,max_frames=F) as enc:
        enc.set_bitrate(br); enc.set_complexity(cx); enc.set_vbr(False)
        if force: enc.set_force_channels(force)
        out,lens,rng=enc.encode_float_multi(pcm.reshape(S,F,fs*ch),fs)
    ref={}
    res=[]
    for s in range(S):
        if force: res.append(int((lens[s]<=0).sum())); continue
        ro,rl,rr=_ref_c_encode(pcm[s],fs,ch,br,0,cx)
        res.append(sum(1 for f in range(F) if lens[s,f]!=rl[f] or not np.array_equal(ro[f,:rl[f]],out[s,f,:rl[f]])))
    same_as_0=[bool(np.array_equal(out[s],out[0])) for s in range(S)]
    print(tag,'bad',res,'lens',lens[:,0].tolist(),'zero lens',int((lens==0).sum()),'same as stream0',same_as_0)
one=synth.stream_pcm(1,960*F,ch,base_seed=777)
run(np.stack([one]*S),'identical streams')
pcm=np.stack([synth.stream_pcm(s,960*F,ch,base_seed=777) for s in range(S)])
run(pcm,'distinct streams')
run(pcm,'distinct, force mono',force=1)
