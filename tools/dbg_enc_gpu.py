import sys, os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),'tests'))
import numpy as np
from opus_codec_b200 import synth
from opus_codec_b200.batch import BatchEncoder
from test_gpu_encode import _ref_c_encode
for (ch,br,fs,vbr,cx) in [(2,96000,960,0,6),(2,96000,960,0,10),(1,64000,960,0,10),(2,96000,960,1,5)]:
    S=6; pcm=np.stack([synth.stream_pcm(s,48000,ch,base_seed=777) for s in range(S)]); F=pcm.shape[1]//(fs*ch)
    with BatchEncoder(S,48000,ch,device=0,max_frames=F) as enc:
        enc.set_bitrate(br); enc.set_complexity(cx); enc.set_vbr(vbr!=0); enc.set_vbr_constraint(vbr==2)
        out,lens,rng=enc.encode_float_multi(pcm.reshape(S,F,fs*ch),fs)
        fr=enc.final_range()
        print(ch,br,fs,vbr,cx,'lens min/max',lens.min(),lens.max(),'final_range eq',(fr==rng[:,-1]).tolist(), 'kernel ms', enc.kernel_ms())
    for s in range(S):
        ro,rl,rr=_ref_c_encode(pcm[s],fs,ch,br,vbr,cx)
        same=((ro==out[s]).all(axis=1)&(rl==lens[s]))
        print('   stream',s,'identical',same.sum(),'/',len(same),'first bad',int(np.argmin(same)) if not same.all() else -1, 'lens',lens[s][:6].tolist(), rl[:6].tolist())
