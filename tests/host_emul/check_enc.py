import sys, ctypes as C, os
ROOT=os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0,ROOT)
import numpy as np
from opus_codec_b200 import synth
from oracle import refpy
L=C.CDLL(os.path.join(ROOT,'tests/host_emul/libemul.so'))
def P(a,t): return a.ctypes.data_as(C.POINTER(t))
def emul_encode(pcm, fs, ch, bitrate, nbytes, vbr=0, complexity=10, max_bytes=1275):
    pcm=np.ascontiguousarray(pcm,np.float32); nf=pcm.size//(fs*ch)
    out=np.zeros((nf,max_bytes),np.uint8); lens=np.zeros(nf,np.int32); rng=np.zeros(nf,np.uint32)
    r=L.emul_celt_encode_stream(P(pcm,C.c_float),nf,fs,ch,bitrate,vbr,complexity,nbytes,P(out,C.c_ubyte),max_bytes,P(lens,C.c_int),P(rng,C.c_uint32))
    assert r==0, r
    return out,lens,rng
if __name__=='__main__':
    cfgs=[(1,64000,960,159,0,10)] if len(sys.argv)<2 else [tuple(int(v) for v in a.split(',')) for a in sys.argv[1:]]
    for (ch,br,fs,nb,vbr,cx) in cfgs:
        for s in range(3):
            pcm=synth.stream_pcm(s,48000*2,ch)
            a=refpy.celt_encode_stream(pcm,fs,ch,br,nb,vbr=vbr,complexity=cx,pure_c=True)
            b=emul_encode(pcm,fs,ch,br,nb,vbr=vbr,complexity=cx)
            same=(a[0]==b[0]).all(axis=1)&(a[1]==b[1])
            first=int(np.argmin(same)) if not same.all() else -1
            print('ch%d br%d fs%d nb%d vbr%d cx%d stream %d: identical frames %d/%d first bad %d  lens ref %s ours %s'%(ch,br,fs,nb,vbr,cx,s,same.sum(),len(same),first,a[1][:3],b[1][:3]))
            if first>=0:
                x=a[0][first,:a[1][first]]; y=b[0][first,:b[1][first]]
                n=min(len(x),len(y)); d=np.nonzero(x[:n]!=y[:n])[0]
                print('   first differing byte', int(d[0]) if len(d) else 'len', 'rng ref %08x ours %08x'%(a[2][first],b[2][first]))
