"""Dev aid: host-emulated product decoder vs the reference decoder with lost packets / DTX payloads."""
import sys, ctypes as C, os
ROOT=os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0,ROOT); sys.path.insert(0,os.path.join(ROOT,'tests'))
import numpy as np
from conftest import golden_names, load_golden
from oracle import refpy
L=C.CDLL(os.path.join(ROOT,'tests/host_emul/libemul.so'))
def P(a,t): return a.ctypes.data_as(C.POINTER(t))
def loss_pattern(nf, kind, seed):
    r=np.random.default_rng(seed); lost=np.zeros(nf,bool)
    if kind=='single': lost[[5,11,20]]=True
    elif kind=='burst': lost[7:10]=True; lost[20:32]=True
    elif kind=='start': lost[0:2]=True; lost[4]=True
    elif kind=='long': lost[10:40]=True
    elif kind=='random': lost=r.random(nf)<0.15
    return lost
worst=0
names=[a for a in sys.argv[1:] if not a.startswith('-')] or golden_names()
for name in names:
    g=load_golden(name)
    for kind in ('single','burst','start','long','random','dtx'):
        s=0
        pk=np.ascontiguousarray(g['packets'][s]).copy(); ln=np.ascontiguousarray(g['lens'][s]).copy(); nf=pk.shape[0]; fs=g['frame_size']; dc=g['dec_channels']
        if kind=='dtx':
            lost=loss_pattern(nf,'burst',1); ln[lost]=np.where(np.arange(nf)[lost]%2==0,1,2)
        else:
            lost=loss_pattern(nf,kind,hash(name)%1000); ln[lost]=0
        rpcm,rrng,rsmp=refpy.decode_stream(pk,ln,fs,dc,pure_c='-c' in sys.argv)[:3]
        pcm=np.zeros((nf,fs*dc),np.float32); rng=np.zeros(nf,np.uint32); smp=np.zeros(nf,np.int32)
        L.emul_decode_stream(P(pk,C.c_ubyte),P(ln,C.c_int),pk.shape[1],nf,fs,dc,P(pcm,C.c_float),P(rng,C.c_uint32),P(smp,C.c_int),None)
        ok_s=(smp==rsmp).all(); ok_r=(rng==rrng).all()
        err=np.abs(pcm-rpcm).reshape(nf,-1).max(axis=1)
        worst=max(worst,float(err.max()))
        bad=np.nonzero(err>1e-4)[0]
        print('%-30s %-7s lost %2d  samples %s ranges %s  max err %.2e  first bad frame %s'%(name,kind,int(lost.sum()),ok_s,ok_r,err.max(),(int(bad[0]),bool(lost[bad[0]])) if len(bad) else None))
        if not ok_s: print('   samples', smp[:12], rsmp[:12])
        if not ok_r: print('   ranges differ at', np.nonzero(rng!=rrng)[0][:8])
print('worst %.3e'%worst)
