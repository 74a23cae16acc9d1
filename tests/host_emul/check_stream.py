import sys, ctypes as C, os
ROOT=os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0,ROOT); sys.path.insert(0,os.path.join(ROOT,'tests'))
import numpy as np
from conftest import golden_names, load_golden
from oracle import oraclepy
L=C.CDLL(os.path.join(ROOT,'tests/host_emul/libemul.so'))
def P(a,t): return a.ctypes.data_as(C.POINTER(t))
worst=0
for name in (sys.argv[1:] or golden_names()):
    g=load_golden(name)
    for s in range(g['packets'].shape[0]):
        pk=np.ascontiguousarray(g['packets'][s]); ln=np.ascontiguousarray(g['lens'][s]); nf=pk.shape[0]; fs=g['frame_size']; dc=g['dec_channels']
        opcm,orng,osmp,taps=oraclepy.decode_stream(pk,ln,fs,dc,want_taps=True)
        pcm=np.zeros((nf,fs*dc),np.float32); rng=np.zeros(nf,np.uint32); smp=np.zeros(nf,np.int32); X=np.zeros((nf,1920),np.float32)
        L.emul_decode_stream(P(pk,C.c_ubyte),P(ln,C.c_int),pk.shape[1],nf,fs,dc,P(pcm,C.c_float),P(rng,C.c_uint32),P(smp,C.c_int),P(X,C.c_float))
        assert (smp==osmp).all(), (name,s,smp[:5],osmp[:5])
        assert (rng==orng).all()
        N=fs; Cc=g['channels']
        EB=[0,1,2,3,4,5,6,7,8,10,12,14,16,20,24,28,34,40,48,60,78,100]
        def xdiff(f):
            t=taps[f]; M=1<<t.LM; lim=M*EB[t.end]; a=np.array(t.X[:Cc*N]).reshape(Cc,N)[:,:lim]; b=X[f,:Cc*N].reshape(Cc,N)[:,:lim]
            return np.abs(a-b)
        xerr=max(float(xdiff(f).max()) for f in range(nf))
        perr=float(np.abs(pcm-opcm).max())
        worst=max(worst,perr)
        flag = '' if (xerr<1e-5 and perr<1e-5) else '   <<<<<<<<'
        print('%-28s s%d X err %.2e pcm err %.2e%s'%(name,s,xerr,perr,flag))
        if flag:
            for f in range(nf):
                e=xdiff(f).reshape(-1)
                if e.max()>1e-5:
                    print('  first bad X frame',f,'idx',int(e.argmax()),'transient',taps[f].transient,'LM',taps[f].LM,'dual',taps[f].dual_stereo,'int',taps[f].intensity); break
            for f in range(nf):
                e=np.abs(pcm[f]-opcm[f])
                if e.max()>1e-5:
                    print('  first bad pcm frame',f,'idx',int(e.argmax()),'transient',taps[f].transient, 'pf', taps[f].pf_on, taps[f].pf_pitch,'acol',taps[f].anti_collapse_on); break
print('worst pcm err %.3e'%worst)
