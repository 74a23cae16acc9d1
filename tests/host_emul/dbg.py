import sys, ctypes as C, os
ROOT=os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0,ROOT); sys.path.insert(0,os.path.join(ROOT,'tests'))
import numpy as np
from conftest import load_golden
from oracle import oraclepy
L=C.CDLL(os.path.join(ROOT,'tests/host_emul/libemul.so'))
def P(a,t): return a.ctypes.data_as(C.POINTER(t))
name=sys.argv[1]; s=int(sys.argv[2])
g=load_golden(name)
pk=np.ascontiguousarray(g['packets'][s]); ln=np.ascontiguousarray(g['lens'][s]); nf=pk.shape[0]; fs=g['frame_size']; dc=g['dec_channels']
opcm,orng,osmp,taps=oraclepy.decode_stream(pk,ln,fs,dc,want_taps=True)
pcm=np.zeros((nf,fs*dc),np.float32); rng=np.zeros(nf,np.uint32); smp=np.zeros(nf,np.int32); X=np.zeros((nf,1920),np.float32)
L.emul_decode_stream(P(pk,C.c_ubyte),P(ln,C.c_int),pk.shape[1],nf,fs,dc,P(pcm,C.c_float),P(rng,C.c_uint32),P(smp,C.c_int),P(X,C.c_float))
EB=[0,1,2,3,4,5,6,7,8,10,12,14,16,20,24,28,34,40,48,60,78,100]
N=fs; Cc=g['channels']
for f in range(nf):
    t=taps[f]; M=1<<t.LM
    a=np.array(t.X[:Cc*N]).reshape(Cc,N); b=X[f,:Cc*N].reshape(Cc,N)
    bad=[]
    for i in range(t.end):
        for c in range(Cc):
            e=np.abs(a[c,M*EB[i]:M*EB[i+1]]-b[c,M*EB[i]:M*EB[i+1]]).max()
            if e>1e-5: bad.append((i,c,float(e)))
    if bad:
        print('frame',f,'LM',t.LM,'tr',t.transient,'int',t.intensity,'dual',t.dual_stereo,'spread',t.spread,'tf',list(t.tf_res)); print(' bad bands',bad[:8])
        i,c,_=bad[0]
        print(' oracle X',a[c,M*EB[i]:M*EB[i+1]][:12]); print(' emul   X',b[c,M*EB[i]:M*EB[i+1]][:12])
        print(' oracle Y',a[1-c,M*EB[i]:M*EB[i+1]][:12]); print(' emul   Y',b[1-c,M*EB[i]:M*EB[i+1]][:12])
        break
# dump leaves/bands of the bad frame
class Leaf(C.Structure):
    _fields_=[('off',C.c_uint16),('n',C.c_uint8),('K',C.c_uint8),('kind',C.c_uint8),('B',C.c_uint8),('lcg_before',C.c_uint16),('gain',C.c_float)]
class Band(C.Structure):
    _fields_=[('leaf_begin_a',C.c_uint16),('leaf_begin_b',C.c_uint16),('leaf_cnt_a',C.c_uint8),('leaf_cnt_b',C.c_uint8),('eff_lowband',C.c_int16),('imid',C.c_int16),('iside',C.c_int16),('mode',C.c_uint8),('flags',C.c_uint8),('pad',C.c_uint8*2)]
buf=C.create_string_buffer(L.emul_ir_size())
L.emul_decode_symbols(P(np.ascontiguousarray(pk[f]),C.c_ubyte),int(ln[f]),dc,960,buf)
hs=L.emul_hdr_size()
bands=(Band*21).from_buffer(buf,hs)
leaves=(Leaf*672).from_buffer(buf,hs+C.sizeof(Band)*21)
for i in range(max(0,bad[0][0]-2),21):
    b=bands[i]
    print('band',i,'mode',b.mode,'flags',b.flags,'efflb',b.eff_lowband,'imid',b.imid,'iside',b.iside)
    for l in list(range(b.leaf_begin_a,b.leaf_begin_a+b.leaf_cnt_a))+list(range(b.leaf_begin_b,b.leaf_begin_b+b.leaf_cnt_b)):
        lf=leaves[l]; print('    leaf',l,'off',lf.off,'n',lf.n,'K',lf.K,'kind',lf.kind,'B',lf.B,'lcg',lf.lcg_before,'gain',lf.gain)
