#!/usr/bin/env python
"""Mutation fuzzer for the native data.info reader (m6a_info_count / m6a_info_read): corrupted index files must be
refused with a status or read without touching memory outside the caller's buffers (guard bands; run with an ASan/UBSan
build of the library via M6A_LIB + LD_PRELOAD for the reads).   python tools/fuzz_info.py [iterations]"""
import sys, os, tempfile, ctypes as C, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from m6anet_b200 import _cabi
L=_cabi.lib()
good=open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden', 'bundled', 'data.info'), 'rb').read()[:3000]
rng=np.random.default_rng(1)
d=tempfile.mkdtemp(); path=os.path.join(d,'data.info').encode()
st={'ok':0,'rej':0}
for it in range(int(sys.argv[1]) if len(sys.argv)>1 else 2000):
    raw=bytearray(good)
    for _ in range(int(rng.integers(1,6))):
        k=rng.integers(0,4); n=len(raw)
        if n<5: break
        if k==0: raw[int(rng.integers(0,n))]=int(rng.integers(0,256))
        elif k==1: raw[int(rng.integers(0,n))]=int(rng.choice(list(b',\n\r-0 9')))
        elif k==2: a=int(rng.integers(0,n)); del raw[a:a+int(rng.integers(1,30))]
        else: del raw[int(rng.integers(0,n)):]
    open(path,'wb').write(raw)
    n_,nb=C.c_int64(0),C.c_int64(0)
    rc=L.m6a_info_count(path,C.byref(n_),C.byref(nb))
    if rc!=0: st['rej']+=1; continue
    n,nbv=int(n_.value),int(nb.value)
    G=16
    buf=np.full(max(nbv,1)+2*G,0x55,np.uint8); off=np.full(n+1+2*G,-7,np.int64); cols=[np.full(max(n,1)+2*G,-7,np.int64) for _ in range(4)]
    vp=lambda a,o=G: C.c_void_p(a.ctypes.data+o*a.itemsize)
    rc=L.m6a_info_read(path,n,nbv,vp(buf),vp(off),*[vp(c) for c in cols])
    assert rc<=0
    for a,fill in [(buf,0x55),(off,-7)]+[(c,-7) for c in cols]:
        assert (a[:G]==fill).all() and (a[len(a)-G:]==fill).all(), it
    st['ok' if rc==0 else 'rej']+=1
print(st)
