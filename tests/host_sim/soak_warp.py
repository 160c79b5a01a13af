"""TEST INFRASTRUCTURE ONLY: randomized soak of the whole-warp shrinking-window walk (32 host threads in lock step,
tests/host_sim/sim_band.cpp) against the oracle, far beyond what the CPU suite runs every time.

    python -m pytest tests/test_host_sim.py -q        # builds tests/host_sim/libsim_band.so
    python tests/host_sim/soak_warp.py <seed> <seconds>

Regimes: read lengths 30-2200, error rates 0-30 %, truncated / unrelated / identical targets, thresholds at the
answers, random, at the cap (pilot) or around a percentile, lanes without a pair.  Round 1: six seeds x 240 s =
43 000 walks, 1.4 M lane results, all equal to the oracle.
"""
import os
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
import ctypes, sys, time
sys.path.insert(0, ROOT)
import numpy as np
from isocon_b200 import workloads
from oracle import oracle as O
L=ctypes.CDLL(os.path.join(HERE, 'libsim_band.so'))
L.sim_warp_diag_run.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.c_char_p, ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int), ctypes.c_int, ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int)]
seed=int(sys.argv[1]); budget=float(sys.argv[2])
rng=np.random.default_rng(seed)
t0=time.time(); runs=0; lanes=0
while time.time()-t0 < budget:
    Ln=int(rng.integers(30, 2200))
    root=rng.integers(0,4,size=Ln,dtype=np.uint8)
    err=float(rng.choice([0.0,0.005,0.02,0.05,0.1,0.18,0.3]))
    def mk():
        base = root
        if rng.random()<0.3: base = workloads._mutate(rng, root, 0.01,0.01,0.01)
        r = workloads._mutate(rng, base, err*0.45, err*0.3, err*0.25)
        if rng.random()<0.05 and r.size>40: r = r[:int(rng.integers(r.size//2, r.size))]
        return r
    q=workloads._to_str(mk()).encode()
    if len(q)==0: continue
    ts=[workloads._to_str(mk()).encode() for _ in range(32)]
    if rng.random()<0.2: ts[int(rng.integers(0,32))]=workloads._to_str(rng.integers(0,4,size=max(1,Ln+int(rng.integers(-20,20))),dtype=np.uint8)).encode()
    if rng.random()<0.1: ts[int(rng.integers(0,32))]=q
    ts=[t if len(t)>0 else b"A" for t in ts]
    ts.sort(key=len)
    ds=[O.ed_myers64(q,t,-1) for t in ts]
    style=int(rng.integers(0,4))
    if style==0: ks=[min(400,max(0,d+int(rng.integers(-5,6)))) for d in ds]
    elif style==1: ks=[int(rng.integers(0,401)) for _ in ds]
    elif style==2: ks=[400]*32
    else:
        b=int(np.percentile(ds,int(rng.integers(5,95)))); ks=[min(400,max(0,b+int(rng.integers(-40,40)))) for _ in ds]
    for l in range(32):
        if rng.random()<0.05: ks[l]=-1
    toff=np.zeros(33,dtype=np.int32); toff[1:]=np.cumsum([len(t) for t in ts])
    res=(ctypes.c_int*32)(); out=(ctypes.c_int*3)()
    for narrow in (1,int(rng.integers(1,7))):
        rc=L.sim_warp_diag_run(q,len(q),b"".join(ts),toff.ctypes.data_as(ctypes.POINTER(ctypes.c_int)),(ctypes.c_int*32)(*ks),narrow,res,out)
        if rc in (-1,-99): continue
        assert rc==0, rc
        for l in range(32):
            k=ks[l]; want=-1 if (k<0 or abs(len(ts[l])-len(q))>k or ds[l]>k) else ds[l]
            assert res[l]==want,(seed,runs,narrow,l,k,ds[l],res[l],len(q),len(ts[l]))
        runs+=1; lanes+=32
print("seed",seed,"runs",runs,"lanes",lanes,"ok")
