#!/usr/bin/env python
"""One-off check, BUILD container only (imports the reference via oracle/ref_shim.py): fit + transform() of new trials +
sample_posterior() by the reference and by this package over the oracle stand-in.  Round 1: Hstep=False cases agree to
1e-15 (samples included); with the H-step on omega is an L-BFGS-B end point (1e-12 .. 1e-7) and the samples, which go
through an SVD of the covariance, are not comparable; a trial length without a prior factor raises KeyError in both."""
import sys, copy, os
os.environ.setdefault("OPENBLAS_NUM_THREADS","1")
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np
from oracle import ref_shim
import oracle_engine, vlgp_b200, vlgp_b200.engine as engine_mod
from vlgp_b200.synth import make_trials
ref=ref_shim.load()
def relerr(x, r):
    x, r = np.asarray(x, float), np.asarray(r, float); return float(np.max(np.abs(x-r))/max(np.max(np.abs(r)),1e-300))
rng=np.random.default_rng(3)
for case in range(6):
    N,L=int(rng.integers(5,12)),int(rng.integers(1,4))
    kw=dict(max_iter=2,min_iter=2,Hstep=bool(case%2==0))
    new_lengths=[100,100,100] if case<5 else [100,71]
    res=[]
    for which in ("ref","our"):
        tr=make_trials(3,100,N,L,seed=case)
        new=[]
        for i,T in enumerate(new_lengths): new+=make_trials(1,T,N,L,seed=100+10*case+i)
        np.random.seed(case)
        if which=="ref":
            r=ref.fit(tr,L,**copy.deepcopy(kw))
            try: out=ref.transform(new,r["params"],r["config"])
            except Exception as e: print(case,"ref raises",type(e).__name__,e); res.append(None); continue
            smp=None
            np.random.seed(9); smp=ref.sample_posterior(out[0], r["params"], 3)
        else:
            engine_mod._ENGINE=oracle_engine.OracleEngine()
            r=vlgp_b200.fit(tr,L,**copy.deepcopy(kw))
            try: out=vlgp_b200.transform(new,r["params"],r["config"])
            except Exception as e: print(case,"our raises",type(e).__name__,e); res.append(None); continue
            np.random.seed(9); smp=vlgp_b200.sample_posterior(out[0], r["params"], 3)
        res.append((out,r,smp))
    if None in res: continue
    (o1,r1,s1),(o2,r2,s2)=res
    e={k: relerr(np.concatenate([t[k] for t in o2]), np.concatenate([t[k] for t in o1])) for k in ("mu","v","w")}
    e["omega"]=relerr(r2["params"]["omega"], r1["params"]["omega"]); e["sample"]=relerr(s2,s1)
    print(case, "Hstep",kw["Hstep"], new_lengths, {k:"%.1e"%v for k,v in e.items()}, "returns same list:", o2 is not None and len(o2)==len(o1), sorted(o1[0].keys())==sorted(o2[0].keys()))
