#!/usr/bin/env python
"""One-off probe, run in the BUILD container only (imports the unmodified reference through oracle/ref_shim.py): the
same unusual inputs through the reference's fit() and through this package's host code over the oracle stand-in --
integer / bool / float32 counts, extra user keys, user-supplied mu and x, a single trial, a silent neuron, counts above
255, unknown keyword arguments, empty / malformed trial lists, zero iteration counts.  Prints the worst relative
difference over mu, v, w, a, b, noise, omega, whether the returned dict surfaces agree, or the exception types.

Findings (round 1): everything agrees to 1e-15 and raises the same exception types, except
  * float32 counts: 5e-8 -- the reference's bias b inherits float32 from mean(y) and stays float32 through its in-place
    updates (vlgp/preprocess.py:22, vlgp/core.py:219); here b is float64;
  * counts of several hundred per bin (rates clipped at exp(10), weights ~1e4): 1e-3 -- the reference's variance formula
    (vlgp/core.py:107-111) cancels terms of size |A| down to 1/|A|, so its own result moves by 1e-6 with the LAPACK
    driver behind solve(); not a regime of spike-count data;
  * history=2 without regressors: the reference fails on a shape mismatch, this package says NotImplementedError.
"""
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
def run(name, mk, L, kw, seed=0, fn="fit"):
    outs=[]
    for which in ("ref","our"):
        tr=mk()
        np.random.seed(seed)
        try:
            if which=="ref": r=ref.fit(tr,L,**copy.deepcopy(kw))
            else:
                engine_mod._ENGINE=oracle_engine.OracleEngine(); r=vlgp_b200.fit(tr,L,**copy.deepcopy(kw))
        except Exception as e:
            outs.append(("EXC", type(e).__name__, str(e)[:70])); continue
        outs.append((tr,r))
    if outs[0][0]=="EXC" or outs[1][0]=="EXC":
        print("%-34s ref: %s | our: %s" % (name, outs[0][1:3] if outs[0][0]=="EXC" else "ok", outs[1][1:3] if outs[1][0]=="EXC" else "ok")); return
    (t1,r1),(t2,r2)=outs
    e={k: relerr(np.concatenate([t[k] for t in t2]), np.concatenate([t[k] for t in t1])) for k in ("mu","v","w")}
    e.update({k: relerr(r2["params"][k], r1["params"][k]) for k in ("a","b","noise","omega")})
    keys1=[sorted(t.keys()) for t in t1]; keys2=[sorted(t.keys()) for t in t2]
    dt=[(k, t1[0][k].dtype, t2[0][k].dtype) for k in t1[0] if hasattr(t1[0][k],'dtype') and t1[0][k].dtype!=t2[0][k].dtype]
    pk=sorted(set(r1["params"])^set(r2["params"])); ck=sorted(set(r1["config"])^set(r2["config"]))
    print("%-34s worst %.1e  keys_equal %s dtype_diff %s params_keydiff %s config_keydiff %s" % (name, max(e.values()), keys1==keys2, dt, pk, ck))
kw=dict(max_iter=2,min_iter=2,Hstep=False)
rng=np.random.default_rng(0)
run("int64 counts", lambda: [dict(y=np.random.default_rng(5+i).poisson(0.3,(100,7))) for i in range(3)], 2, kw)
run("extra user keys", lambda: [dict(y=np.random.default_rng(5+i).poisson(0.3,(100,7)).astype(float), id=i, ID="t%d"%i) for i in range(3)], 2, kw)
run("user mu", lambda: [dict(y=np.random.default_rng(5+i).poisson(0.3,(100,7)).astype(float), mu=np.random.default_rng(9+i).standard_normal((100,2))*0.1) for i in range(3)], 2, kw)
run("user x ones", lambda: [dict(y=np.random.default_rng(5+i).poisson(0.3,(100,7)).astype(float), x=np.ones((100,1,7))) for i in range(3)], 2, kw)
run("single trial L=1", lambda: [dict(y=np.random.default_rng(5).poisson(0.3,(150,7)).astype(float))], 1, kw)
run("silent neuron", lambda: [dict(y=np.concatenate([np.random.default_rng(5+i).poisson(0.3,(100,6)), np.zeros((100,1),int)],1).astype(float)) for i in range(3)], 2, kw)
run("counts > 255", lambda: [dict(y=np.random.default_rng(5+i).poisson(3.0,(100,7)).astype(float)*(1+100*(i==0))) for i in range(3)], 2, kw)
run("float32 y", lambda: [dict(y=np.random.default_rng(5+i).poisson(0.3,(100,7)).astype(np.float32)) for i in range(3)], 2, kw)
run("bool y", lambda: [dict(y=np.random.default_rng(5+i).poisson(0.3,(100,7))>0) for i in range(3)], 2, kw)
run("unknown kwargs", lambda: [dict(y=np.random.default_rng(5+i).poisson(0.3,(100,7)).astype(float)) for i in range(3)], 2, dict(kw, foo=1, bar="x"))
run("empty trial list", lambda: [], 2, kw)
run("missing y", lambda: [dict(z=1)], 2, kw)
run("n_factors > neurons", lambda: [dict(y=np.random.default_rng(5+i).poisson(0.3,(100,3)).astype(float)) for i in range(3)], 4, kw)
run("history=2 no x", lambda: [dict(y=np.random.default_rng(5+i).poisson(0.3,(100,7)).astype(float)) for i in range(3)], 2, dict(kw, history=2))
run("max_iter=0", lambda: [dict(y=np.random.default_rng(5+i).poisson(0.3,(100,7)).astype(float)) for i in range(3)], 2, dict(max_iter=0,min_iter=0,Hstep=False))
run("Eniter=0", lambda: [dict(y=np.random.default_rng(5+i).poisson(0.3,(100,7)).astype(float)) for i in range(3)], 2, dict(kw, Eniter=0))
run("Mniter=0", lambda: [dict(y=np.random.default_rng(5+i).poisson(0.3,(100,7)).astype(float)) for i in range(3)], 2, dict(kw, Mniter=0))
