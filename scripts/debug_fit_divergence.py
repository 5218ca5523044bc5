"""Debug helper (GPU): run the tutorial fit pipeline twice (general vs segment E-step kernel) and report where the
two trajectories separate; also time single H-step objective evaluations."""
import copy, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vlgp_b200 import core, preprocess
from vlgp_b200.core import Session
from vlgp_b200.gp import make_cholesky
from vlgp_b200.synth import make_trials
from vlgp_b200.util import cut_trials


def prepare():
    trials = make_trials(10, 200, 30, 3, seed=0)
    config = preprocess.get_config(max_iter=1, min_iter=1)
    params = preprocess.get_params(trials, 3, omega_bound=config["omega_bound"])
    np.random.seed(0)
    preprocess.initialize(trials, params, config)
    preprocess.fill_params(params)
    preprocess.fill_trials(trials)
    params.pop("transform")
    make_cholesky(trials, params, config)
    core.update_w(trials, params, config)
    core.update_v(trials, params, config)
    segs = list(cut_trials(trials, params, config))
    make_cholesky(segs, params, config)
    preprocess.fill_trials(segs)
    return segs, params, config


def rel(a, b):
    return float(np.max(np.abs(a - b)) / np.max(np.abs(b)))


segs, params, config = prepare()
runs = {}
for mode in ("generic", "seg"):
    if mode == "generic":
        os.environ["VLGP_FORCE_GENERIC_ESTEP"] = "1"
    else:
        os.environ.pop("VLGP_FORCE_GENERIC_ESTEP", None)
    s, p, c = copy.deepcopy(segs), copy.deepcopy(params), copy.deepcopy(config)
    hist = []
    sys.stdout = open(os.devnull, "w")
    for it in range(3):
        # E-step only, then the rest, recording the state in between
        sess = Session(s, p)
        core._constrain_loading_dev(sess, p, c)
        sess.ts.estep(c["Eniter"], c["dmu_bound"], c["method"])
        st_e = sess.ts.get_state()
        core._mstep_dev(sess, p, c)
        a_m = p["a"].copy()
        core._hstep_dev(sess, s, p, c)
        sess.pull(s)
        sess.close()
        hist.append(dict(e=st_e, a=a_m, omega=p["omega"].copy(), nfev=c["hstep_nfev"][-1]))
    sys.stdout = sys.__stdout__
    runs[mode] = hist
for it in range(3):
    g, q = runs["generic"][it], runs["seg"][it]
    print("it", it, "estep mu", rel(q["e"]["mu"], g["e"]["mu"]), "v", rel(q["e"]["v"], g["e"]["v"]), "w",
          rel(q["e"]["w"], g["e"]["w"]), "a", rel(q["a"], g["a"]), "omega", rel(q["omega"], g["omega"]), q["omega"],
          g["nfev"], q["nfev"])

# H-step evaluation latency
segs, params, config = prepare()
with Session(segs, params) as sess:
    sess.ts.hstep_prepare()
    h = np.array([1.0, 0.02, 1e-4])
    for _ in range(5):
        sess.ts.hstep_objective(0, h)
    t0 = time.perf_counter()
    for _ in range(200):
        sess.ts.hstep_objective(0, h)
    print("hstep_objective latency (40 segments): %.1f us" % ((time.perf_counter() - t0) / 200 * 1e6))
