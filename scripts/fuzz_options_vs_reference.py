#!/usr/bin/env python
"""One-off check, run in the BUILD container only (imports the unmodified reference through oracle/ref_shim.py): whole
fits under RANDOM combinations of the reference's keyword arguments -- likelihood mix, method, Hessian / gradient
M-step, learning rate, step bounds, window, iteration counts, constraints, user-supplied initial values -- by the
reference and by this package's host code over the oracle stand-in engine.  Pins the oracle restatement and the host
orchestration on option combinations beyond the committed golden cases (tests/golden/fit_options.npz, vem_options.npz).

    python scripts/fuzz_options_vs_reference.py [n_cases] [seed]
"""
import copy
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")


def relerr(x, ref):
    x, ref = np.asarray(x, float), np.asarray(ref, float)
    return float(np.max(np.abs(x - ref)) / max(np.max(np.abs(ref)), 1e-300))


def main():
    from oracle import ref_shim
    import oracle_engine
    import vlgp_b200
    import vlgp_b200.engine as engine_mod
    from vlgp_b200.synth import make_trials

    ref = ref_shim.load()
    n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 10
    rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 7)
    worst = 0.0
    for case in range(n_cases):
        N, L = int(rng.integers(5, 14)), int(rng.integers(1, 4))
        window = int(rng.choice([25, 40, 50]))
        n_trials = int(rng.integers(2, 5))
        T = window * int(rng.integers(2, 5))                    # multiples of the window: no overlap in this script
        kw = dict(max_iter=int(rng.integers(1, 4)), window=window)
        kw["min_iter"] = kw["max_iter"]
        if rng.random() < 0.5:
            kw["Hstep"] = False
        if rng.random() < 0.3:
            kw["method"] = "MAP"
        if rng.random() < 0.3:
            kw["use_hessian"] = False
            kw["learning_rate"] = float(rng.choice([0.01, 0.05]))
        if rng.random() < 0.3:
            kw["Eniter"], kw["Mniter"] = int(rng.integers(1, 8)), int(rng.integers(1, 8))
        if rng.random() < 0.3:
            kw["da_bound"], kw["db_bound"], kw["dmu_bound"] = 0.5, 0.3, 0.7
        kw["constrain_loading"] = rng.choice(["fro", "none", 1, 2, "svd"]).item() if rng.random() < 0.6 else "fro"
        if kw["constrain_loading"] in ("1", "2"):
            kw["constrain_loading"] = int(kw["constrain_loading"])
        if rng.random() < 0.4:
            kw["constrain_latent"] = str(rng.choice(["location", "scale", "both"]))
        if rng.random() < 0.3:
            kw["lik"] = [str(x) for x in rng.choice(["poisson", "gaussian"], N)]
        if rng.random() < 0.25:
            kw["omega"] = rng.uniform(1e-3, 2e-2, L)
            kw["sigma"] = rng.uniform(0.5, 1.5, L)
        seed = int(rng.integers(1 << 30))

        def trials():
            out = make_trials(n_trials, T, N, L, seed=500 + case)
            if "lik" in kw:                                      # Gaussian channels get real-valued observations
                g = np.array(kw["lik"]) == "gaussian"
                r2 = np.random.default_rng(900 + case)
                for tr in out:
                    tr["y"] = tr["y"].astype(float)
                    tr["y"][:, g] = r2.standard_normal((T, int(g.sum()))) * 0.5 + 0.2
            return out

        t_ref, t_our = trials(), trials()
        try:
            np.random.seed(seed)
            r_ref = ref.fit(t_ref, L, **copy.deepcopy(kw))
        except Exception as e:  # the reference itself rejects some combinations: the port must reject them too
            engine_mod._ENGINE = oracle_engine.OracleEngine()
            try:
                np.random.seed(seed)
                vlgp_b200.fit(t_our, L, **copy.deepcopy(kw))
                print("case %2d: reference raised %r but the port did not  %s" % (case, e, kw))
                worst = max(worst, 1.0)
            except Exception as e2:
                print("case %2d: both raise (%s / %s)" % (case, type(e).__name__, type(e2).__name__))
            continue
        engine_mod._ENGINE = oracle_engine.OracleEngine()
        np.random.seed(seed)
        r_our = vlgp_b200.fit(t_our, L, **copy.deepcopy(kw))
        errs = {k: relerr(np.stack([t[k] for t in t_our]), np.stack([t[k] for t in t_ref])) for k in ("mu", "v", "w")}
        errs.update({k: relerr(r_our["params"][k], r_ref["params"][k]) for k in ("a", "b", "noise", "omega", "sigma")})
        same_it = r_our["config"]["runtime"]["it"] == r_ref["config"]["runtime"]["it"]
        # With the H-step on, omega is an L-BFGS-B end point: it agrees to ~1e-12, not bit for bit, and the rank-50
        # factors of the UNCUT trials (built from it after vem) have exact pivot ties that such a difference can flip --
        # the final mu / v / w then differ at the percent level although everything vem computed agrees (DESIGN.md
        # section 5, "known limit"; with Hstep=False the same cases agree to 1e-15).  Judge those cases on the
        # parameters, report the rest.
        hstep_on = kw.get("Hstep", True)
        judged = ("a", "b", "noise", "omega", "sigma") if hstep_on else tuple(errs)
        w = max(errs[k] for k in judged)
        if hstep_on and max(errs.values()) > 1e-7:
            print("         (pivot-tie sensitive case: mu %.1e v %.1e w %.1e with omega equal to %.1e)"
                  % (errs["mu"], errs["v"], errs["w"], errs["omega"]))
        worst = max(worst, w if same_it else 1.0)
        show = {k: (v if not isinstance(v, (list, np.ndarray)) else "...") for k, v in kw.items()}
        print("case %2d N=%2d L=%d trials=%d T=%3d  worst %.1e (%s) it %s  %s" % (case, N, L, n_trials, T, w,
                                                                                max(judged, key=errs.get), same_it, show))
    print("worst relative difference over %d cases: %.2e" % (n_cases, worst))
    sys.exit(0 if worst < 1e-7 else 1)


if __name__ == "__main__":
    main()
