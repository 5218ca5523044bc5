#!/usr/bin/env python
"""One-off check, run in the BUILD container only (it imports the unmodified reference from /root/reference through
oracle/ref_shim.py): whole fits on RANDOM trial lengths that are not multiples of the window -- overlapping, aliased
segments (vlgp/util.py:482-498) -- by the reference and by this package's host code (api.fit / core.vem / _Aliasing)
over the oracle stand-in engine (tests/oracle_engine.py).  Extends the three committed cases of
tests/golden/fit_overlap.npz to arbitrary junction patterns (zero-overlap junctions, chains of several windows).

    python scripts/fuzz_overlap_vs_reference.py [n_cases]

Prints one line per case and the worst relative differences; exits non-zero above 1e-8.
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
    n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    rng = np.random.default_rng(2024)
    options = [dict(), dict(Hstep=False), dict(Hstep=False, constrain_latent="both"), dict(constrain_loading=2),
               dict(Hstep=False, constrain_latent="location", constrain_loading="fro")]
    worst = 0.0
    for case in range(n_cases):
        n_trials = int(rng.integers(2, 6))
        lengths = [int(rng.integers(51, 330)) for _ in range(n_trials)]
        if case % 3 == 0:
            lengths[0] = 50 * int(rng.integers(1, 4))            # one trial without overlap in the mix
        N, L = int(rng.integers(8, 16)), int(rng.integers(1, 4))
        kw = dict(options[case % len(options)], max_iter=2, min_iter=2)
        seed = int(rng.integers(1 << 30))

        def trials():
            out = []
            for i, T in enumerate(lengths):
                out += make_trials(1, T, N, L, seed=1000 * case + i)
            return out

        t_ref, t_our = trials(), trials()
        np.random.seed(seed)
        r_ref = ref.fit(t_ref, L, **copy.deepcopy(kw))
        engine_mod._ENGINE = oracle_engine.OracleEngine()
        np.random.seed(seed)
        r_our = vlgp_b200.fit(t_our, L, **copy.deepcopy(kw))
        errs = {k: relerr(np.concatenate([t[k] for t in t_our]), np.concatenate([t[k] for t in t_ref]))
                for k in ("mu", "v", "w")}
        errs.update({k: relerr(r_our["params"][k], r_ref["params"][k]) for k in ("a", "b", "noise", "omega")})
        levels = max((len(d) for n, d in engine_mod._ENGINE.log if n == "estep" and isinstance(d, tuple)), default=0)
        sub = sum(1 for n, d in engine_mod._ENGINE.log if n == "estep" and isinstance(d, tuple))
        w = max(errs.values())
        worst = max(worst, w)
        print("case %2d lengths %-28s N=%2d L=%d %-60s subset E-steps %3d  worst %.1e (%s)"
              % (case, lengths, N, L, {k: v for k, v in kw.items() if k not in ("max_iter", "min_iter")}, sub, w,
                 max(errs, key=errs.get)))
    print("worst relative difference over %d cases: %.2e" % (n_cases, worst))
    sys.exit(0 if worst < 1e-8 else 1)


if __name__ == "__main__":
    main()
