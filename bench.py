#!/usr/bin/env python
"""Benchmark of the vLGP variational-EM hot path (BASELINE.json: EM-iterations/sec).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config config2] [--impl ours|reference]

One "step" = one EM iteration of ``vem`` on the cut segments: constrain_loading + E-step (Eniter=25), M-step
(Mniter=25), H-step (L-BFGS-B over omega per latent + new prior factors) -- vlgp/core.py:298-326 -- with the
reference's default configuration.  Workload at every N: BASELINE config 2 (256 trials x T=1000 x 100 neurons x 5
latents, Poisson, fp64; S = 5120 segments of 50 bins), synthetic spike trains, initialised exactly like ``fit``
(FactorAnalysis + update_w/update_v + cut_trials).  N > 1 (torchrun, one process per GPU) shards the TRIALS over
ranks: strong scaling of the same job, one NCCL allreduce per M-step Newton iteration / H-step evaluation.

Printed JSON (rank 0): ``value`` = EM-iterations/sec with the state resident in HBM (CUDA events on the engine's stream,
max over ranks); ``e2e`` = the same iteration through the public ``vem(splits, params, config)`` call with host
buffers (upload + iteration + download each step); ``roofline`` = the E-step kernel (dominant) against the measured
FP64 peak; ``cpu_baseline`` = the NumPy oracle port of the reference timed on a bounded sample on this host.

``--impl reference`` times the reference's algorithm on the host cores (the NumPy/SciPy oracle port: the reference
itself is pure Python and does not travel to the GPU box) on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import copy
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")     # the reference is fastest single-threaded

import numpy as np  # noqa: E402

METRIC = "EM-iterations/sec"
UNIT = "EM-iter/s"


def build_problem(cfgname, n_trials=None, verbose=False):
    """Segments + params + config after the reference's own pre-vem pipeline (vlgp/api.py:28-60), on the host for the
    oracle / as input of the device session.  Deterministic: np.random.seed(0) before the RNG-consuming steps."""
    from vlgp_b200.synth import CONFIGS, make_trials
    from vlgp_b200 import preprocess

    c = dict(CONFIGS[cfgname])
    if n_trials is not None:
        c["n_trials"] = n_trials
    trials = make_trials(c["n_trials"], c["T"], c["N"], c["L"], seed=0, latents=c.get("latents", "sine"))
    config = preprocess.get_config()
    params = preprocess.get_params(trials, c["L"], omega_bound=config["omega_bound"])
    np.random.seed(0)
    preprocess.initialize(trials, params, config)
    preprocess.fill_params(params)
    preprocess.fill_trials(trials)
    params.pop("transform", None)
    return trials, params, config, c


def cut(trials, params, config):
    from vlgp_b200.util import cut_trials
    from vlgp_b200.preprocess import fill_trials

    segs = list(cut_trials(trials, params, config))
    fill_trials(segs)
    return segs


# ----------------------------------------------------------------------------------------------------------------------
# CPU arm: NumPy/SciPy oracle port of the reference, bounded sample
# ----------------------------------------------------------------------------------------------------------------------
def cpu_em_iteration_time(cfgname, sample_trials, steps, warmup, all_threads_probe=None):
    """Seconds per EM iteration of the oracle on ``sample_trials`` trials of the workload, and the segment counts."""
    from oracle import vlgp_oracle as orc

    trials, params, config, c = build_problem(cfgname, n_trials=sample_trials)
    lengths = [t["y"].shape[0] for t in trials]
    params["cholesky"] = orc.make_cholesky(lengths, params["omega"], params["sigma"], params["rank"])
    orc.update_w(trials, params)
    orc.update_v(trials, params, config)
    segs = cut(trials, params, config)
    segs = [copy.deepcopy(s) for s in segs]
    params["cholesky"] = orc.make_cholesky([config["window"]], params["omega"], params["sigma"], params["rank"])
    config["max_iter"] = 1
    config["min_iter"] = 1
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        orc.vem(segs, params, config)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    if all_threads_probe is not None:
        # the reference at one BLAS thread and at one per core, the better one is the baseline.
        # One more iteration with the pool opened up (the operands are 50 x 50: more threads have never helped).
        try:
            from threadpoolctl import threadpool_limits

            ncpu = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
            with threadpool_limits(limits=ncpu, user_api="blas"):
                t0 = time.perf_counter()
                orc.vem(segs, params, config)
                all_threads_probe.update(threads=ncpu, sec=time.perf_counter() - t0)
        except Exception as e:  # pragma: no cover - diagnostics only
            all_threads_probe.update(error=repr(e))
    return float(np.mean(times)), len(segs), c


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    os.environ["VLGP_NO_FASTPACK"] = "1"      # the CPU arm loads none of the product's native code
    out_fd = _claim_stdout()
    from vlgp_b200.synth import CONFIGS

    c = CONFIGS[args.config]
    full_trials = c["n_trials"]
    sample = max(1, min(args.cpu_sample_trials, full_trials))
    # keep the whole --steps K --warmup W run within a few minutes: about 0.8-1.4 s per trial and EM iteration on one
    # host core for the named workloads, so the sample shrinks when many steps are asked for (linear extrapolation in
    # trials either way; `linearity` below evidences it)
    budget_s = 240.0
    per_trial_s = 1.4 * (c["N"] / 100.0) * ((c["T"] if isinstance(c["T"], int) else sum(c["T"]) / 2) / 1000.0)
    while sample > 2 and (args.steps + args.warmup + 2) * sample * per_trial_s > budget_s:
        sample -= 1
    probe = {}
    sec, nseg, _ = cpu_em_iteration_time(args.config, sample, args.steps, args.warmup, all_threads_probe=probe)
    # evidence for the linear extrapolation in the number of trials: one more iteration on half the sample
    half = max(1, sample // 2)
    sec_half, nseg_half, _ = cpu_em_iteration_time(args.config, half, 1, 0) if half < sample else (sec, nseg, None)
    threads = int(os.environ.get("OPENBLAS_NUM_THREADS", os.environ.get("OMP_NUM_THREADS", "0")) or 0) or os.cpu_count()
    sec_one = sec
    if probe.get("sec") and probe["sec"] < sec:     # the opened-up pool won: that is the baseline then
        sec, threads = probe["sec"], probe["threads"]
    full_sec = sec * full_trials / sample          # E-, M- and H-step cost are all linear in the number of segments
    value = 1.0 / full_sec
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": full_sec * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args.config, c, args.gpus),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": "%d of %d trials (%d of %d segments), %.2f s per EM iteration measured, scaled "
                                   "linearly in segments" % (sample, full_trials, nseg, nseg * full_trials // sample,
                                                             sec),
                         "linearity": {"trials": [half, sample], "segments": [nseg_half, nseg],
                                       "sec_per_em_iteration": [sec_half, sec_one],
                                       "sec_per_segment": [sec_half / nseg_half, sec_one / nseg]},
                         "blas_threads_tried": {"1": sec_one, str(probe.get("threads", "all")): probe.get("sec")},
                         "note": "seconds per EM iteration of the sample at each BLAS pool size; the reference is a "
                                 "single Python process, the faster setting is reported"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    _emit(out_fd, line)


def workload_config(name, c, gpus):
    T = c["T"]
    return {"workload": "%s: %d trials x T=%s x %d neurons x %d latents, Poisson, window=50 rank=50 Eniter=25 Mniter=25 "
                        "Hstep=True" % (name, c["n_trials"], T if isinstance(T, int) else "U[%d,%d]" % tuple(T), c["N"],
                                        c["L"]),
            "segments": c["n_trials"] * (T // 50) if isinstance(T, int) else "overlapping windows of 50 bins "
                        "(vlgp/util.py:482-498), processed with the reference's in-place semantics",
            "sharding": "trials over %d rank(s), sum-allreduce of M-/H-step statistics (in-kernel over NVLink peer memory; "
                        "NCCL when peers cannot be mapped)" % gpus,
            "l2": "256 MiB write between steps evicts the 126 MB L2 (inside the timed region, <0.1 ms/step)"}


# ----------------------------------------------------------------------------------------------------------------------
# clocks
# ----------------------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        self.nvml = []              # (sm MHz, max MHz, reasons bitmask) polled in-process every 10 ms
        self._stop = threading.Event()
        self._thr = None
        self._armed = threading.Event()

    def _poll(self):
        # nvidia-smi needs up to a second to start and then reports every 50 ms: too coarse for a 100 ms timed region
        # (8 GPUs).  NVML in a thread answers in ~50 us per query and releases the GIL while it does.
        try:
            import pynvml as nv

            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.gpu)
            mx = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
                nv.nvmlDeviceGetCurrentClocksThrottleReasons
            while not self._stop.is_set():
                if self._armed.is_set():
                    self.nvml.append((float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)), float(mx), int(reasons(h))))
                self._stop.wait(0.01)
        except Exception:      # noqa: BLE001 -- no NVML: the nvidia-smi samples remain
            pass

    def arm(self):
        """Samples count from here (the sampler itself is started earlier so that it is up by now)."""
        self._armed.set()
        try:
            self._mark = os.path.getsize(self.f.name)
        except OSError:
            self._mark = 0

    def start(self):
        self._mark = 0
        self._thr = threading.Thread(target=self._poll, daemon=True)
        self._thr.start()
        try:
            self.p = subprocess.Popen(["nvidia-smi", "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms",
                                       "50", "-i", str(self.gpu)], stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        self._stop.set()
        if self._thr is not None:
            self._thr.join(timeout=2)
        if self.p is not None:
            self.p.terminate()
            try:
                self.p.wait(timeout=5)
            except Exception:
                self.p.kill()
        self.f.flush()
        self.f.seek(self._mark)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for v, m, bits in self.nvml:
            sm.append(v)
            mx.append(m)
            for nm, bit in zip(names, (0x8, 0x40, 0x20, 0x4)):
                if bits & bit:
                    reasons.add(nm)
        for ln in self.f.read().splitlines():
            parts = [x.strip() for x in ln.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                mx.append(float(parts[2]))
            except ValueError:
                continue
            for nm, val in zip(names, parts[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        if sm:
            out = {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(np.max(mx)), "reasons": sorted(reasons),
                   "samples": len(sm)}
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        return out


# ----------------------------------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------------------------------
def estep_flops(S, W, N, L, ncols, n_iter, rank):
    """Algorithmic flops of one E-step launch (symmetric-aware minimum, c_exp = 20 flops per
    exp): per segment-iteration 8 W L N + 2 c_exp W N for the two rate passes, and per latent the r x r Gram
    (W r^2), its Cholesky (r^3/3), the solve + matvecs of the mean step (2 r^2 + 8 W r) and the W triangular solves of
    the variance (W r^2).  Returned for r = rank (what the reference executes) and for r = the number of non-zero
    columns of each latent's factor (what is mathematically required)."""
    def per_latent(r):
        return (W * r * r + r ** 3 / 3.0 + 2 * r * r + 8 * W * r) + (W * r * r)

    base = 8.0 * W * L * N + 2 * 20.0 * W * N
    full = S * n_iter * (base + L * per_latent(rank))
    eff = S * n_iter * (base + sum(per_latent(int(r)) for r in ncols))
    return full, eff


def _claim_stdout():
    """The contract is ONE JSON line on stdout.  Libraries loaded below may print there on their own (NCCL announces
    its version from C when a communicator is split): point file descriptor 1 at stderr for the rest of the process and
    return a private duplicate of the real stdout for the final line."""
    sys.stdout.flush()
    real = os.dup(1)
    os.dup2(2, 1)
    return real


def _emit(fd, line):
    os.write(fd, (json.dumps(line) + "\n").encode())


def solo_parity(args, traj, n_iter):
    """Correctness witness of a multi-rank run: rank 0 repeats the same n_iter EM iterations on ALL trials with a
    private single-rank context (no communicator) and reports the relative differences of the loading, bias, omega and
    the norm of the posterior means against what the N ranks arrived at after the timed steps."""
    from vlgp_b200 import core, engine
    from vlgp_b200.core import Session
    from vlgp_b200.gp import make_cholesky

    shared = engine._ENGINE
    solo = engine.Engine(shared.device)
    engine._ENGINE = solo
    quiet, stdout = open(os.devnull, "w"), sys.stdout
    sys.stdout = quiet
    try:
        trials, params, config, c = build_problem(args.config, n_trials=getattr(args, "n_trials", None))
        make_cholesky(trials, params, config)
        core.update_w(trials, params, config)
        core.update_v(trials, params, config)
        segs = cut(trials, params, config)
        make_cholesky(segs, params, config)
        config["max_iter"] = config["min_iter"] = 1
        with Session(segs, params) as s:
            for _ in range(n_iter):
                solo.flush_l2()
                core._em_iteration(s, segs, params, config)
                s.ts.norms()
            mu_sq = float(s.ts.norms()[0])
    finally:
        sys.stdout = stdout
        engine._ENGINE = shared
        solo.close()

    def rel(x, ref):
        return float(np.max(np.abs(np.asarray(x) - np.asarray(ref))) / max(np.max(np.abs(ref)), 1e-300))

    return {"vs": "single-rank run of the same %d EM iterations on rank 0 (private context, no communicator)" % n_iter,
            "rel_diff_a": rel(traj["a"], params["a"]), "rel_diff_b": rel(traj["b"], params["b"]),
            "rel_diff_omega": rel(traj["omega"], params["omega"]),
            "rel_diff_mu_norm": abs(np.sqrt(traj["mu_sq"]) - np.sqrt(mu_sq)) / np.sqrt(mu_sq)}


def run_ours(args):
    out_fd = _claim_stdout()
    from vlgp_b200 import core, dist
    from vlgp_b200.core import Session
    from vlgp_b200.gp import make_cholesky

    eng = dist.init_from_env()
    world, rank = dist.world_size(), dist.rank()
    if world != args.gpus:
        raise SystemExit("--gpus %d but WORLD_SIZE=%d (launch N>1 with torch.distributed.run)" % (args.gpus, world))

    # ---- problem: every rank builds the same global initial state, then keeps its shard of trials -------------------
    trials, params, config, c = build_problem(args.config, n_trials=getattr(args, "n_trials", None))
    if getattr(args, "dtype", "f64") == "f32":
        config["dtype"] = "float32"
    lo, hi = dist.shard_bounds(len(trials), world, rank)
    my_trials = trials[lo:hi]
    make_cholesky(my_trials, params, config)
    core.update_w(my_trials, params, config)
    core.update_v(my_trials, params, config)
    segs = cut(my_trials, params, config)
    make_cholesky(segs, params, config)
    config["max_iter"] = 1
    config["min_iter"] = 1
    W, N, L = config["window"], c["N"], c["L"]
    S_local = len(segs)
    S_total = int(sum(-(-tr["y"].shape[0] // W) for tr in trials))

    peaks = {}
    if rank == 0:
        peaks = eng.peak_fp64()
        peaks["hbm_gbs_copy"] = eng.peak_hbm(1 << 30)

    # ---- device-resident arm ---------------------------------------------------------------------------------------
    s = Session(segs, params)
    state0 = s.ts.get_state(("mu", "v", "w"))
    y_u8 = getattr(s.ts, "y_stored", None) == 1
    p0 = copy.deepcopy(params)
    quiet = open(os.devnull, "w")

    def step():
        eng.flush_l2()
        core._em_iteration(s, segs, params, config)
        s.ts.norms()                              # convergence bookkeeping of vem (vlgp/core.py:350-354)

    stdout = sys.stdout
    sys.stdout = quiet
    try:
        sampler = ClockSampler(eng.device)
        sampler.start()                          # before the warm-up: nvidia-smi / NVML are up when the timed region starts
        for _ in range(args.warmup):
            step()
        eng.profile_enable(0x1)                  # E-step launches only: they end with a sync anyway
        c0 = eng.counters()
        eng.sync()
        dist.barrier()
        sampler.arm()                            # only samples from here on are reported
        eng.timer_start()
        t0 = time.perf_counter()
        split = np.zeros(3)
        for _ in range(args.steps):
            eng.flush_l2()
            split += core._em_iteration(s, segs, params, config)
            s.ts.norms()
        ms = eng.timer_stop()
        wall = time.perf_counter() - t0
        dist.barrier()
        clocks = sampler.stop()
        c1 = eng.counters()
        e_ms, e_n = eng.profile_get(0)
        eng.profile_enable(0)
    finally:
        sys.stdout = stdout
    ms = float(eng.allreduce(np.array([ms]), op="max")[0])
    launches = c1["launches"] - c0["launches"]
    # state after the timed steps, for the multi-rank correctness witness below
    traj = {"a": np.array(params["a"], copy=True), "b": np.array(params["b"], copy=True),
            "omega": np.array(params["omega"], copy=True), "mu_sq": float(s.ts.norms()[0])}
    # one more EM iteration outside the timed region with the H-step segment kernel and the M-step statistics kernel
    # timed individually (each timed launch adds a synchronisation, so this is kept out of `value`)
    sys.stdout = quiet
    try:
        nf0 = sum(sum(x) for x in config.get("hstep_nfev", []))
        eng.profile_enable(0x6)
        config["overlap_mh"] = False              # time the two kernels alone, not while they share the GPU
        core._em_iteration(s, segs, params, config)
        config.pop("overlap_mh")
        h_ms, h_n = eng.profile_get(2)
        m_ms, m_n = eng.profile_get(1)
        eng.profile_enable(0)
        h_evals = sum(sum(x) for x in config.get("hstep_nfev", [])) - nf0
    finally:
        sys.stdout = stdout
    ncols = [int((np.abs(params["cholesky"][W][l]).sum(axis=0) > 0).sum()) for l in range(L)]
    nfev = config.get("hstep_nfev", [])

    # ---- end-to-end arm: public vem() with host buffers ------------------------------------------------------------
    sys.stdout = quiet
    try:
        # the caller's side of fit() (vlgp/api.py:52-60): trials own their arrays, the segments handed to vem() are the
        # VIEWS util.cut_trials returns.  (A deep copy of the segment list would turn them into 5120 x 4 scattered
        # little arrays, which no caller of the reference's API produces and which the host gather pays ~6 ms for.)
        e2e_params = copy.deepcopy(p0)
        try:
            e2e_trials = copy.deepcopy(my_trials)      # state as the resident arm started from (update_w / update_v)
            e2e_segs = cut(e2e_trials, e2e_params, config)
            if any(sg["y"].shape[0] != W for sg in e2e_segs):
                raise ValueError("unexpected segment length")
        except Exception:      # noqa: BLE001 -- keep the line: independent copies of the resident arm's segments
            e2e_segs = copy.deepcopy(segs)
            for i, sg in enumerate(e2e_segs):
                for k in ("mu", "v", "w"):
                    sg[k][...] = state0[k][i * W:(i + 1) * W]
        e2e_steps = max(1, min(args.steps, 10))
        core.vem(e2e_segs, e2e_params, config)     # warm-up
        dist.barrier()
        eng.sync()
        h2d = d2h = 0
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            sess = Session(e2e_segs, e2e_params)
            core.vem(e2e_segs, e2e_params, config, session=sess)
            h2d, d2h = sess.ts.h2d_bytes, sess.ts.d2h_bytes
            sess.close()
        eng.sync()
        e2e_s = (time.perf_counter() - t0) / e2e_steps
    finally:
        sys.stdout = stdout
    e2e_s = float(eng.allreduce(np.array([e2e_s]), op="max")[0])
    s.close()

    if rank != 0:
        return
    parity = None
    if world > 1:
        parity = solo_parity(args, traj, args.warmup + args.steps)
    ms_per_step = ms / args.steps
    value = 1e3 / ms_per_step
    f_full, f_eff = estep_flops(S_local, W, N, L, ncols, config["Eniter"], params["rank"])
    e_avg_ms = e_ms / max(e_n, 1)
    peak = peaks.get("dfma_tflops") or None
    mp = {}
    try:
        mp = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    traffic = None
    try:        # DRAM bytes per E-step launch from the committed `ncu --set full` capture of the same workload
        if args.config == "config2" and args.gpus == 1 and getattr(args, "n_trials", None) is None:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "estep_traffic.json")))["dram_bytes_per_launch"]
    except Exception:
        pass
    roof = {"bound": "fp64", "kernel": "estep (all %d Newton iterations of every segment in one launch)" % config["Eniter"],
            "achieved": f_eff / (e_avg_ms * 1e-3) / 1e12 if e_avg_ms else None,
            "achieved_reference_flops": f_full / (e_avg_ms * 1e-3) / 1e12 if e_avg_ms else None,
            "peak": peak, "unit": "TFLOP/s", "peak_source": "measured in this run: register-resident DFMA loop "
            "(vlgp_peak_fp64); mma.sync.m8n8k4.f64 gives %.1f" % peaks.get("dmma_tflops", float("nan")),
            "frac": (f_eff / (e_avg_ms * 1e-3) / 1e12 / peak) if (e_avg_ms and peak) else None,
            "traffic": traffic, "ms_per_launch": e_avg_ms, "launches_timed": e_n,
            "share_of_step": e_avg_ms / ms_per_step if ms_per_step else None,
            "factor_columns": ncols,
            "hbm": {"peak_copy_gbs_this_run": peaks.get("hbm_gbs_copy"), "peak_measured_json": mp.get("hbm_gbs")}}
    # the same launch against the HBM roofline (one read of y, read + write of mu, v, w, dmu):
    # three orders of magnitude below the copy peak, i.e. this kernel is not bandwidth-bound
    e_bytes = float(S_local * W) * (N * (1 if y_u8 else 8) + 7 * L * 8)
    hbm_peak = mp.get("hbm_gbs") or peaks.get("hbm_gbs_copy")
    roof["hbm"].update({"algorithmic_bytes_per_launch": e_bytes,
                        "achieved_gbs": e_bytes / (e_avg_ms * 1e-3) / 1e9 if e_avg_ms else None,
                        "peak_gbs": hbm_peak,
                        "peak_source": "MEASURED_PEAKS.json" if mp.get("hbm_gbs") else "device copy measured in this run",
                        "frac": (e_bytes / (e_avg_ms * 1e-3) / 1e9 / hbm_peak) if (e_avg_ms and hbm_peak) else None})
    # secondary rooflines: the H-step per-segment kernel runs on the FP64 tensor pipe (mma.sync.m8n8k4.f64), the M-step
    # statistics kernel on the FP64 pipe; neither is HBM-bound (their GB/s are listed for completeness)
    nb = (W + 7) // 8
    if 1 <= W - 8 * (nb - 1) <= 2 and 3 <= nb <= 7 and not os.environ.get("VLGP_HSTEP_NO_SCHUR"):
        nb -= 1                                                    # bordered kernel: the sweep runs on the 8 (nb - 1) core rows
    dmma_per_seg = ((nb - 1) + (nb - 1) * nb // 2) * nb * 2      # per pivot block: nb-1 panel + nb(nb-1)/2 update tile products
    h_flops = 2.0 * 256.0 * dmma_per_seg * S_local * max(h_evals, 1)
    # useful work of one segment-evaluation: the W x W symmetric inverse (W^3 flops) plus d K d, the trace and the dK
    # contraction (4 W^2); the DMMA count above also multiplies the padding of W up to a multiple of 8
    h_useful = (float(W) ** 3 + 4.0 * W * W) * S_local * max(h_evals, 1)
    h_peak = peaks.get("dmma_tflops")
    roof_h = {"bound": "tensor", "kernel": "hstep_segment_schur / hstep_segment_dmma (FP64 tensor pipe, DMMA)",
              "achieved": h_useful / (h_ms * 1e-3) / 1e12 if h_ms else None, "peak": h_peak,
              "unit": "TFLOP/s", "frac": (h_useful / (h_ms * 1e-3) / 1e12 / h_peak) if (h_ms and h_peak) else None,
              "flops": "useful: W^3 + 4 W^2 per segment-evaluation (symmetric inverse, d K d, trace, dK contraction)",
              "frac_issued_dmma": (h_flops / (h_ms * 1e-3) / 1e12 / h_peak) if (h_ms and h_peak) else None,
              "peak_source": "measured in this run: mma.sync.m8n8k4.f64 loop (vlgp_peak_fp64)",
              "ms_total": h_ms, "launches_timed": h_n, "evaluations": h_evals, "dmma_per_segment_evaluation": dmma_per_seg}
    m_flops = float(S_local * W) * N * (10 * L + L * L + 20 + 4)
    m_bytes = float(S_local * W) * (N * 1 + 2 * L * 8)
    roof_m = {"bound": "fp64", "kernel": "mstep_stats_tma (TMA-staged (mu, v) rows / count tile; average over the Newton "
              "iterations of one M-step)", "achieved": m_flops * m_n / (m_ms * 1e-3) / 1e12 if m_ms else None,
              "peak": peak, "unit": "TFLOP/s", "frac": (m_flops * m_n / (m_ms * 1e-3) / 1e12 / peak) if (m_ms and peak) else None,
              "ms_per_launch": m_ms / max(m_n, 1), "launches_timed": m_n,
              "hbm_view": {"algorithmic_GBps": m_bytes * m_n / (m_ms * 1e-3) / 1e9 if m_ms else None,
                           "peak_copy_GBps": peaks.get("hbm_gbs_copy")}}
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64" if getattr(args, "dtype", "f64") == "f64" else "f32 (E-step rate passes; Gram / inverse / mean step, M- and H-step f64)",
        "data": "synthetic", "config": workload_config(args.config, c, args.gpus),
        "split_ms": {"estep": split[0] / args.steps * 1e3, "mstep": split[1] / args.steps * 1e3,
                     "hstep": split[2] / args.steps * 1e3, "wall_per_step": wall / args.steps * 1e3,
                     "note": "the M-step runs on a second stream under the host-driven H-step; 'mstep' is the host time "
                             "it costs on top (enqueue + final wait), 'hstep' the wall time of the H-step rounds"},
        "hstep_evals_per_step": float(np.mean([sum(x) for x in nfev])) if nfev else None,
        "solves_per_sec": (2.0 * S_total * L * config["Eniter"]) / (ms_per_step * 1e-3),
        "clocks": clocks, "gpu_launches": int(launches), "roofline": roof, "roofline_hstep": roof_h,
        "roofline_mstep": roof_m,
        "e2e": {"value": 1.0 / e2e_s, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "ms_per_step": e2e_s * 1e3, "api": "vlgp_b200.core.vem(splits, params, config) with host ndarrays; splits = util.cut_trials(trials) "
                       "views, as in fit()"},
    }
    if parity is not None:
        line["parity"] = parity
    if not args.no_cpu and args.gpus == 1:
        sec, nseg, _ = cpu_em_iteration_time(args.config, args.cpu_sample_trials, 1, 0)
        full_sec = sec * c["n_trials"] / args.cpu_sample_trials
        line["cpu_baseline"] = {"value": 1.0 / full_sec, "unit": UNIT, "cores": 1, "kind": "port",
                                "sample": "%d of %d trials (%d segments), %.2f s per EM iteration measured with the "
                                          "NumPy/SciPy oracle, scaled linearly in segments" % (
                                              args.cpu_sample_trials, c["n_trials"], nseg, sec)}
    _emit(out_fd, line)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="config2")
    ap.add_argument("--cpu-sample-trials", type=int, default=12)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--dtype", default="f64", choices=["f64", "f32"],
                    help="f32: config['dtype'] = 'float32' (single-precision E-step rate passes; BASELINE configs[2])")
    ap.add_argument("--n-trials", type=int, default=None,
                    help="experiments only: fewer trials than the named workload (the line's config then says so)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
