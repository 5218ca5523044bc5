"""bench.py's own arm needs a GPU; its FLOW does not.  With the oracle stand-in as the engine (tests/oracle_engine.py)
the whole of run_ours -- problem set-up, warm-up, timed loop, clock sampler, the extra profiling iteration, the
end-to-end arm, the JSON assembly -- runs on the CPU in a child process, and what arrives on the child's stdout must be
exactly ONE line of JSON with the keys of the contract (numbers are meaningless here; only their presence is checked)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

DRIVER = r"""
import sys, types
sys.path.insert(0, {root!r}); sys.path.insert(0, {tests!r})
import oracle_engine, vlgp_b200.engine as engine_mod
engine_mod._ENGINE = oracle_engine.OracleEngine()
print("a library chatting on stdout before the benchmark starts is the caller's problem")   # before _claim_stdout
import bench
args = types.SimpleNamespace(gpus=1, steps=2, warmup=3, impl="ours", config="tiny", cpu_sample_trials=1, no_cpu=True)
from vlgp_b200 import synth
synth.CONFIGS["tiny"] = dict(n_trials=2, T=100, N=6, L=2, dtype="f64")
import os as _os
# a C-level write to fd 1 during the run (what NCCL does when it announces its version) must not reach stdout
real_claim = bench._claim_stdout
def claim_then_chat():
    fd = real_claim()
    _os.write(1, b"NCCL version 9.9.9+fake\n")
    return fd
bench._claim_stdout = claim_then_chat
bench.run_ours(args)
"""


def test_run_ours_prints_exactly_one_json_line_with_the_contract_keys(tmp_path):
    script = tmp_path / "drive_bench.py"
    script.write_text(DRIVER.format(root=ROOT, tests=os.path.join(ROOT, "tests")))
    env = dict(os.environ, OPENBLAS_NUM_THREADS="1")
    env.pop("WORLD_SIZE", None)
    r = subprocess.run([sys.executable, str(script)], capture_output=True, text=True, timeout=600, cwd=str(tmp_path), env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    # the first line was printed by the driver before bench took stdout over; after that: the JSON line and nothing else
    assert len(lines) == 2 and lines[0].startswith("a library chatting"), lines
    assert "NCCL version 9.9.9+fake" in r.stderr and "NCCL version" not in r.stdout.split("\n", 1)[1]
    d = json.loads(lines[1])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "roofline", "e2e", "clocks", "gpu_launches", "split_ms"):
        assert k in d, k
    assert d["metric"] == "EM-iterations/sec" and d["n_gpus"] == 1 and d["steps"] == 2 and d["warmup"] == 3
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["e2e"]["value"] > 0
    assert set(d["roofline"]) >= {"bound", "achieved", "peak", "unit", "frac", "traffic"}
    # the dominant kernel is FP64-bound; the same launch against the HBM roofline (MEASURED_PEAKS.json when present)
    assert set(d["roofline"]["hbm"]) >= {"algorithmic_bytes_per_launch", "achieved_gbs", "peak_gbs", "peak_source", "frac"}
    assert set(d["e2e"]) >= {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"}
    assert set(d["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}        # no GPU here: the values are None
    assert "workload" in d["config"] and "cpu_baseline" not in d            # --no-cpu


REF_DRIVER = r"""
import sys, types
sys.path.insert(0, {root!r})
import bench
from vlgp_b200 import synth
synth.CONFIGS["tiny"] = dict(n_trials=3, T=100, N=6, L=2, dtype="f64")
args = types.SimpleNamespace(gpus=1, steps=1, warmup=0, impl="reference", config="tiny", cpu_sample_trials=2, no_cpu=False)
bench.run_reference(args)
"""


def test_reference_arm_prints_one_json_line_with_its_keys(tmp_path):
    """`bench.py --impl reference` needs no GPU at all: the oracle port on a bounded sample, scaled to the workload; the
    line carries impl / cpu_baseline (kind, cores, sample, both BLAS pool sizes tried) / e2e with zero copy bytes."""
    script = tmp_path / "drive_ref.py"
    script.write_text(REF_DRIVER.format(root=ROOT))
    env = dict(os.environ, OPENBLAS_NUM_THREADS="1", RANK="0")
    r = subprocess.run([sys.executable, str(script)], capture_output=True, text=True, timeout=600, cwd=str(tmp_path), env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "EM-iterations/sec" and d["gpu_launches"] == 0
    assert d["value"] > 0 and d["higher_is_better"] is True and d["vs_baseline"] is None
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and "2 of 3 trials" in cb["sample"] and cb["value"] == d["value"]
    assert set(cb["blas_threads_tried"]) >= {"1"} and len(cb["blas_threads_tried"]) == 2
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # ranks other than 0 print nothing and exit 0
    r2 = subprocess.run([sys.executable, str(script)], capture_output=True, text=True, timeout=120, cwd=str(tmp_path),
                        env=dict(env, RANK="1"))
    assert r2.returncode == 0 and r2.stdout.strip() == ""
