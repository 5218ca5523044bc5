"""Small driver for ncu: config-2 problem, a few EM iterations on the device-resident session (no timing here)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from vlgp_b200 import core
from vlgp_b200.core import Session
from vlgp_b200.gp import make_cholesky

n_iter = int(sys.argv[1]) if len(sys.argv) > 1 else 2
trials, params, config, c = bench.build_problem("config2")
make_cholesky(trials, params, config)
core.update_w(trials, params, config)
core.update_v(trials, params, config)
segs = bench.cut(trials, params, config)
make_cholesky(segs, params, config)
config["max_iter"] = config["min_iter"] = 1
sys.stdout = open(os.devnull, "w")
with Session(segs, params) as s:
    for _ in range(n_iter):
        core._em_iteration(s, segs, params, config)
sys.stdout = sys.__stdout__
print("done", params["omega"])
