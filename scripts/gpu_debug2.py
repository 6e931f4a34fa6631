"""Scratch: find the first trust-region iteration where the GPU path and the oracle diverge."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
import numpy as np
from gf2_loader import load
gf2 = load()
from importlib import import_module
synth = import_module("gf2_b200.synth")
import gf2_oracle as orc
nl = int(sys.argv[1]); prior = sys.argv[2]; n = 3
w = synth.make_windows(n, n_landmarks=nl, prior=prior)
orc.imu_preintegrate(w)
s = gf2.Solver(n, 11, w["max_landmarks"], w["max_obs"], max_imu_samples=w["n_imu_samples"])
np.set_printoptions(linewidth=220, precision=6)
for it in range(1, 9):
    opts = gf2.abi.default_opts(max_iterations=it)
    s.upload(w, preintegrate="records")
    sg = s.solve(opts, n)
    got = s.get_states(n)
    wo = {k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in w.items()}
    so = orc.solve_batch(wo, opts, n_threads=3)
    print("it", it, "gpu", [(x["iterations"], x["successful_steps"], x["termination"], float(x["final_cost"])) for x in sg])
    print("      orc", [(x["iterations"], x["successful_steps"], x["termination"], float(x["final_cost"])) for x in so],
          "pose diff", np.abs(got["para_pose"] - wo["para_pose"]).max())
summ, tr = orc.solve_window_trace({k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in w.items()}, 0, gf2.abi.default_opts())
print(tr)
opts = gf2.abi.default_opts(); s.upload(w, preintegrate="records"); s.solve(opts, n); print(s.get_trace(n)[0, :8])
