"""Marginalization leg alone (for ncu): upload, solve, then gf2_marginalize a few times."""
import sys, os, importlib
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
from gf2_loader import load
gf2 = load(); synth = importlib.import_module("gf2_b200.synth")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
base = synth.make_windows(64, n_landmarks=1000, prior_stride=80)
w = {k: (np.concatenate([v] * ((n + 63) // 64))[:n] if isinstance(v, np.ndarray) and v.ndim >= 1 and v.shape[0] == 64 and k != "imu_noise" else v) for k, v in base.items()}
opts = gf2.abi.default_opts()
s = gf2.Solver(n, 11, w["max_landmarks"], w["max_obs"], max_imu_samples=w["n_imu_samples"], max_prior_rows=80)
s.upload(w, preintegrate="device"); s.solve(opts, n)
for _ in range(reps):
    s.set_prior(w)
    st, m = s.marginalize(opts, 0, n)
    print("marginalize ms", s.last_marginalize_ms(), "ok", (st == 0).mean(), "m", m[0])
