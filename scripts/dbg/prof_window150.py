"""Dev helper for ncu: launch list of one-window solves of a 150-landmark window (the single-robot shape)."""
import sys, os, importlib, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from gf2_loader import load
gf2 = load(); synth = importlib.import_module("gf2_b200.synth")
w = synth.make_windows(1, n_landmarks=150, prior_stride=80)
s = gf2.Solver(1, w["n_frames"], w["max_landmarks"], w["max_obs"], max_imu_samples=w["n_imu_samples"], max_prior_rows=80)
s.upload(w, preintegrate="device"); s.snapshot(1)
opts = gf2.abi.default_opts()
for _ in range(3):
    s.restore(1); s.set_prior(w); s.solve(opts, 1); st, m = s.marginalize(opts, 0, 1)
    print(s.last_timing(), s.last_marginalize_ms(), st)
