"""Dev helper for ncu: launch list of config-4 solves (W10-F1000 + wheel + 5,000 planes), B windows."""
import sys, os, importlib, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from gf2_loader import load
gf2 = load(); synth = importlib.import_module("gf2_b200.synth")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1184
base = synth.make_windows(8, config_id=4, n_landmarks=1000, wheel=True, n_planes=5000)
w = {k: (np.concatenate([v] * (B // 8 + 1))[:B] if isinstance(v, np.ndarray) and v.ndim >= 1 and v.shape[0] == 8 and k not in ("imu_noise", "wheel_noise") else v) for k, v in base.items()}
s = gf2.Solver(B, w["n_frames"], w["max_landmarks"], w["max_obs"], max_planes=w["max_planes"], max_imu_samples=w["n_imu_samples"], use_wheel=True, max_wheel_samples=w["n_wheel_samples"])
s.upload(w, preintegrate="device"); s.snapshot(B)
opts = gf2.abi.default_opts()
for _ in range(2):
    s.restore(B); s.solve(opts, B); print(s.last_timing())
