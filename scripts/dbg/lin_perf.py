"""Dev helper: per-kernel phase times of an 8-iteration solve of B W10-F1000 windows (the bench's resident leg without the rest).
usage: python scripts/dbg/lin_perf.py [B] [reps] [sweep]   (sweep: 0 auto, 1 batch kernel, 2 window kernel; GF2_LIB=<path> selects a variant build of libgf2_b200.so)"""
import sys, os, importlib, shutil, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
if os.environ.get("GF2_LIB"):
    shutil.copy(os.environ["GF2_LIB"], os.path.join(ROOT, "ground-fusion2_b200", "libgf2_b200.so"))
from gf2_loader import load
gf2 = load(); synth = importlib.import_module("gf2_b200.synth")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
sweep = int(sys.argv[3]) if len(sys.argv) > 3 else 0
distinct = 16
base = synth.make_windows(distinct, n_landmarks=1000, prior_stride=8)
w = {k: (np.concatenate([v] * (B // distinct + 1))[:B] if isinstance(v, np.ndarray) and v.ndim >= 1 and v.shape[0] == distinct and k != "imu_noise" else v) for k, v in base.items()}
s = gf2.Solver(B, w["n_frames"], w["max_landmarks"], w["max_obs"], max_imu_samples=w["n_imu_samples"], max_prior_rows=8, sweep=sweep)
s.upload(w, preintegrate="device"); s.snapshot(B)
opts = gf2.abi.default_opts()
acc = {}
for r in range(reps + 1):
    s.restore(B); summ = s.solve(opts, B)
    t = s.last_timing()
    if r: 
        for k, v in t.items(): acc[k] = acc.get(k, 0.0) + v / reps
nl = acc["linearize_launches"]
print(f"{os.environ.get('GF2_LIB', 'default')} sweep={sweep}: B={B} lin {acc['linearize_ms'] / nl:.4f} ms/launch ({nl:.0f})  solve {acc['solve_ms'] / nl:.4f}  step {acc['step_ms'] / nl:.4f}  total {acc['total_ms']:.3f} ms  "
      f"iters {summ['iterations'].mean():.2f} cost ratio {np.median(summ['final_cost'] / summ['initial_cost']):.3e}")
