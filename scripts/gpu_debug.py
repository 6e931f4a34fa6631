"""Scratch parity probe (run on the GPU box): device path vs oracle on a few windows, verbose."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
import numpy as np
from gf2_loader import load
gf2 = load()
from importlib import import_module
synth = import_module("gf2_b200.synth")
import gf2_oracle as orc
abi = gf2.abi

nl = int(sys.argv[1]) if len(sys.argv) > 1 else 200
nw = int(sys.argv[2]) if len(sys.argv) > 2 else 3
prior = sys.argv[3] if len(sys.argv) > 3 else "anchor"
w = synth.make_windows(nw, n_landmarks=nl, prior=prior)
print("devices", gf2.device_count())
s = gf2.Solver(nw, 11, w["max_landmarks"], w["max_obs"], max_imu_samples=w["n_imu_samples"])
s.upload(w, preintegrate="device")
rec_d = s.get_imu(nw)
rec_o = orc.imu_preintegrate(w)
for f in ("sum_dt", "delta_p", "delta_q", "delta_v", "jacobian", "covariance"):
    a, b = rec_d[f], rec_o[f]
    print("preint", f, "max rel diff", np.abs(a - b).max() / max(1e-300, np.abs(b).max()))
opts = abi.default_opts()
s.set_imu(rec_o)  # identical records for the solver parity below
S, g, cost = s.linearize(opts, nw)
for i in range(min(nw, 2)):
    So, go, co, ete, etr = orc.linearize_window(w, i, opts)
    print("win", i, "D", So.shape, "cost gpu", cost[i], "oracle", co, "rel", abs(cost[i] - co) / co)
    print("  S max rel diff", np.abs(S[i] - So).max() / np.abs(So).max(), " g max rel diff", np.abs(g[i] - go).max() / np.abs(go).max())
    d = np.abs(S[i] - So); r, c = np.unravel_index(d.argmax(), d.shape); print("  worst at", r, c, S[i][r, c], So[r, c])
    # per-block diagnostics
    bd = np.zeros((11, 11))
    for a in range(11):
        for b in range(11):
            bd[a, b] = np.abs(S[i][15*a:15*a+15, 15*b:15*b+15] - So[15*a:15*a+15, 15*b:15*b+15]).max()
    np.set_printoptions(linewidth=200, precision=2)
    print(bd / np.abs(So).max())
# full solve
wo = {k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in w.items()}
t = time.time(); so = orc.solve_batch(wo, opts, n_threads=4); print("oracle solve s", time.time() - t)
s.set_states(w); s.set_landmarks(w)
t = time.time(); sg = s.solve(opts, nw); print("gpu solve s", time.time() - t, s.last_timing())
out = s.get_states(nw); lam = s.get_landmarks(nw)
print("oracle summaries", so)
print("gpu summaries   ", sg)
for i in range(nw):
    dp = np.abs(out["para_pose"][i] - wo["para_pose"][i]).max(); ds = np.abs(out["para_speedbias"][i] - wo["para_speedbias"][i]).max()
    dl = np.abs(lam[i] - wo["inv_depth"][i]).max()
    print("win", i, "pose diff", dp, "speedbias diff", ds, "invdepth diff", dl, " (moved by", np.abs(wo["para_pose"][i] - w["para_pose"][i]).max(), ")")
