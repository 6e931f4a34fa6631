"""Debug helper: where does the device reduced system differ from the oracle's? (per frame-block maxima)"""
import sys, importlib, numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "oracle")
from gf2_loader import load
gf2 = load(); synth = importlib.import_module("gf2_b200.synth"); import gf2_oracle as orc
nl = int(sys.argv[1]) if len(sys.argv) > 1 else 200
w = synth.make_windows(1, n_landmarks=nl, prior="anchor", sorted_landmarks=True)
orc.imu_preintegrate(w)
s = gf2.Solver(1, w["n_frames"], w["max_landmarks"], w["max_obs"], max_imu_samples=w["n_imu_samples"])
s.upload(w, preintegrate="records")
opts = gf2.abi.default_opts()
S, g, cost = s.linearize(opts, 1)
So, go, co, _, _ = orc.linearize_window(w, 0, opts)
E = np.abs(S[0] - So)
np.set_printoptions(linewidth=250, precision=2)
print("cost", cost[0], co)
print("pose-pose block error maxima (rows i, cols j) relative to the block's max")
B = np.zeros((11, 11)); R = np.zeros((11, 11))
for i in range(11):
    for j in range(11):
        B[i, j] = E[15*i:15*i+6, 15*j:15*j+6].max(); R[i, j] = B[i, j] / max(np.abs(So[15*i:15*i+6, 15*j:15*j+6]).max(), 1e-300)
print(R)
print("g err", np.abs(g[0]-go).reshape(11,15)[:, :6].max(axis=1) / np.abs(go).max())
