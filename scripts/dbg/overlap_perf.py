"""Dev helper: does splitting a resident batch over K handles / streams / host threads (kernels of different chunks overlap on the SMs) beat one
handle? usage: python scripts/dbg/overlap_perf.py [B] [K...]"""
import sys, os, importlib, time, numpy as np
from concurrent.futures import ThreadPoolExecutor
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from gf2_loader import load
gf2 = load(); synth = importlib.import_module("gf2_b200.synth")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
Ks = [int(a) for a in sys.argv[2:]] or [1, 2, 4]
distinct = 16
base = synth.make_windows(distinct, n_landmarks=1000, prior_stride=8)
def tile(n):
    return {k: (np.concatenate([v] * (n // distinct + 1))[:n] if isinstance(v, np.ndarray) and v.ndim >= 1 and v.shape[0] == distinct and k != "imu_noise" else v) for k, v in base.items()}
opts = gf2.abi.default_opts()
for K in Ks:
    n = B // K
    w = tile(n)
    ss = []
    for c in range(K):
        s = gf2.Solver(n, w["n_frames"], w["max_landmarks"], w["max_obs"], max_imu_samples=w["n_imu_samples"], max_prior_rows=8)
        s.upload(w, preintegrate="device"); s.snapshot(n); ss.append(s)
    pool = ThreadPoolExecutor(K)
    def step(s):
        s.restore(n); s.imu_preintegrate_resident(w["imu_noise"], n); s.solve(opts, n)
    for _ in range(2): list(pool.map(step, ss))
    t0 = time.perf_counter(); reps = 4
    for _ in range(reps): list(pool.map(step, ss))
    dt = (time.perf_counter() - t0) / reps
    print(f"K={K}: {1e3 * dt:.2f} ms per {n * K} windows -> {n * K / dt:.0f} solves/s")
    for s in ss: s.close()
    pool.shutdown()
