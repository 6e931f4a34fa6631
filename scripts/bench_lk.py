"""Front-end benchmark (BASELINE.json config 3): pyramidal LK forward + backward check, 640x480, 4 levels (maxLevel 3),
300 corners, batch of independent 30 Hz synthetic streams on one B200; cv2 (the reference's own library) timed beside it.
Prints one JSON line. Not the driver's bench (bench.py is the solver metric)."""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
import numpy as np
from gf2_loader import load
import lk_oracle as lk

ap = argparse.ArgumentParser(); ap.add_argument("--streams", type=int, default=64); ap.add_argument("--steps", type=int, default=20)
args = ap.parse_args()
gf2 = load()
S = args.streams
base = [lk.synthetic_pair(s % 8, shift=(2.0 + 0.3 * (s % 8), -1.0)) for s in range(min(S, 8))]
prev = np.stack([base[s % len(base)][0] for s in range(S)]); cur = np.stack([base[s % len(base)][1] for s in range(S)])
npts = min(len(b[2]) for b in base)
pts = np.stack([base[s % len(base)][2][:npts] for s in range(S)])
t = gf2.Tracker(640, 480, max_pts=npts, max_streams=S)
for _ in range(3):
    t.track_fb(prev, cur, pts)
lk_ms = tot_ms = 0.0
t0 = time.perf_counter()
for _ in range(args.steps):
    out, ok = t.track_fb(prev, cur, pts)
    tm = t.last_timing(); lk_ms += tm["lk_ms"]; tot_ms += tm["total_ms"]
wall = time.perf_counter() - t0
# algorithmic bytes of one frame pair
bytes_lk = 640 * 480 * (1 + 2 * (0.25 + 0.0625 + 0.015625)) + npts * 6 * ((21 + 2) ** 2 + (21 + 8) ** 2)
line = {"metric": "LK frame pairs/sec (640x480, 4 levels, fwd+bwd)", "value": S * args.steps / wall, "unit": "frame pairs/s", "streams": S, "points": npts,
        "device_ms_per_batch": tot_ms / args.steps, "lk_kernel_ms_per_batch": lk_ms / args.steps, "tracked_fraction": float(ok.mean()),
        "roofline": {"bound": "hbm", "achieved_GBps": bytes_lk * S / (lk_ms / args.steps / 1e3) / 1e9, "algorithmic_bytes_per_pair": bytes_lk}}
try:
    import cv2
    crit = (cv2.TERM_CRITERIA_COUNT + cv2.TERM_CRITERIA_EPS, 30, 0.01)
    for threads in (1, 0):
        cv2.setNumThreads(threads)
        t0 = time.perf_counter(); n = 0
        while time.perf_counter() - t0 < 3.0:
            s = n % len(base)
            c, st, _ = cv2.calcOpticalFlowPyrLK(base[s][0], base[s][1], base[s][2][:npts].reshape(-1, 1, 2), None, winSize=(21, 21), maxLevel=3, criteria=crit)
            cv2.calcOpticalFlowPyrLK(base[s][1], base[s][0], c, base[s][2][:npts].reshape(-1, 1, 2).copy(), winSize=(21, 21), maxLevel=1, criteria=crit, flags=cv2.OPTFLOW_USE_INITIAL_FLOW)
            n += 1
        line[f"cv2_threads_{threads or 'all'}"] = n / (time.perf_counter() - t0)
except ImportError:
    pass
print(json.dumps(line))
