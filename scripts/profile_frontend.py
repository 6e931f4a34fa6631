"""Driver for the ncu captures of the front-end and LIO kernels (k_clahe_*, k_gftt_*, k_lk, k_lio_factors): one batched CLAHE + detect +
track call over 64 streams and one LIO scan, after a warm-up call of each. Run under
  ncu --set full --clock-control none --import-source on -k regex:"k_gftt|k_clahe|k_lio" -s <warm-up launches> -c <n> python scripts/profile_frontend.py"""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gf2_loader import load  # noqa: E402

gf2 = load()
synth = importlib.import_module("gf2_b200.synth")
S = 64
base = [synth.image_pair(s, shift=(2.0 + 0.3 * s, -1.0)) for s in range(8)]
prev = np.stack([base[s % 8][0] for s in range(S)]); cur = np.stack([base[s % 8][1] for s in range(S)])
npts = min(len(b[2]) for b in base)
pts = np.stack([base[s % 8][2][:npts] for s in range(S)])
t = gf2.Tracker(640, 480, max_pts=npts, max_streams=S)
t.set_equalize(40.0, (8, 8))
mask = np.full((S, 480, 640), 255, np.uint8); mask[:, 100:300, 200:500] = 0
scene = synth.lio_scene(7, n_map_points=100000, n_keypoints=3000)
o = gf2.abi.default_lio_opts(translation_begin=scene["translation_begin"], rotation=scene["rotation"], translation=scene["translation"])
h = gf2.Lio(max_voxels=len(scene["keys"]), max_keypoints=len(scene["keypoints"]), max_points_per_voxel=scene["max_points_per_voxel"])
h.set_map(scene["keys"], scene["n_points"], scene["points"])
for rep in range(2):   # pass 0 = warm-up (4 CLAHE + 3 detector + LK launches + 1 LIO launch per pass)
    t.track_fb(prev, cur, pts)
    corners = t.detect(None, 40, mask=mask, n_streams=S)
    fac, _, _, _ = h.build_factors(scene["keypoints"], o)
    print("pass", rep, "detect_ms", t.last_timing()["detect_ms"], "corners", sum(len(c) for c in corners), "lio kernel_ms", h.last_timing()["kernel_ms"], "residuals", len(fac))
