"""Generates tests/golden/clahe_golden.npz: outputs of cv2.createCLAHE().apply (the call of VE/rosNodeTest.cpp:271-276) on images of
tests/golden/gftt_golden.npz (textured, saturated / flat regions that trigger the clip + residual redistribution) and on a
low-contrast image; stored as every 8th row + checksums (bit-exact comparison). Run once here; committed for the GPU box."""
import os
import sys

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
g = np.load(os.path.join(ROOT, "tests", "golden", "gftt_golden.npz"))
out = {"cv2_version": cv2.__version__}
cases = [(0, 40.0, (8, 8)), (2, 40.0, (8, 8)), (3, 40.0, (8, 8)), (2, 2.0, (8, 8)), (0, 3.0, (16, 12)), (4, 40.0, (8, 8))]
out["cases"] = np.array([(i, c, t[0], t[1]) for i, c, t in cases], np.float64)
for k, (i, clip, tiles) in enumerate(cases):
    r = cv2.createCLAHE(clip, tiles).apply(g["imgs"][i])
    out[f"rows{k}"] = r[::8].copy()
    out[f"sum{k}"] = np.int64(r.astype(np.int64).sum())
    out[f"xor{k}"] = np.bitwise_xor.reduce((r.astype(np.uint32) * (np.arange(r.size, dtype=np.uint32).reshape(r.shape) | 1)).ravel())
    print(k, i, clip, tiles, "changed pixels", int((r != g["imgs"][i]).sum()))
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "clahe_golden.npz"), **out)
