"""Generates tests/golden/lk_golden.npz: inputs (synthetic 640x480 pair + points) and the outputs of the Python cv2 build in
this image (cv2.calcOpticalFlowPyrLK, the very function the reference calls) for the two call shapes of trackImage:
forward (maxLevel 3) and reverse (maxLevel 1, OPTFLOW_USE_INITIAL_FLOW). Run once here; the .npz is committed because the GPU
box must not depend on anything but the repo."""
import os
import sys

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
import lk_oracle as lk  # noqa: E402

cv2.setNumThreads(1)
prev, cur, pts = lk.synthetic_pair(0)
extra = np.array([[1.5, 1.5], [638.2, 2.0], [3.0, 477.0], [637.0, 478.5], [10.2, 200.0], [320.0, 9.9], [0.3, 240.0], [639.4, 100.0],
                  [-5.0, 50.0], [700.0, 300.0]], np.float32)  # border, out-of-image
flat = prev.copy(); flat[200:260, 300:380] = 128  # a textureless patch -> minEig failures
pts2 = np.concatenate([pts, extra, np.array([[340.0, 230.0], [325.5, 215.25]], np.float32)])
crit = (cv2.TERM_CRITERIA_COUNT + cv2.TERM_CRITERIA_EPS, 30, 0.01)
fwd, fst, ferr = cv2.calcOpticalFlowPyrLK(flat, cur, pts2.reshape(-1, 1, 2), None, winSize=(21, 21), maxLevel=3, criteria=crit)
rev, rst, _ = cv2.calcOpticalFlowPyrLK(cur, flat, fwd.copy(), pts2.reshape(-1, 1, 2).copy(), winSize=(21, 21), maxLevel=1, criteria=crit,
                                       flags=cv2.OPTFLOW_USE_INITIAL_FLOW)
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "lk_golden.npz"), prev=flat, cur=cur, pts=pts2, fwd=fwd.reshape(-1, 2), fst=fst.ravel(),
                    ferr=ferr.ravel(), rev=rev.reshape(-1, 2), rst=rst.ravel(), cv2_version=cv2.__version__)
print("points", len(pts2), "forward ok", int(fst.sum()), "reverse ok", int(rst.sum()), "cv2", cv2.__version__)
