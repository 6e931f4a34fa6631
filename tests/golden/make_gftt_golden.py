"""Generates tests/golden/gftt_golden.npz: synthetic 640x480 images + masks and the outputs of the Python cv2 build in this image
(cv2.cornerMinEigenVal and cv2.goodFeaturesToTrack, the very function the reference calls at feature_tracker.cpp:198) for the
argument shapes of trackImage: qualityLevel 0.01, minDistance MIN_DIST = 30, maxCorners = MAX_CNT - tracked, mask from setMask().
Run once here; the .npz is committed because the GPU box must not depend on anything but the repo."""
import os
import sys

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
import lk_oracle as lk  # noqa: E402

prev, cur, pts = lk.synthetic_pair(0)
rng = np.random.default_rng(7)
imgs, masks, maxc = [], [], []
# 0: first frame of a stream: no mask restrictions, MAX_CNT = 300
imgs.append(prev); masks.append(np.full(prev.shape, 255, np.uint8)); maxc.append(300)
# 1: steady state: mask with MIN_DIST circles around 120 tracked points, 30 new corners wanted
m = np.full(cur.shape, 255, np.uint8)
for p in pts[:120]:
    cv2.circle(m, (int(round(float(p[0]))), int(round(float(p[1])))), 30, 0, -1)
imgs.append(cur); masks.append(m); maxc.append(30)
# 2: saturated / textureless regions and a strong isolated corner pattern, small budget
im2 = cur.copy(); im2[100:220, 200:400] = 255; im2[300:400, 50:150] = 0; im2[320:360, 80:120] = 200
imgs.append(im2); masks.append(np.full(cur.shape, 255, np.uint8)); maxc.append(8)
# 3: low-contrast noise image (many near-equal eigenvalues), half-plane mask
im3 = rng.integers(120, 136, size=cur.shape).astype(np.uint8)
m3 = np.zeros(cur.shape, np.uint8); m3[:, 320:] = 255
imgs.append(im3); masks.append(m3); maxc.append(150)
# 4: constant image: no corners at all
imgs.append(np.full(cur.shape, 77, np.uint8)); masks.append(np.full(cur.shape, 255, np.uint8)); maxc.append(50)
# 5: empty mask
imgs.append(cur); masks.append(np.zeros(cur.shape, np.uint8)); maxc.append(50)
out = {"imgs": np.stack(imgs), "masks": np.stack(masks), "max_corners": np.array(maxc, np.int32), "cv2_version": cv2.__version__}
for i, (im, mk, mc) in enumerate(zip(imgs, masks, maxc)):
    c = cv2.goodFeaturesToTrack(im, mc, 0.01, 30, mask=mk)
    c = np.zeros((0, 2), np.float32) if c is None else c.reshape(-1, 2)
    out[f"corners{i}"] = c
    print(i, "corners", len(c))
# the eigenvalue map of two images, stored as a bit pattern checksum + a sparse sample (the full maps are 1.2 MB each)
for i in (0, 2):
    e = cv2.cornerMinEigenVal(imgs[i], 3, 3)
    out[f"eig_rows{i}"] = e[::16].copy()             # every 16th row, bit-exact comparison
    out[f"eig_xor{i}"] = np.bitwise_xor.reduce(e.view(np.uint32).ravel())
    out[f"eig_sum{i}"] = e.astype(np.float64).sum()
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "gftt_golden.npz"), **out)
