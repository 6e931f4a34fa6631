"""TEST INFRASTRUCTURE — numpy restatement of cv::goodFeaturesToTrack as FeatureTracker::trackImage calls it
(VE/featureTracker/feature_tracker.cpp:198: goodFeaturesToTrack(cur_img, n_pts, MAX_CNT - cur_pts.size(), 0.01, MIN_DIST, mask),
i.e. blockSize 3, gradientSize 3, useHarrisDetector false).

OpenCV is an un-vendored dependency of the reference ("OpenCV 4", GF/vins_estimator/CMakeLists.txt:84); the algorithm is
restated from its published implementation (modules/imgproc/src/featureselect.cpp, corner.cpp, deriv.cpp, box_filter) and
PINNED against the Python cv2 build in this image (tests/test_gftt_oracle.py): the min-eigenvalue map is BIT-EXACT with
cv2.cornerMinEigenVal and the corner list (order included) is identical to cv2.goodFeaturesToTrack; golden vectors in
tests/golden/gftt_golden.npz made by tests/golden/make_gftt_golden.py.

The float32 operation order that reproduces cv2 bit for bit (probed against cv2 4.13.0, documented here because the CUDA
kernels k_gftt_cov / k_gftt_eig follow the same sequence):
  Sobel (CV_8U -> CV_32F, scale = 1 / (2^(ksize-1) * blockSize * 255) folded into the SMOOTHING kernel [1 2 1]):
    Dx: row pass d = I[x+1] - I[x-1] (exact), column pass fma(d[y-1] + d[y+1], s, fl(d[y] * 2s))
    Dy: row pass r = fma(I[x+1], s, fma(I[x], 2s, fl(I[x-1] * s))), column pass fl(r[y+1] - r[y-1])
  cov = (fl(Dx*Dx), fl(Dx*Dy), fl(Dy*Dy))
  boxFilter 3x3, normalize = false, BORDER_REFLECT_101, accumulated in DOUBLE: row sums (c[x-1] + c[x]) + c[x+1]; the column
    pass is a RUNNING sum down the image (SUM = r[-1] + r[0]; per row s = SUM + r[y+1], out = float(s), SUM = s - r[y-1]) whose
    rounding history is part of the result (Dx can be ~1e-10 instead of 0, so the double sums are not exact)
  eig = fl(fl(a + c) - sqrt(fl(fl(t*t) + fl(b*b)))), a = xx/2, c = yy/2, b = xy, t = fl(a - c)   (no FMA contraction)
This is the sequence of cv2's vector code path. For image widths that are not a multiple of its unrolled SIMD step (32 floats
here: probed, W = 176 differs in columns >= 164) cv2 computes the trailing columns of the Sobel passes with non-fused code,
which differs by 1 ulp there; bit-exactness is therefore claimed (and tested) for widths that are multiples of 32 — 640 in
every shipped config.
"""
import numpy as np

f32 = np.float32
f64 = np.float64


def _fma32(a, b, c):
    """fl32(a*b + c) for float32 arrays: the product of two float32 is exact in float64; the sum is rounded to 53 bits and then to
    24, which differs from a true fused multiply-add only in double-rounding ties (never observed against cv2)."""
    return (a.astype(f64) * b.astype(f64) + c.astype(f64)).astype(f32)


def sobel_scaled(img, block_size=3):
    """(Dx, Dy) of cv::cornerEigenValsVecs: Sobel 3x3, CV_32F, scale 1/(4 * block_size * 255), BORDER_REFLECT_101."""
    scale = 1.0 / (4.0 * block_size * 255.0)
    s1 = f32(scale)
    s2 = f32(2.0 * scale)
    p = np.pad(img.astype(np.int32), 1, mode="reflect").astype(f32)   # REFLECT_101
    d = p[:, 2:] - p[:, :-2]                                           # (H+2, W) exact integers
    Dx = _fma32(d[:-2] + d[2:], np.full_like(d[2:], s1), d[1:-1] * s2)
    r = _fma32(p[:, 2:], np.full_like(d, s1), _fma32(p[:, 1:-1], np.full_like(d, s2), p[:, :-2] * s1))
    Dy = (r[2:] - r[:-2]).astype(f32)
    return Dx, Dy


def box3_running(a):
    """cv::boxFilter(a, 3x3, normalize=false, BORDER_REFLECT_101) for a CV_32F image: double accumulators, running column sum."""
    H, W = a.shape
    p = np.pad(a.astype(f64), 1, mode="reflect")
    r = (p[:, :-2] + p[:, 1:-1]) + p[:, 2:]
    out = np.empty((H, W), f32)
    SUM = r[0] + r[1]
    for i in range(H):
        s0 = SUM + r[i + 2]
        out[i] = s0.astype(f32)
        SUM = s0 - r[i]
    return out


def corner_min_eigen_val(img, block_size=3):
    """cv::cornerMinEigenVal(img, blockSize=3, ksize=3), bit-exact with cv2 (see the module docstring)."""
    assert block_size == 3, "the running-sum restatement is written for the reference's block size 3"
    Dx, Dy = sobel_scaled(img, block_size)
    a = box3_running(Dx * Dx) * f32(0.5)
    b = box3_running(Dx * Dy)
    c = box3_running(Dy * Dy) * f32(0.5)
    t = a - c
    return ((a + c) - np.sqrt(t * t + b * b)).astype(f32)


def candidates(eig, quality_level=0.01, mask=None):
    """featureselect.cpp: minMaxLoc (masked) -> threshold TOZERO at maxVal*qualityLevel -> 3x3 dilate -> local maxima strictly
    inside the image border. Returns (flat index y*W+x, value) unsorted."""
    H, W = eig.shape
    m = np.ones((H, W), bool) if mask is None else (mask != 0)
    max_val = float(eig[m].max()) if m.any() else 0.0
    thr = f32(max_val * quality_level)
    e = np.where(eig > thr, eig, f32(0))
    p = np.pad(e, 1, mode="constant", constant_values=-np.inf)
    dil = np.max([p[i:i + H, j:j + W] for i in range(3) for j in range(3)], axis=0)
    ok = (e != 0) & (e == dil) & m
    ok[0, :] = ok[-1, :] = False
    ok[:, 0] = ok[:, -1] = False
    idx = np.flatnonzero(ok)
    return idx, e.ravel()[idx]


def select_min_distance(idx, val, W, H, max_corners, min_distance):
    """Sort by (value desc, address desc) — greaterThanPtr — then the grid-accelerated greedy min-distance selection."""
    order = np.lexsort((-idx, -val.astype(f64)))
    out = []
    if min_distance >= 1:
        cell = int(round(min_distance))   # cvRound
        gw = (W + cell - 1) // cell
        gh = (H + cell - 1) // cell
        grid = [[] for _ in range(gw * gh)]
        md2 = min_distance * min_distance
        for k in order:
            y, x = divmod(int(idx[k]), W)
            xc, yc = x // cell, y // cell
            x1, y1, x2, y2 = max(0, xc - 1), max(0, yc - 1), min(gw - 1, xc + 1), min(gh - 1, yc + 1)
            good = True
            for yy in range(y1, y2 + 1):
                for xx in range(x1, x2 + 1):
                    for (px, py) in grid[yy * gw + xx]:
                        dx, dy = x - px, y - py
                        if dx * dx + dy * dy < md2:
                            good = False
                            break
                    if not good:
                        break
                if not good:
                    break
            if good:
                grid[yc * gw + xc].append((x, y))
                out.append((x, y))
                if max_corners > 0 and len(out) == max_corners:
                    break
    else:
        for k in order:
            y, x = divmod(int(idx[k]), W)
            out.append((x, y))
            if max_corners > 0 and len(out) == max_corners:
                break
    return np.array(out, f32).reshape(-1, 2)


def good_features_to_track(img, max_corners, quality_level=0.01, min_distance=30, mask=None):
    """cv::goodFeaturesToTrack(img, corners, maxCorners, qualityLevel, minDistance, mask) with the defaults the reference uses."""
    eig = corner_min_eigen_val(img)
    idx, val = candidates(eig, quality_level, mask)
    if len(idx) == 0:
        return np.zeros((0, 2), f32)
    return select_min_distance(idx, val, img.shape[1], img.shape[0], max_corners, float(min_distance))
