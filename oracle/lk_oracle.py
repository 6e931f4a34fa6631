"""TEST INFRASTRUCTURE — numpy restatement of cv::calcOpticalFlowPyrLK as FeatureTracker::trackImage calls it
(VE/featureTracker/feature_tracker.cpp:122,132,135,141: win 21x21, maxLevel 3 / 1, TermCriteria(COUNT+EPS, 30, 0.01),
flags 0 / OPTFLOW_USE_INITIAL_FLOW, minEigThreshold 1e-4).

OpenCV is an un-vendored dependency of the reference ("OpenCV 4", GF/vins_estimator/CMakeLists.txt:84); the algorithm is
restated from its published implementation (modules/video/src/lkpyramid.cpp) and PINNED against the
Python cv2 build in this image (tests/test_lk_oracle.py: identical status, positions within 1e-4 px; golden vectors in
tests/golden/lk_*.npz made by tests/golden/make_lk_golden.py).

Differences from OpenCV, by design: the window sums A11/A12/A22/b1/b2 are accumulated exactly in int64 (OpenCV accumulates
in float32, in an order that depends on its SIMD path), so results are order independent; they agree with cv2 to float32
rounding of those sums."""
import numpy as np

W_BITS = 14
FLT_SCALE = np.float32(1.0 / (1 << 20))
FLT_EPSILON = np.float32(1.1920929e-07)


def _reflect101(i, n):
    i = np.asarray(i)
    if n == 1:
        return np.zeros_like(i)
    p = 2 * (n - 1)
    i = np.mod(i, p)
    return np.where(i >= n, p - i, i)


def pyr_down(img):
    """cv::pyrDown for CV_8U: 5x5 [1 4 6 4 1] separable, BORDER_REFLECT_101, (sum + 128) >> 8, size ((w+1)/2, (h+1)/2)."""
    h, w = img.shape
    oh, ow = (h + 1) // 2, (w + 1) // 2
    k = np.array([1, 4, 6, 4, 1], np.int32)
    src = img.astype(np.int32)
    xs = _reflect101(2 * np.arange(ow)[:, None] + np.arange(-2, 3)[None, :], w)   # [ow][5]
    rows = (src[:, xs] * k).sum(-1)                                               # [h][ow]
    ys = _reflect101(2 * np.arange(oh)[:, None] + np.arange(-2, 3)[None, :], h)   # [oh][5]
    out = (rows[ys, :] * k[None, :, None]).sum(1)                                 # [oh][ow]
    return ((out + 128) >> 8).astype(np.uint8)


def build_pyramid(img, max_level):
    pyr = [np.ascontiguousarray(img)]
    for _ in range(max_level):
        pyr.append(pyr_down(pyr[-1]))
    return pyr


def scharr_deriv(img):
    """calcSharrDeriv: un-normalised 3x3 Scharr, int16, BORDER_REFLECT_101 at the image edge. Returns (dx, dy)."""
    h, w = img.shape
    s = img.astype(np.int32)
    ym, yp = _reflect101(np.arange(h) - 1, h), _reflect101(np.arange(h) + 1, h)
    xm, xp = _reflect101(np.arange(w) - 1, w), _reflect101(np.arange(w) + 1, w)
    t0 = (s[ym] + s[yp]) * 3 + s * 10     # vertical smoothing [3 10 3]
    t1 = s[yp] - s[ym]                    # vertical difference
    dx = t0[:, xp] - t0[:, xm]
    dy = (t1[:, xm] + t1[:, xp]) * 3 + t1 * 10
    return dx.astype(np.int16), dy.astype(np.int16)


def _descale(x, n):
    return (x + (1 << (n - 1))) >> n


def _patch(img_pad, pad, ip, iw, win):
    """sum of the four neighbours weighted by the 14-bit fixed-point bilinear weights (int64)."""
    y0, x0 = ip[1] + pad, ip[0] + pad
    a = img_pad[y0:y0 + win + 1, x0:x0 + win + 1].astype(np.int64)
    return a[:-1, :-1] * iw[0] + a[:-1, 1:] * iw[1] + a[1:, :-1] * iw[2] + a[1:, 1:] * iw[3]


def _weights(px, py, ipx, ipy):
    a = np.float32(px - np.float32(ipx)); b = np.float32(py - np.float32(ipy))
    one = np.float32(1.0)
    s = np.float32(1 << W_BITS)
    iw00 = int(np.rint(np.float32((one - a) * (one - b)) * s))
    iw01 = int(np.rint(np.float32(a * (one - b)) * s))
    iw10 = int(np.rint(np.float32((one - a) * b) * s))
    iw11 = (1 << W_BITS) - iw00 - iw01 - iw10
    return iw00, iw01, iw10, iw11


def calc_optical_flow_pyr_lk(prev, cur, prev_pts, next_pts=None, win=21, max_level=3, max_iters=30, eps=0.01,
                             use_initial_flow=False, min_eig=1e-4):
    """Returns (next_pts [n,2] float32, status [n] uint8, err [n] float32)."""
    prev_pts = np.asarray(prev_pts, np.float32).reshape(-1, 2)
    n = len(prev_pts)
    # OpenCV lowers maxLevel while a level is smaller than the window
    lvl = 0
    h, w = prev.shape
    while lvl < max_level and ((w + 1) // 2 > win and (h + 1) // 2 > win):
        w, h = (w + 1) // 2, (h + 1) // 2; lvl += 1
    max_level = lvl
    ppyr = build_pyramid(prev, max_level); cpyr = build_pyramid(cur, max_level)
    status = np.ones(n, np.uint8); err = np.zeros(n, np.float32)
    out = np.zeros((n, 2), np.float32)
    if use_initial_flow:
        out[:] = np.asarray(next_pts, np.float32).reshape(-1, 2)
    half = np.float32((win - 1) * 0.5)
    eps2 = float(eps) * float(eps)  # criteria.epsilon *= criteria.epsilon in double; delta.ddot(delta) is a double dot product
    pad = win + 2
    for level in range(max_level, -1, -1):
        I = ppyr[level]; J = cpyr[level]
        rows, cols = I.shape
        dx, dy = scharr_deriv(I)
        # intensity images are padded by reflection (pyramid border), derivative buffers by zeros
        yi = _reflect101(np.arange(-pad, rows + pad), rows); xi = _reflect101(np.arange(-pad, cols + pad), cols)
        Ipad = I[yi][:, xi]; Jpad = J[yi][:, xi]
        dxp = np.zeros((rows + 2 * pad, cols + 2 * pad), np.int16); dyp = np.zeros_like(dxp)
        dxp[pad:pad + rows, pad:pad + cols] = dx; dyp[pad:pad + rows, pad:pad + cols] = dy
        scale = np.float32(1.0 / (1 << level))
        for i in range(n):
            ppx = np.float32(prev_pts[i, 0] * scale); ppy = np.float32(prev_pts[i, 1] * scale)
            if level == max_level:
                if use_initial_flow:
                    nx = np.float32(out[i, 0] * scale); ny = np.float32(out[i, 1] * scale)
                else:
                    nx, ny = ppx, ppy
            else:
                nx = np.float32(out[i, 0] * np.float32(2.0)); ny = np.float32(out[i, 1] * np.float32(2.0))
            out[i] = (nx, ny)
            px = np.float32(ppx - half); py = np.float32(ppy - half)
            ipx = int(np.floor(px)); ipy = int(np.floor(py))
            if ipx < -win or ipx >= cols or ipy < -win or ipy >= rows:
                if level == 0:
                    status[i] = 0; err[i] = 0
                continue
            iw = _weights(px, py, ipx, ipy)
            Ival = _descale(_patch(Ipad, pad, (ipx, ipy), iw, win), W_BITS - 5)
            Ix = _descale(_patch(dxp, pad, (ipx, ipy), iw, win), W_BITS)
            Iy = _descale(_patch(dyp, pad, (ipx, ipy), iw, win), W_BITS)
            A11 = np.float32(np.float32(int((Ix * Ix).sum())) * FLT_SCALE)
            A12 = np.float32(np.float32(int((Ix * Iy).sum())) * FLT_SCALE)
            A22 = np.float32(np.float32(int((Iy * Iy).sum())) * FLT_SCALE)
            D = np.float32(A11 * A22 - A12 * A12)
            me = np.float32((A22 + A11 - np.sqrt(np.float32((A11 - A22) * (A11 - A22) + np.float32(4.0) * A12 * A12), dtype=np.float32)) / np.float32(2 * win * win))
            if me < np.float32(min_eig) or D < FLT_EPSILON:
                if level == 0:
                    status[i] = 0
                continue
            D = np.float32(1.0) / D
            nx = np.float32(nx - half); ny = np.float32(ny - half)
            pdx = np.float32(0); pdy = np.float32(0)
            diff = None
            for j in range(max_iters):
                inx = int(np.floor(nx)); iny = int(np.floor(ny))
                if inx < -win or inx >= cols or iny < -win or iny >= rows:
                    if level == 0:
                        status[i] = 0
                    break
                jw = _weights(nx, ny, inx, iny)
                diff = _descale(_patch(Jpad, pad, (inx, iny), jw, win), W_BITS - 5) - Ival
                b1 = np.float32(np.float32(int((diff * Ix).sum())) * FLT_SCALE)
                b2 = np.float32(np.float32(int((diff * Iy).sum())) * FLT_SCALE)
                ddx = np.float32(np.float32(A12 * b2 - A22 * b1) * D)
                ddy = np.float32(np.float32(A12 * b1 - A11 * b2) * D)
                nx = np.float32(nx + ddx); ny = np.float32(ny + ddy)
                out[i] = (np.float32(nx + half), np.float32(ny + half))
                if float(ddx) * float(ddx) + float(ddy) * float(ddy) <= eps2:
                    break
                if j > 0 and abs(float(np.float32(ddx + pdx))) < 0.01 and abs(float(np.float32(ddy + pdy))) < 0.01:
                    out[i, 0] -= np.float32(ddx * np.float32(0.5)); out[i, 1] -= np.float32(ddy * np.float32(0.5))
                    break
                pdx, pdy = ddx, ddy
            if status[i] and level == 0:
                # err = mean |J - I| / 32 at the final position (measured like OpenCV, on the window at nextPts - halfWin)
                fx = np.float32(out[i, 0] - half); fy = np.float32(out[i, 1] - half)
                inx = int(np.floor(fx)); iny = int(np.floor(fy))
                if inx < -win or inx >= cols or iny < -win or iny >= rows:
                    status[i] = 0; err[i] = 0
                    continue
                jw = _weights(fx, fy, inx, iny)
                d = _descale(_patch(Jpad, pad, (inx, iny), jw, win), W_BITS - 5) - Ival
                err[i] = np.float32(np.float32(np.abs(d).sum()) / np.float32(32 * win * win))
    return out, status, err


def track_forward_backward(prev, cur, prev_pts, win=21, max_level=3):
    """The LK part of trackImage without prediction (feature_tracker.cpp:135-153): forward at maxLevel 3, reverse at maxLevel 1
    with OPTFLOW_USE_INITIAL_FLOW, keep iff both succeed and the round trip is <= 0.5 px (distance() :22-28 in double)."""
    cur_pts, status, _ = calc_optical_flow_pyr_lk(prev, cur, prev_pts, win=win, max_level=max_level)
    rev, rstatus, _ = calc_optical_flow_pyr_lk(cur, prev, cur_pts, next_pts=np.asarray(prev_pts, np.float32).copy(), win=win, max_level=1,
                                               use_initial_flow=True)
    d = np.asarray(prev_pts, np.float64) - rev.astype(np.float64)
    dist = np.sqrt(d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1])
    ok = (status == 1) & (rstatus == 1) & (dist <= 0.5)
    return cur_pts, ok.astype(np.uint8)


def track_image_lk(prev, cur, prev_pts, predict_pts=None, flow_back=True, win=21, max_level=3):
    """The whole LK stage of trackImage (feature_tracker.cpp:113-153). With a prediction (:118-131): forward at maxLevel 1 from
    predict_pts with OPTFLOW_USE_INITIAL_FLOW; fewer than 10 successes -> forward again at maxLevel 3 from prev_pts (flags 0).
    Returns (cur_pts, status, used_fallback)."""
    prev_pts = np.asarray(prev_pts, np.float32)
    fallback = False
    if predict_pts is not None:
        cur_pts, status, _ = calc_optical_flow_pyr_lk(prev, cur, prev_pts, next_pts=np.asarray(predict_pts, np.float32).copy(), win=win, max_level=1,
                                                      use_initial_flow=True)
        if int(status.sum()) < 10:
            fallback = True
            cur_pts, status, _ = calc_optical_flow_pyr_lk(prev, cur, prev_pts, win=win, max_level=max_level)
    else:
        cur_pts, status, _ = calc_optical_flow_pyr_lk(prev, cur, prev_pts, win=win, max_level=max_level)
    ok = status == 1
    if flow_back:
        rev, rstatus, _ = calc_optical_flow_pyr_lk(cur, prev, cur_pts, next_pts=prev_pts.copy(), win=win, max_level=1, use_initial_flow=True)
        d = prev_pts.astype(np.float64) - rev.astype(np.float64)
        ok = ok & (rstatus == 1) & (np.sqrt(d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) <= 0.5)
    return cur_pts, ok.astype(np.uint8), fallback


def synthetic_pair(*a, **kw):
    """Synthetic image pair + corners: the generator lives with the other synthetic inputs (gf2_b200.synth.image_pair)."""
    import importlib
    from gf2_loader import load
    load()
    return importlib.import_module("gf2_b200.synth").image_pair(*a, **kw)


