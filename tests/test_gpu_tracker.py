"""GPU tests of the CUDA LK tracker through the C ABI: against the committed cv2 golden vectors, the numpy oracle and (when
importable on the box) cv2 itself. status must be bit-exact; positions within 1e-3 px (float32 rounding of the window sums:
OpenCV accumulates them in float32, the kernel exactly in int64)."""
import os

import numpy as np
import pytest

import lk_oracle as lk

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "lk_golden.npz")


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


def test_forward_and_reverse_match_golden(gf2, gold):
    n = len(gold["pts"])
    t = gf2.Tracker(640, 480, max_pts=n, max_streams=1)
    out, st, err = t.track(gold["prev"], gold["cur"], gold["pts"], max_level=3)
    assert np.array_equal(st[0], gold["fst"])
    ok = gold["fst"] == 1
    assert np.abs(out[0] - gold["fwd"])[ok].max() < 1e-3
    assert np.abs(err[0] - gold["ferr"])[ok].max() < 1e-2
    # reverse call shape: maxLevel 1 + OPTFLOW_USE_INITIAL_FLOW
    r, rst, _ = t.track(gold["cur"], gold["prev"], gold["fwd"], init_pts=gold["pts"], max_level=1, flags=gf2.abi.LK_USE_INITIAL_FLOW)
    assert np.array_equal(rst[0], gold["rst"])
    assert np.abs(r[0] - gold["rev"])[gold["rst"] == 1].max() < 1e-3
    t.close()


def test_matches_oracle_bitwise_on_status_and_tightly_on_positions(gf2):
    prev, cur, pts = lk.synthetic_pair(5, shift=(-4.2, 3.1))
    o, ost, oerr = lk.calc_optical_flow_pyr_lk(prev, cur, pts, max_level=3)
    t = gf2.Tracker(640, 480, max_pts=len(pts))
    out, st, err = t.track(prev, cur, pts)
    assert np.array_equal(st[0], ost)
    # same integer sums, same float32 operation order -> identical positions
    assert np.abs(out[0] - o)[ost == 1].max() <= 1e-5
    t.close()


def test_forward_backward_fused_and_pyramid_reuse(gf2):
    frames = [lk.synthetic_pair(7, shift=(2.0 * k, -1.0 * k))[1] for k in range(3)]
    prev0, _, pts = lk.synthetic_pair(7, shift=(0.0, 0.0))
    t = gf2.Tracker(640, 480, max_pts=len(pts))
    cp, ok = t.track_fb(prev0, frames[1], pts)
    ecp, eok = lk.track_forward_backward(prev0, frames[1], pts)
    assert np.array_equal(ok[0], eok)
    assert np.abs(cp[0] - ecp)[eok == 1].max() <= 1e-5
    # next frame: prev = None reuses the cached pyramid of frames[1] (prev_img = cur_img, feature_tracker.cpp:307)
    good = cp[0][ok[0] == 1]
    cp2, ok2 = t.track_fb(None, frames[2], good)
    ecp2, eok2 = lk.track_forward_backward(frames[1], frames[2], good)
    assert np.array_equal(ok2[0][:len(good)], eok2)
    assert np.abs(cp2[0][:len(good)] - ecp2)[eok2 == 1].max() <= 1e-5
    t.close()


def test_batched_streams_and_ragged_counts(gf2):
    pairs = [lk.synthetic_pair(10 + s, shift=(1.5 * s, 0.7 * s)) for s in range(3)]
    n = [len(p[2]) for p in pairs]; n[1] = 17; n[2] = 0
    maxp = max(n)
    t = gf2.Tracker(640, 480, max_pts=maxp, max_streams=3)
    prev = np.stack([p[0] for p in pairs]); cur = np.stack([p[1] for p in pairs])
    pts = np.zeros((3, maxp, 2), np.float32)
    for s in range(3):
        pts[s, :n[s]] = pairs[s][2][:n[s]]
    out, st, _ = t.track(prev, cur, pts, n_pts=n)
    for s in range(3):
        if n[s] == 0:
            continue
        o, ost, _ = lk.calc_optical_flow_pyr_lk(pairs[s][0], pairs[s][1], pairs[s][2][:n[s]])
        assert np.array_equal(st[s, :n[s]], ost)
        assert np.abs(out[s, :n[s]] - o)[ost == 1].max() <= 1e-5
    t.close()


def test_live_cv2_when_available(gf2):
    cv2 = pytest.importorskip("cv2")
    cv2.setNumThreads(1)
    prev, cur, pts = lk.synthetic_pair(21, shift=(5.0, 2.5))
    crit = (cv2.TERM_CRITERIA_COUNT + cv2.TERM_CRITERIA_EPS, 30, 0.01)
    c, cst, _ = cv2.calcOpticalFlowPyrLK(prev, cur, pts.reshape(-1, 1, 2), None, winSize=(21, 21), maxLevel=3, criteria=crit)
    t = gf2.Tracker(640, 480, max_pts=len(pts))
    out, st, _ = t.track(prev, cur, pts)
    assert np.array_equal(st[0], cst.ravel())
    assert np.abs(out[0] - c.reshape(-1, 2))[cst.ravel() == 1].max() < 1e-3
    t.close()


def test_tracker_rejects_bad_arguments(gf2):
    with pytest.raises(gf2.Gf2Error, match="21x21"):
        gf2.Tracker(640, 480, win=15)
    t = gf2.Tracker(640, 480, max_pts=8)
    with pytest.raises(gf2.Gf2Error, match="no pyramid"):
        t.track(None, np.zeros((480, 640), np.uint8), np.zeros((4, 2), np.float32))
    t.close()
