"""CPU tests pinning the numpy LK oracle: against the committed cv2 golden vectors (tests/golden/lk_golden.npz, made by
make_lk_golden.py with the cv2 build the reference's OpenCV call maps to) and, when cv2 is importable, against cv2 live."""
import os

import numpy as np
import pytest

import lk_oracle as lk

GOLD = os.path.join(os.path.dirname(__file__), "golden", "lk_golden.npz")


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


def test_forward_matches_cv2_golden(gold):
    o, st, err = lk.calc_optical_flow_pyr_lk(gold["prev"], gold["cur"], gold["pts"], max_level=3)
    assert np.array_equal(st, gold["fst"])                      # status: exact
    ok = gold["fst"] == 1
    assert 0 < (~ok).sum() < 20                                 # the fixture holds real failures (border, flat patch, outside)
    assert np.abs(o - gold["fwd"])[ok].max() < 1e-3             # positions: float32 rounding of the window sums
    assert np.abs(err - gold["ferr"])[ok].max() < 1e-2


def test_reverse_with_initial_flow_matches_cv2_golden(gold):
    r, rst, _ = lk.calc_optical_flow_pyr_lk(gold["cur"], gold["prev"], gold["fwd"], next_pts=gold["pts"].copy(), max_level=1, use_initial_flow=True)
    assert np.array_equal(rst, gold["rst"])
    ok = gold["rst"] == 1
    assert np.abs(r - gold["rev"])[ok].max() < 1e-3


def test_pyramid_and_derivatives_match_cv2():
    cv2 = pytest.importorskip("cv2")
    prev, _, _ = lk.synthetic_pair(3, w=321, h=243)  # odd sizes exercise (w+1)/2 and the reflect borders
    assert np.array_equal(lk.pyr_down(prev), cv2.pyrDown(prev))
    dx, dy = lk.scharr_deriv(prev)
    assert np.array_equal(dx, cv2.Scharr(prev, cv2.CV_16S, 1, 0, borderType=cv2.BORDER_REFLECT_101))
    assert np.array_equal(dy, cv2.Scharr(prev, cv2.CV_16S, 0, 1, borderType=cv2.BORDER_REFLECT_101))


@pytest.mark.parametrize("seed,shift", [(1, (3.3, -2.1)), (2, (-6.0, 4.5))])
def test_live_cv2(seed, shift):
    cv2 = pytest.importorskip("cv2")
    cv2.setNumThreads(1)
    prev, cur, pts = lk.synthetic_pair(seed, shift=shift)
    crit = (cv2.TERM_CRITERIA_COUNT + cv2.TERM_CRITERIA_EPS, 30, 0.01)
    c, cst, _ = cv2.calcOpticalFlowPyrLK(prev, cur, pts.reshape(-1, 1, 2), None, winSize=(21, 21), maxLevel=3, criteria=crit)
    o, st, _ = lk.calc_optical_flow_pyr_lk(prev, cur, pts, max_level=3)
    assert np.array_equal(st, cst.ravel())
    assert np.abs(o - c.reshape(-1, 2))[st == 1].max() < 1e-3
    # forward/backward glue of trackImage (feature_tracker.cpp:137-153)
    cp, ok = lk.track_forward_backward(prev, cur, pts)
    assert ok.mean() > 0.9
    flow = (cp - pts)[ok == 1].mean(0)
    assert np.abs(flow - np.array(shift)).max() < 0.3


def test_prediction_path_against_live_cv2():
    """trackImage with hasPrediction (feature_tracker.cpp:118-131): level-1 LK from predicted positions, fall-back to level 3."""
    cv2 = pytest.importorskip("cv2")
    cv2.setNumThreads(1)
    prev, cur, pts = lk.synthetic_pair(9, shift=(6.0, -3.5))
    crit = (cv2.TERM_CRITERIA_COUNT + cv2.TERM_CRITERIA_EPS, 30, 0.01)
    for pred, expect_fallback in ((pts + np.float32([5.5, -3.0]), False), (pts + np.float32([2000.0, 0.0]), True)):
        c, st, fb = lk.track_image_lk(prev, cur, pts, predict_pts=pred, flow_back=False)
        r, rst, _ = cv2.calcOpticalFlowPyrLK(prev, cur, pts.reshape(-1, 1, 2), pred.reshape(-1, 1, 2).copy(), winSize=(21, 21), maxLevel=1, criteria=crit,
                                             flags=cv2.OPTFLOW_USE_INITIAL_FLOW)
        if int(rst.sum()) < 10:
            r, rst, _ = cv2.calcOpticalFlowPyrLK(prev, cur, pts.reshape(-1, 1, 2), None, winSize=(21, 21), maxLevel=3)
        assert fb == expect_fallback
        assert np.array_equal(st, rst.ravel())
        assert np.abs(c - r.reshape(-1, 2))[st == 1].max() < 1e-3
