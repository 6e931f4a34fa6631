"""GPU tests of the CUDA goodFeaturesToTrack (gf2_tracker_detect / gf2_tracker_min_eigen_map) and of the prediction path of the LK
stage (gf2_tracker_track_image) through the C ABI: against the committed cv2 golden vectors, the numpy oracles and (when
importable on the box) cv2 itself. The score map must be bit-exact, corner lists identical in content AND order (feature ids are
assigned in that order, feature_tracker.cpp:85-93)."""
import os

import numpy as np
import pytest

import gftt_oracle as gftt
import lk_oracle as lk

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "gftt_golden.npz")


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


def test_min_eigen_map_bit_exact(gf2, gold):
    t = gf2.Tracker(640, 480, max_pts=8, max_streams=2)
    e = t.min_eigen_map(np.stack([gold["imgs"][0], gold["imgs"][2]]))
    for k, i in enumerate((0, 2)):
        assert np.array_equal(e[k, ::16].view(np.uint32), gold[f"eig_rows{i}"].view(np.uint32))
        assert np.bitwise_xor.reduce(e[k].view(np.uint32).ravel()) == gold[f"eig_xor{i}"]
        assert e[k].astype(np.float64).sum() == gold[f"eig_sum{i}"]
        assert np.array_equal(e[k].view(np.uint32), gftt.corner_min_eigen_val(gold["imgs"][i]).view(np.uint32))
    t.close()


def test_corner_lists_identical_to_cv2_golden_batched(gf2, gold):
    S = 6
    t = gf2.Tracker(640, 480, max_pts=300, max_streams=S)
    out = t.detect(gold["imgs"], gold["max_corners"], mask=gold["masks"], quality_level=0.01, min_distance=30.0)
    for i in range(S):
        assert np.array_equal(out[i], gold[f"corners{i}"]), i
    # one stream at a time, no mask argument where the golden mask is all-ones
    one = t.detect(gold["imgs"][0], int(gold["max_corners"][0]))
    assert np.array_equal(one[0], gold["corners0"])
    assert t.last_timing()["candidates"] > 1000
    t.close()


def test_detect_on_cached_image_and_skipped_streams(gf2, gold):
    prev, cur, pts = lk.synthetic_pair(0)
    t = gf2.Tracker(640, 480, max_pts=300, max_streams=2)
    t.track_fb(np.stack([prev, prev]), np.stack([cur, gold["imgs"][2]]), np.stack([pts, pts]))
    # img = None: the detector runs on the `cur` images the tracker holds; max_corners 0 skips a stream (n_pts.clear())
    out = t.detect(None, [30, 0], mask=np.stack([gold["masks"][1], gold["masks"][2]]), n_streams=2)
    assert np.array_equal(out[0], gold["corners1"]) and len(out[1]) == 0
    out = t.detect(None, [0, 8], n_streams=2)
    assert len(out[0]) == 0 and np.array_equal(out[1], gold["corners2"])
    with pytest.raises(gf2.Gf2Error, match="negative"):
        t.detect(None, [-1, 8], n_streams=2)
    t.close()
    t2 = gf2.Tracker(640, 480, max_pts=8)
    with pytest.raises(gf2.Gf2Error, match="no image is cached"):
        t2.detect(None, 5)
    t2.close()


def test_live_cv2_when_available(gf2):
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(3)
    W, H = 320, 200
    t = gf2.Tracker(W, H, max_pts=400, max_streams=1, max_level=1)
    for k in range(3):
        img = cv2.GaussianBlur(rng.integers(0, 256, size=(H, W)).astype(np.uint8), (0, 0), 1.2 + 0.4 * k)
        img = cv2.normalize(img, None, 0, 255, cv2.NORM_MINMAX).astype(np.uint8)
        assert np.array_equal(t.min_eigen_map(img)[0].view(np.uint32), cv2.cornerMinEigenVal(img, 3, 3).view(np.uint32))
        mask = (rng.random(img.shape) > 0.3).astype(np.uint8) * 255
        for mc, md in ((25, 12.0), (400, 3.0), (60, 7.5)):
            ref = cv2.goodFeaturesToTrack(img, mc, 0.01, md, mask=mask)
            ref = np.zeros((0, 2), np.float32) if ref is None else ref.reshape(-1, 2)
            assert np.array_equal(t.detect(img, mc, mask=mask, min_distance=md)[0], ref)
    t.close()


def test_prediction_path_and_per_stream_fallback(gf2):
    """hasPrediction (feature_tracker.cpp:118-131): stream 0 gets a good prediction (level-1 result kept), stream 1 a prediction far
    outside the image (< 10 successes -> redone at level 3 from prev_pts): decided per stream on the device."""
    prev, cur, pts = lk.synthetic_pair(9, shift=(6.0, -3.5))
    preds = [pts + np.float32([5.5, -3.0]), pts + np.float32([2000.0, 0.0])]
    t = gf2.Tracker(640, 480, max_pts=len(pts), max_streams=2)
    for flow_back in (False, True):
        out, st = t.track_image(np.stack([prev, prev]), np.stack([cur, cur]), np.stack([pts, pts]), predict_pts=np.stack(preds), flow_back=flow_back)
        for s in range(2):
            c, ok, fb = lk.track_image_lk(prev, cur, pts, predict_pts=preds[s], flow_back=flow_back)
            assert fb == (s == 1)
            assert np.array_equal(st[s], ok)
            assert np.abs(out[s] - c)[ok == 1].max() <= 1e-5
    # without a prediction the call equals track_fb
    a, sa = t.track_image(prev, cur, pts)
    b, sb = t.track_fb(prev, cur, pts)
    assert np.array_equal(a, b) and np.array_equal(sa, sb)
    t.close()


def test_clahe_bit_exact_and_fused_into_the_front_end(gf2, gold):
    """cv::createCLAHE()->apply of the node (VE/rosNodeTest.cpp:271-276, equalize: 1 in m3dgr.yaml) on the device: bit-exact with the cv2
    golden vectors; switched on inside the tracker, track + detect see the equalised images without a host round trip."""
    import clahe_oracle
    from test_clahe_oracle import check_against_golden
    t = gf2.Tracker(640, 480, max_pts=300, max_streams=2)
    check_against_golden(lambda img, clip, tiles: t.equalize(img, clip, tiles)[0])
    both = t.equalize(np.stack([gold["imgs"][0], gold["imgs"][2]]))
    assert np.array_equal(both[0], clahe_oracle.apply(gold["imgs"][0])) and np.array_equal(both[1], clahe_oracle.apply(gold["imgs"][2]))
    with pytest.raises(gf2.Gf2Error, match="does not divide"):
        t.equalize(gold["imgs"][0], 40.0, (7, 8))
    # fused: the same result as equalising on the host first
    prev, cur, pts = lk.synthetic_pair(3, shift=(3.0, 1.0))
    dark = lambda im: (im // 3 + 20).astype(np.uint8)   # low contrast: CLAHE changes it a lot
    pe, ce = clahe_oracle.apply(dark(prev)), clahe_oracle.apply(dark(cur))
    ref_out, ref_ok = t.track_fb(pe, ce, pts)
    ref_new = t.detect(ce, 40)
    t.set_equalize(40.0, (8, 8))
    out, ok = t.track_fb(dark(prev), dark(cur), pts)
    assert np.array_equal(out, ref_out) and np.array_equal(ok, ref_ok)
    assert np.array_equal(t.get_image(1)[0], ce)
    assert np.array_equal(t.detect(None, 40)[0], ref_new[0])          # on the cached (equalised) image
    assert np.array_equal(t.detect(dark(cur), 40)[0], ref_new[0])     # on an uploaded image: equalised on the way in
    t.set_equalize(0.0)
    assert not np.array_equal(t.detect(dark(cur), 40)[0], ref_new[0])
    t.close()
