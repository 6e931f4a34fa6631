"""CPU tests pinning the numpy goodFeaturesToTrack oracle: against the committed cv2 golden vectors (tests/golden/gftt_golden.npz,
made by make_gftt_golden.py with the cv2 build the reference's OpenCV call maps to) and, when cv2 is importable, against cv2 live."""
import os

import numpy as np
import pytest

import gftt_oracle as gftt

GOLD = os.path.join(os.path.dirname(__file__), "golden", "gftt_golden.npz")


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


@pytest.mark.parametrize("i", [0, 2])
def test_min_eigen_map_bit_exact_vs_cv2_golden(gold, i):
    e = gftt.corner_min_eigen_val(gold["imgs"][i])
    assert np.array_equal(e[::16].view(np.uint32), gold[f"eig_rows{i}"].view(np.uint32))
    assert np.bitwise_xor.reduce(e.view(np.uint32).ravel()) == gold[f"eig_xor{i}"]
    assert e.astype(np.float64).sum() == gold[f"eig_sum{i}"]


@pytest.mark.parametrize("i", range(6))
def test_corner_list_identical_to_cv2_golden(gold, i):
    c = gftt.good_features_to_track(gold["imgs"][i], int(gold["max_corners"][i]), 0.01, 30, gold["masks"][i])
    ref = gold[f"corners{i}"]
    assert c.shape == ref.shape
    assert np.array_equal(c, ref)          # same corners in the same order: feature ids follow this order (addPoints, :85-93)


def test_live_cv2_random_images():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(11)
    for k in range(3):
        base = rng.integers(0, 256, size=(120 + 17 * k, 160 + 32 * k)).astype(np.uint8)   # widths: multiples of 32, see gftt_oracle docstring
        img = cv2.GaussianBlur(base, (0, 0), 1.5 + 0.5 * k)
        img = cv2.normalize(img, None, 0, 255, cv2.NORM_MINMAX).astype(np.uint8)
        assert np.array_equal(gftt.corner_min_eigen_val(img).view(np.uint32), cv2.cornerMinEigenVal(img, 3, 3).view(np.uint32))
        mask = (rng.random(img.shape) > 0.2).astype(np.uint8) * 255
        for mc, md in ((20, 10), (0, 7), (500, 1), (15, 0)):
            ref = cv2.goodFeaturesToTrack(img, mc, 0.01, md, mask=mask)
            ref = np.zeros((0, 2), np.float32) if ref is None else ref.reshape(-1, 2)
            assert np.array_equal(gftt.good_features_to_track(img, mc, 0.01, md, mask), ref)


@pytest.mark.parametrize("i", [0, 1, 2, 3])
def test_library_corner_selection_matches_cv2_golden(gf2, gold, i):
    """The host half of gf2_tracker_detect (sort order + min-distance grid, integer-exact) runs without a device: feed it the
    oracle's candidates and compare with cv2's corner list."""
    img, mask = gold["imgs"][i], gold["masks"][i]
    idx, val = gftt.candidates(gftt.corner_min_eigen_val(img), 0.01, mask)
    got = gf2.detect_select(idx, val, img.shape[1], img.shape[0], int(gold["max_corners"][i]), 30.0)
    assert np.array_equal(got, gold[f"corners{i}"])


def test_library_corner_selection_ties_and_fractional_distance(gf2):
    rng = np.random.default_rng(5)
    W, H = 96, 64
    idx = rng.choice(W * H, 800, replace=False)
    val = rng.integers(1, 6, 800).astype(np.float32) * np.float32(0.125)      # many exact ties -> the address tie-break decides
    for mc, md in ((50, 7.5), (1000, 0.0), (10, 2.5), (1000, 30.0)):
        ref = gftt.select_min_distance(idx, val, W, H, mc, md)
        assert np.array_equal(gf2.detect_select(idx, val, W, H, mc, md), ref)
