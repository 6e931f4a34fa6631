"""CPU tests pinning the numpy CLAHE oracle bit-exactly: committed cv2 golden vectors + live cv2 when importable."""
import os

import numpy as np
import pytest

import clahe_oracle as clahe

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def check_against_golden(fn):
    g = np.load(os.path.join(GOLD, "clahe_golden.npz")); imgs = np.load(os.path.join(GOLD, "gftt_golden.npz"))["imgs"]
    for k, (i, clip, tx, ty) in enumerate(g["cases"]):
        r = fn(imgs[int(i)], float(clip), (int(tx), int(ty)))
        assert np.array_equal(r[::8], g[f"rows{k}"]), k
        assert int(r.astype(np.int64).sum()) == int(g[f"sum{k}"]), k
        assert np.bitwise_xor.reduce((r.astype(np.uint32) * (np.arange(r.size, dtype=np.uint32).reshape(r.shape) | 1)).ravel()) == g[f"xor{k}"], k


def test_clahe_bit_exact_vs_cv2_golden():
    check_against_golden(clahe.apply)


def test_clahe_live_cv2():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(2)
    for k, (shape, clip, tiles) in enumerate([((480, 640), 40.0, (8, 8)), ((120, 160), 1.5, (4, 4)), ((96, 128), 0.0, (8, 8)), ((240, 320), 4.0, (8, 6))]):
        img = cv2.GaussianBlur(rng.integers(0, 256, size=shape).astype(np.uint8), (0, 0), 2.0)
        if k % 2:
            img[: shape[0] // 3] = 17
        ref = cv2.createCLAHE(clip, tiles).apply(img)
        assert np.array_equal(clahe.apply(img, clip, tiles), ref), k
