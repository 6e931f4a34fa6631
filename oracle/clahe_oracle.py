"""TEST INFRASTRUCTURE — numpy restatement of cv::CLAHE::apply for 8-bit images as the reference's node calls it when EQUALIZE is set
(VE/rosNodeTest.cpp:271-276: cv::createCLAHE() -> clipLimit 40.0, tileGridSize 8x8; `equalize: 1` in config/realsense/m3dgr.yaml:16).

OpenCV is an un-vendored dependency ("OpenCV 4", GF/vins_estimator/CMakeLists.txt:84); the algorithm is restated from its published
implementation (modules/imgproc/src/clahe.cpp: CLAHE_CalcLut_Body, CLAHE_Interpolation_Body) and PINNED bit-exactly against the
Python cv2 build in this image (tests/test_clahe_oracle.py; golden vectors tests/golden/clahe_golden.npz made by
tests/golden/make_clahe_golden.py). Tile grids that do not divide the image (cv pads by reflection) are not restated."""
import numpy as np

f32 = np.float32


def tile_luts(img, clip_limit=40.0, tiles=(8, 8)):
    """Per-tile lookup tables [tiles_y * tiles_x][256] (CLAHE_CalcLut_Body)."""
    H, W = img.shape
    tx, ty = tiles
    assert W % tx == 0 and H % ty == 0, "tile grid must divide the image"
    tw, th = W // tx, H // ty
    area = tw * th
    cl = max(int(clip_limit * area / 256), 1) if clip_limit > 0 else 0
    lut = np.zeros((ty * tx, 256), np.uint8)
    scale = f32(255) / f32(area)
    for j in range(ty):
        for i in range(tx):
            hist = np.bincount(img[j * th:(j + 1) * th, i * tw:(i + 1) * tw].ravel(), minlength=256).astype(np.int64)
            if cl > 0:
                clipped = int(np.maximum(hist - cl, 0).sum())
                hist = np.minimum(hist, cl)
                batch = clipped // 256
                resid = clipped - batch * 256
                hist += batch
                if resid != 0:
                    step = max(256 // resid, 1)
                    k = 0
                    while k < 256 and resid > 0:
                        hist[k] += 1
                        k += step
                        resid -= 1
            lut[j * tx + i] = np.clip(np.rint(np.cumsum(hist).astype(f32) * scale), 0, 255).astype(np.uint8)
    return lut


def apply(img, clip_limit=40.0, tiles=(8, 8)):
    """cv2.createCLAHE(clip_limit, tiles).apply(img): bilinear blend of the neighbouring tile LUTs in float32 (CLAHE_Interpolation_Body)."""
    H, W = img.shape
    tx, ty = tiles
    lut = tile_luts(img, clip_limit, tiles)
    tw, th = W // tx, H // ty
    inv_tw = f32(1.0) / f32(tw)
    inv_th = f32(1.0) / f32(th)
    txf = (np.arange(W).astype(f32) * inv_tw - f32(0.5)).astype(f32)
    tx1 = np.floor(txf).astype(np.int64)
    xa = (txf - tx1.astype(f32)).astype(f32)
    xa1 = (f32(1.0) - xa).astype(f32)
    tx2 = np.minimum(tx1 + 1, tx - 1)
    tx1 = np.maximum(tx1, 0)
    tyf = (np.arange(H).astype(f32) * inv_th - f32(0.5)).astype(f32)
    ty1 = np.floor(tyf).astype(np.int64)
    ya = (tyf - ty1.astype(f32)).astype(f32)
    ya1 = (f32(1.0) - ya).astype(f32)
    ty2 = np.minimum(ty1 + 1, ty - 1)
    ty1 = np.maximum(ty1, 0)
    v = img.astype(np.int64)
    l11 = lut[ty1[:, None] * tx + tx1[None, :], v].astype(f32)
    l12 = lut[ty1[:, None] * tx + tx2[None, :], v].astype(f32)
    l21 = lut[ty2[:, None] * tx + tx1[None, :], v].astype(f32)
    l22 = lut[ty2[:, None] * tx + tx2[None, :], v].astype(f32)
    top = ((l11 * xa1[None]).astype(f32) + (l12 * xa[None]).astype(f32)).astype(f32)
    bot = ((l21 * xa1[None]).astype(f32) + (l22 * xa[None]).astype(f32)).astype(f32)
    res = ((top * ya1[:, None]).astype(f32) + (bot * ya[:, None]).astype(f32)).astype(f32)
    return np.clip(np.rint(res), 0, 255).astype(np.uint8)
