"""CPU-only check of the detector KERNELS' arithmetic: ground-fusion2_b200/csrc/gf2_tracker_detect.cuh is compiled for the host with shims
(tests/emul/detect_emul.cpp, one CUDA thread at a time) and its min-eigenvalue map / candidate keys are compared with the cv2-pinned oracle and the
cv2 golden vectors. This is test infrastructure for boxes without a GPU (the GPU suite checks the real kernels); the product has no CPU path."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import gftt_oracle as gftt

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden", "gftt_golden.npz")


@pytest.fixture(scope="module")
def emul(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("emul") / "detect_emul.so")
    r = subprocess.run(["g++", "-O1", "-ffp-contract=off", "-std=c++17", "-shared", "-fPIC", "-o", so, os.path.join(HERE, "emul", "detect_emul.cpp")], capture_output=True, text=True)
    if r.returncode != 0:
        pytest.skip("host emulation did not compile: " + r.stderr[-300:])
    return C.CDLL(so)


@pytest.mark.parametrize("i", [1, 2, 3])
def test_detector_kernels_emulated_on_the_host_match_cv2(emul, gf2, i):
    g = np.load(GOLD)
    img = np.ascontiguousarray(g["imgs"][i]); mask = np.ascontiguousarray(g["masks"][i]); H, W = img.shape
    cap = W * H // 4
    eig = np.zeros((H, W), np.float32); keys = np.zeros(cap, np.uint64); count = np.zeros(1, np.int32); want = np.array([int(g["max_corners"][i])], np.int32)
    P = lambda a: a.ctypes.data_as(C.c_void_p)
    emul.detect_emul(P(img), P(mask), W, H, 1, P(want), C.c_double(0.01), P(eig), P(keys), cap, P(count))
    ref = gftt.corner_min_eigen_val(img)
    assert np.array_equal(eig.view(np.uint32), ref.view(np.uint32))                       # k_gftt_cov + k_gftt_eig: bit-exact score map
    idx, val = gftt.candidates(ref, 0.01, mask)
    k = keys[:count[0]]
    assert count[0] == len(idx) and np.array_equal(np.sort((k & np.uint64(0xffffffff)).astype(np.int64)), np.sort(idx))   # k_gftt_nms: same candidates
    # keys -> the library's host selection -> cv2's corner list
    out = np.zeros((max(int(want[0]), 1), 2), np.float32); n = np.zeros(1, np.int32)
    assert gf2.lib().gf2_detect_select(P(np.ascontiguousarray(k)), int(count[0]), W, H, int(want[0]), C.c_double(30.0), P(out), P(n)) == 0
    assert np.array_equal(out[:n[0]], g[f"corners{i}"])
