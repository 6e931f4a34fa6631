"""TEST INFRASTRUCTURE — ctypes binding of oracle/_ref/libgf2_ref.so: the REFERENCE's own factor sources (compiled unmodified from
/root/reference by oracle/Makefile against the header stand-ins of oracle/shim) behind the same call signatures as gf2_oracle.
Exists only where /root/reference does (this container); the GPU box sees the golden vectors it produced instead
(tests/golden/ref_golden.npz, tests/golden/make_ref_golden.py). Only tests/ may import this."""
import ctypes as C
import os
import numpy as np

import gf2_oracle as orc

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_ref", "libgf2_ref.so")


def available():
    if not os.path.exists(LIB_PATH) and os.path.isdir("/root/reference/Ground-Fusion++"):
        orc.build()
    return os.path.exists(LIB_PATH)


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not available():
            raise RuntimeError("oracle/_ref/libgf2_ref.so is not built (the reference tree is not present on this box)")
        _lib = C.CDLL(LIB_PATH)
    return _lib


_p = orc._p


def factor_eval(kind, consts, params, extra=None, want_jac=True):
    nres, blocks = {0: (2, [7, 7, 7, 1, 1]), 1: (15, [7, 9, 7, 9]), 2: (6, [7, 7, 7, 1, 1, 1, 1]), 3: (1, [3, 4]), 4: (1, [3, 4, 3, 4])}[kind]
    params = np.ascontiguousarray(params, dtype=np.float64)
    assert params.size == sum(blocks)
    res = np.zeros(nres); jac = np.zeros(nres * sum(blocks))
    consts = np.ascontiguousarray(consts)
    ex = np.ascontiguousarray(extra if extra is not None else [0.0], dtype=np.float64)
    lib().gf2r_factor_eval(int(kind), _p(consts), _p(ex), _p(params), _p(res), _p(jac) if want_jac else None)
    out = []; o = 0
    for s in blocks:
        out.append(jac[o:o + nres * s].reshape(nres, s).copy()); o += nres * s
    return res, out


def imu_preintegrate_one(abi, samples, n, first, lin_bias, noise):
    one = np.zeros(1, abi.IMU_PREINT)
    lib().gf2r_imu_preintegrate(_p(np.ascontiguousarray(samples)), int(n), _p(np.ascontiguousarray(first)), _p(np.ascontiguousarray(lin_bias)),
                                _p(np.ascontiguousarray(noise, dtype=np.float64)), _p(one))
    return one


def wheel_preintegrate_one(abi, samples, n, first, lin, noise):
    one = np.zeros(1, abi.WHEEL_PREINT)
    lib().gf2r_wheel_preintegrate(_p(np.ascontiguousarray(samples)), int(n), _p(np.ascontiguousarray(first)), _p(np.ascontiguousarray(lin)),
                                  _p(np.ascontiguousarray(noise, dtype=np.float64)), _p(one))
    return one


def set_noise(imu_noise=None, wheel_noise=None):
    a = None if imu_noise is None else np.ascontiguousarray(imu_noise, dtype=np.float64)
    b = None if wheel_noise is None else np.ascontiguousarray(wheel_noise, dtype=np.float64)
    lib().gf2r_set_noise(_p(a), _p(b))


def pose_plus(x, delta):
    x = np.ascontiguousarray(x, dtype=np.float64); d = np.ascontiguousarray(delta, dtype=np.float64)
    out = np.zeros(7); jac = np.zeros((7, 6))
    lib().gf2r_pose_plus(_p(x), _p(d), _p(out), _p(jac))
    return out, jac


def marginalize_window(w, i, opts, mode=0, P=96):
    """The reference's MarginalizationInfo on window i of a synth window dict (same return value as gf2_oracle.marginalize_window)."""
    from gf2_loader import load
    abi = load().abi
    keep = []
    b = orc.make_batch(w, keep)
    win = orc.Window()
    orc.lib.gf2o_batch_window(C.byref(b), int(i), C.byref(win))
    J0 = np.zeros((P, P)); r0 = np.zeros(P); nb = C.c_int32(0); m = C.c_int32(0)
    blocks = np.zeros(2 * w["n_frames"] + 8, abi.PRIOR_BLOCK)
    set_noise(w.get("imu_noise"), w.get("wheel_noise"))
    n = lib().gf2r_marginalize_window(C.byref(win), C.byref(opts), int(mode), int(P), _p(J0), _p(r0), C.byref(nb), _p(blocks), C.byref(m))
    if n < 0:
        return {"status": n, "n": 0, "m": m.value}
    return {"status": 0, "n": n, "m": m.value, "J0": J0[:n, :n].copy(), "r0": r0[:n].copy(), "blocks": blocks[:nb.value].copy()}
