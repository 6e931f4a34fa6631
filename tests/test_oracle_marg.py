"""Pins oracle/gf2o_marg.h (restated MarginalizationInfo, VE/factor/marginalization_factor.cpp:12-330) against an
independent numpy restatement built from single-factor evaluations (which tests/test_oracle_factors.py pins by finite
differences) and numpy.linalg.eigh."""
import importlib

import numpy as np
import pytest


@pytest.fixture(scope="module")
def synth(gf2):
    return importlib.import_module("gf2_b200.synth")

EPS = 1e-8


def _huber_correct(res, jacs, delta=1.0):
    sq = float(res @ res)
    if sq > delta * delta:
        r = np.sqrt(sq); rho1 = delta / r; rho2 = -rho1 / (2 * sq)
    else:
        rho1, rho2 = 1.0, 0.0
    s1 = np.sqrt(rho1)
    if sq == 0.0 or rho2 <= 0.0:
        scal, a = s1, 0.0
    else:
        D = 1.0 + 2.0 * sq * rho2 / rho1
        alpha = 1.0 - np.sqrt(D)
        scal = s1 / (1 - alpha); a = alpha / sq
    jacs = [s1 * (J - a * np.outer(res, res @ J)) for J in jacs]
    return res * scal, jacs


def assemble_margin_old(w, i, oracle, opts):
    """A = sum J^T J, b = sum J^T r of the MARGIN_OLD factors (estimator.cpp:3396-3528) in a canonical tangent layout:
    returns (A, b, mm, rr, T): dropped columns mm, kept (touched) columns rr, T = size of the non-landmark layout."""
    F = w["n_frames"]
    T = 15 * F + 17
    base = {2: 15 * F, 3: 15 * F + 6, 4: 15 * F + 7, 5: 15 * F + 13, 6: 15 * F + 14, 7: 15 * F + 15, 8: 15 * F + 16}
    lm0 = [l for l in range(int(w["n_landmarks"][i])) if w["start_frame"][i, l] == 0]
    NT = T + len(lm0)
    A = np.zeros((NT, NT)); b = np.zeros(NT)
    touched = np.zeros(NT, bool)

    def add(res, jacs, cols):
        J = np.zeros((len(res), NT))
        for Jb, c in zip(jacs, cols):
            J[:, c] += Jb[:, :len(c)]
            touched[c] = True
        A[...] += J.T @ J; b[...] += J.T @ res

    pose = lambda f: np.arange(15 * f, 15 * f + 6)
    sb = lambda f: np.arange(15 * f + 6, 15 * f + 15)
    # prior
    n0 = int(w["prior_rows"][i])
    if n0 > 0:
        J0 = w["prior_J0"][i, :n0, :n0]; r0 = w["prior_r0"][i, :n0]
        dx = np.zeros(n0); cols = np.zeros(n0, int)
        for blk in w["prior_blocks"][i, :int(w["prior_nblocks"][i])]:
            kind, idx, off = int(blk["kind"]), int(blk["index"]), int(blk["offset"])
            if kind == 0:
                x = w["para_pose"][i, idx]; x0 = blk["x0"][:7]
                q0 = x0[3:7]; q = x[3:7]
                # q0^-1 * q, xyzw
                w0, v0 = q0[3], -q0[:3]; w1, v1 = q[3], q[:3]
                dq_w = w0 * w1 - v0 @ v1; dq_v = w0 * v1 + w1 * v0 + np.cross(v0, v1)
                d = np.concatenate([x[:3] - x0[:3], 2 * dq_v * (1 if dq_w >= 0 else -1)])
                cols[off:off + 6] = pose(idx); dx[off:off + 6] = d
            elif kind == 1:
                cols[off:off + 9] = sb(idx); dx[off:off + 9] = w["para_speedbias"][i, idx] - blk["x0"][:9]
            else:
                raise AssertionError
        res = r0 + J0 @ dx
        J = np.zeros((n0, NT)); J[:, cols] = J0
        touched[cols] = True
        A += J.T @ J; b += J.T @ res
    # IMU 0
    rec = w["imu"][i, 0]
    params = np.concatenate([w["para_pose"][i, 0], w["para_speedbias"][i, 0], w["para_pose"][i, 1], w["para_speedbias"][i, 1]])
    res, jacs = oracle.factor_eval(1, np.frombuffer(rec.tobytes(), np.uint8), params, extra=[opts.g_norm])
    add(res, jacs, [pose(0), sb(0), pose(1), sb(1)])
    # wheel factor 0 (estimator.cpp:3428-3439)
    if w.get("use_wheel") and "wheel" in w:
        recw = w["wheel"][i, 0]
        params = np.concatenate([w["para_pose"][i, 0], w["para_pose"][i, 1], w["ex_pose_wheel"][i], w["sxsysw"][i], [w["td_wheel"][i]]])
        res, jacs = oracle.factor_eval(2, np.frombuffer(recw.tobytes(), np.uint8), params)
        add(res, jacs, [pose(0), pose(1), np.arange(base[4], base[4] + 6), np.array([base[5]]), np.array([base[6]]), np.array([base[7]]), np.array([base[8]])])
    # projection factors of landmarks hosted in frame 0
    obeg = np.concatenate([[0], np.cumsum(w["track_len"][i])])
    for k, l in enumerate(lm0):
        oi = w["obs"][i, obeg[l]]
        for t in range(1, int(w["track_len"][i, l])):
            oj = w["obs"][i, obeg[l] + t]
            consts = np.array([oi["x"], oi["y"], oi["vx"], oi["vy"], w["frame_td"][i, 0], oj["x"], oj["y"], oj["vx"], oj["vy"], w["frame_td"][i, t], opts.sqrt_info_px], dtype=np.float64)
            params = np.concatenate([w["para_pose"][i, 0], w["para_pose"][i, t], w["ex_pose"][i], [w["inv_depth"][i, l]], [w["td"][i]]])
            res, jacs = oracle.factor_eval(0, consts, params)
            res, jacs = _huber_correct(res, jacs, opts.huber_delta)
            add(res, jacs, [pose(0), pose(t), np.arange(base[2], base[2] + 6), np.array([T + k]), np.array([base[3]])])
    mm = np.concatenate([pose(0), sb(0), T + np.arange(len(lm0))])
    mm = np.array([c for c in mm if touched[c]], dtype=int)
    rr = np.array([c for c in range(T) if touched[c] and c >= 15])
    return A, b, mm, rr, T


def numpy_marginalize_old(w, i, oracle, abi, opts):
    """MARGIN_OLD (estimator.cpp:3396-3595): returns (H, g) over the layout of gf2_oracle.prior_information of the window
    AFTER the slide, plus m."""
    F = w["n_frames"]
    A, b, mm, rr, T = assemble_margin_old(w, i, oracle, opts)
    Amm = 0.5 * (A[np.ix_(mm, mm)] + A[np.ix_(mm, mm)].T)
    ev, V = np.linalg.eigh(Amm)
    inv = np.where(ev > EPS, 1.0 / np.where(ev > EPS, ev, 1.0), 0.0)
    Ainv = (V * inv) @ V.T
    Ar = A[np.ix_(rr, rr)] - A[np.ix_(rr, mm)] @ Ainv @ A[np.ix_(mm, rr)]
    br = b[rr] - A[np.ix_(rr, mm)] @ Ainv @ b[mm]
    S, V2 = np.linalg.eigh(Ar)
    keep = S > EPS
    Hk = (V2[:, keep] * S[keep]) @ V2[:, keep].T
    gk = V2[:, keep] @ (V2[:, keep].T @ br)
    # shift to the indexing after slideWindow: frame f -> f - 1
    H = np.zeros((T, T)); g = np.zeros(T)
    new = np.where(rr < 15 * F, rr - 15, rr)
    H[np.ix_(new, new)] = Hk; g[new] = gk
    return H, g, len(mm)


@pytest.mark.parametrize("prior", ["anchor", "dense"])
def test_margin_old_matches_numpy(gf2, oracle, synth, prior):
    abi = gf2.abi
    w = synth.make_windows(2, n_landmarks=160, prior=prior)
    oracle.imu_preintegrate(w)
    opts = abi.default_opts()
    oracle.solve_batch(w, opts)     # marginalization runs at the solved states
    for i in range(2):
        got = oracle.marginalize_window(w, i, opts, mode=0)
        assert got["status"] == 0
        H, g, m = numpy_marginalize_old(w, i, oracle, abi, opts)
        assert got["m"] == m == 15 + int((w["start_frame"][i, :int(w["n_landmarks"][i])] == 0).sum())
        Ho, go, x0 = oracle.prior_information(got, w["n_frames"])
        scale = np.abs(H).max()
        assert np.abs(Ho - H).max() <= 1e-7 * scale   # pinv(Amm) amplifies the Jacobi-vs-LAPACK eigenvector rounding by cond(Amm)
        assert np.abs(go - g).max() <= 1e-7 * max(1.0, np.abs(g).max())
        # kept blocks: poses 1..10 -> 0..9, speed-bias 1 -> 0, ex-pose, td; x0 = the states at marginalization time
        kinds = sorted((int(b["kind"]), int(b["index"])) for b in got["blocks"])
        assert kinds == sorted([(0, f) for f in range(w["n_frames"] - 1)] + [(1, 0), (2, 0), (3, 0)])
        assert got["n"] == 6 * (w["n_frames"] - 1) + 9 + 6 + 1
        assert np.array_equal(x0[(0, 3)][:7], w["para_pose"][i, 4]) and np.array_equal(x0[(1, 0)], w["para_speedbias"][i, 1])
        assert np.array_equal(x0[(2, 0)][:7], w["ex_pose"][i])


def test_margin_old_with_wheel_matches_numpy(gf2, oracle, synth):
    abi = gf2.abi
    w = synth.make_windows(2, config_id=4, n_landmarks=160, wheel=True, prior="dense")
    oracle.imu_preintegrate(w); oracle.wheel_preintegrate(w)
    w["sxsysw"][1] = [1.01, 0.99, 1.02]; w["td_wheel"][1] = 0.004    # off the linearisation point of the wheel preintegration
    opts = abi.default_opts()
    for i in range(2):
        got = oracle.marginalize_window(w, i, opts, mode=0)
        assert got["status"] == 0
        H, g, m = numpy_marginalize_old(w, i, oracle, abi, opts)
        Ho, go, x0 = oracle.prior_information(got, w["n_frames"])
        assert np.abs(Ho - H).max() <= 1e-7 * np.abs(H).max()
        assert np.abs(go - g).max() <= 1e-7 * max(1.0, np.abs(g).max())
        kinds = sorted((int(b["kind"]), int(b["index"])) for b in got["blocks"])
        assert kinds == sorted([(0, f) for f in range(w["n_frames"] - 1)] + [(1, 0), (2, 0), (3, 0), (4, 0), (5, 0), (6, 0), (7, 0), (8, 0)])
        assert got["n"] == 6 * (w["n_frames"] - 1) + 9 + 6 + 1 + 10


def test_margin_second_new(gf2, oracle, synth):
    """MARGIN_SECOND_NEW (estimator.cpp:3597-3690): only the old prior, dropping the second-newest pose."""
    abi = gf2.abi
    w = synth.make_windows(1, n_landmarks=60, prior="dense")   # dense prior holds poses 0..9 and speed-bias 0
    opts = abi.default_opts()
    F = w["n_frames"]
    w["para_pose"][0, :, :3] += 0.01                              # move off the linearisation point
    got = oracle.marginalize_window(w, 0, opts, mode=1)
    assert got["status"] == 0 and got["m"] == 6 and got["n"] == 6 * (F - 2) + 9
    n0 = int(w["prior_rows"][0]); J0 = w["prior_J0"][0, :n0, :n0]; r0 = w["prior_r0"][0, :n0]
    dx = np.zeros(n0)
    for blk in w["prior_blocks"][0, :int(w["prior_nblocks"][0])]:
        if blk["kind"] == 0:
            dx[blk["offset"]:blk["offset"] + 3] = 0.01
    res = r0 + J0 @ dx
    A = J0.T @ J0; b = J0.T @ res
    drop = np.arange(6 * (F - 2), 6 * (F - 2) + 6)   # pose F-2 = 9 sits at offset 6*9 in the dense prior
    keep = np.array([c for c in range(n0) if c not in drop])
    ev, V = np.linalg.eigh(0.5 * (A[np.ix_(drop, drop)] + A[np.ix_(drop, drop)].T))
    Ainv = (V / ev) @ V.T
    Ar = A[np.ix_(keep, keep)] - A[np.ix_(keep, drop)] @ Ainv @ A[np.ix_(drop, keep)]
    br = b[keep] - A[np.ix_(keep, drop)] @ Ainv @ b[drop]
    Ho, go, x0 = oracle.prior_information(got, F)
    # canonical columns of the kept prior columns: poses 0..8 keep their index, speed-bias 0 too
    cols = np.concatenate([np.concatenate([15 * f + np.arange(6) for f in range(F - 2)]), 6 + np.arange(9)])
    assert np.abs(Ho[np.ix_(cols, cols)] - Ar).max() <= 1e-9 * np.abs(Ar).max()
    assert np.abs(go[cols] - br).max() <= 1e-9 * max(1.0, np.abs(br).max())
    # without the second-newest pose in the prior nothing happens
    w2 = synth.make_windows(1, n_landmarks=60, prior="anchor")
    assert oracle.marginalize_window(w2, 0, opts, mode=1)["status"] == -2


def test_margin_old_rank_deficient_landmark(gf2, oracle, synth):
    """A landmark whose inverse-depth column vanishes (all its observations coincide with a pure-rotation geometry is not
    constructible here; instead the host observation is duplicated so v_l is tiny) must not poison the prior: eigenvalues
    below eps are truncated exactly as in marginalization_factor.cpp:281."""
    abi = gf2.abi
    w = synth.make_windows(1, n_landmarks=80, prior="anchor")
    oracle.imu_preintegrate(w)
    opts = abi.default_opts()
    w["inv_depth"][0, 0] = 1e-9     # landmark 0 (start frame 0) at "infinity": d proj / d lambda ~ 0 after scaling
    got = oracle.marginalize_window(w, 0, opts, mode=0)
    assert got["status"] == 0 and np.isfinite(got["J0"]).all() and np.isfinite(got["r0"]).all()
    H, g, m = numpy_marginalize_old(w, 0, oracle, abi, opts)
    Ho, go, _ = oracle.prior_information(got, w["n_frames"])
    assert np.abs(Ho - H).max() <= 1e-8 * np.abs(H).max()
