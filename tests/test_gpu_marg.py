"""gf2_marginalize (k_marg_build + k_marg_eig) vs the restated MarginalizationInfo of the oracle.

Compared on the order-independent form of a prior (H = J0^T J0, g = J0^T r0 scattered by block, and the blocks' x0): the
reference's own block order is an unordered_map iteration order (see oracle/gf2o_marg.h)."""
import importlib

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def synth(gf2):
    if gf2.device_count() < 1:
        pytest.fail("no CUDA device: the hot path has no CPU fallback")
    return importlib.import_module("gf2_b200.synth")


@pytest.fixture(params=[0, 1], ids=["cholesky", "eigen"])
def marg_eig(request):
    """gf2_solve_opts.marg_eig: 0 = rank-revealing Cholesky factor of the kept system (default), 1 = the reference's eigen-decomposition, literally."""
    return request.param


def _prior_of(pr, i):
    n = int(pr["prior_rows"][i]); nb = int(pr["prior_nblocks"][i])
    return {"n": n, "J0": pr["prior_J0"][i, :n, :n], "r0": pr["prior_r0"][i, :n], "blocks": pr["prior_blocks"][i, :nb]}


def _check_prior(oracle, got, ref, F):
    """H to 1e-8 of its largest entry; g in units of the prior's own sigma (|dg_i| / sqrt(H_ii) <= 5e-6): g is a small
    difference of large terms at the solved state and Arm pinv(Amm) bmm carries cond(Amm) ~ 1e10 of rounding in BOTH
    implementations (eigen-decomposition pseudo-inverse in the oracle, structured elimination on the device)."""
    Hg, gg, xg = oracle.prior_information(got, F)
    Hr, gr, xr = oracle.prior_information(ref, F)
    scale = np.abs(Hr).max()
    assert np.abs(Hg - Hr).max() <= 1e-8 * scale, np.abs(Hg - Hr).max() / scale
    d = np.sqrt(np.diag(Hr)); nz = d > 0
    assert (np.abs(gg - gr)[nz] / d[nz]).max() <= 5e-6
    assert np.abs(gg - gr).max() <= 1e-5 * max(1.0, np.abs(gr).max())
    assert np.array_equal(gg[~nz], gr[~nz])
    assert set(xg) == set(xr)
    for key in xr:
        assert np.array_equal(xg[key], xr[key]), key
    assert got["n"] == ref["n"]


@pytest.mark.parametrize("prior,nl", [("anchor", 200), ("dense", 1000)])
def test_margin_old_matches_oracle(gf2, oracle, synth, prior, nl, marg_eig):
    n = 5
    w = synth.make_windows(n, n_landmarks=nl, prior=prior)
    oracle.imu_preintegrate(w)
    opts = gf2.abi.default_opts(); opts.marg_eig = marg_eig
    s = gf2.Solver(n, w["n_frames"], w["max_landmarks"], w["max_obs"])
    s.upload(w, preintegrate="records")
    s.solve(opts, n)
    st = s.get_states(n); lam = s.get_landmarks(n)
    status, m = s.marginalize(opts, mode=0)
    assert (status == 0).all(), status
    got = s.get_prior(n)
    # the oracle marginalizes at the SAME states (the device's solved ones), so only the marginalization itself is compared
    w["para_pose"][...] = st["para_pose"]; w["para_speedbias"][...] = st["para_speedbias"]; w["inv_depth"][...] = lam
    for i in range(n):
        ref = oracle.marginalize_window(w, i, opts, mode=0)
        assert ref["status"] == 0 and ref["m"] == m[i]
        _check_prior(oracle, _prior_of(got, i), ref, w["n_frames"])
    s.close()


def test_margin_old_with_wheel_matches_oracle(gf2, oracle, synth, marg_eig):
    """config 4 composition: IMU + wheel + projection (+ planes, which never touch frame 0's dropped blocks... they do touch
    pose 0 but the reference does not marginalize LiDAR factors in the VINS window) -> the wheel factor's calibration blocks
    (body_T_wheel, sx, sy, sw, td_wheel) become kept blocks of the prior."""
    n = 4
    w = synth.make_windows(n, config_id=4, n_landmarks=300, wheel=True, prior="dense")
    oracle.imu_preintegrate(w); oracle.wheel_preintegrate(w)
    w["sxsysw"][1] = [1.01, 0.99, 1.02]; w["td_wheel"][2] = 0.004
    opts = gf2.abi.default_opts(); opts.marg_eig = marg_eig
    s = gf2.Solver(n, w["n_frames"], w["max_landmarks"], w["max_obs"], use_wheel=True)
    s.upload(w, preintegrate="records")
    s.solve(opts, n)
    st = s.get_states(n); lam = s.get_landmarks(n)
    status, m = s.marginalize(opts, mode=0)
    assert (status == 0).all(), status
    got = s.get_prior(n)
    w["para_pose"][...] = st["para_pose"]; w["para_speedbias"][...] = st["para_speedbias"]; w["inv_depth"][...] = lam
    for i in range(n):
        ref = oracle.marginalize_window(w, i, opts, mode=0)
        assert ref["status"] == 0 and ref["m"] == m[i] and ref["n"] == 86
        _check_prior(oracle, _prior_of(got, i), ref, w["n_frames"])
    s.close()


def test_prior_chain_stays_resident(gf2, oracle, synth, marg_eig):
    """solve -> marginalize -> (slide) -> solve again with the device-resident prior == oracle doing the same on the host."""
    n = 3
    w = synth.make_windows(n, n_landmarks=300, prior="anchor")
    oracle.imu_preintegrate(w)
    opts = gf2.abi.default_opts(); opts.marg_eig = marg_eig
    s = gf2.Solver(n, w["n_frames"], w["max_landmarks"], w["max_obs"])
    s.upload(w, preintegrate="records")
    s.solve(opts, n)
    st = s.get_states(n); lam = s.get_landmarks(n)
    status, _ = s.marginalize(opts, mode=0)
    assert (status == 0).all()
    # "next window": a fresh set of windows (their frames play the role of the slid window); the prior links them to x0
    w2 = synth.make_windows(n, n_landmarks=300, prior="anchor", first_window=50)
    oracle.imu_preintegrate(w2)
    s.set_states(w2); s.set_landmarks(w2); s.set_imu(w2["imu"])
    s.solve(opts, n)
    got = s.get_states(n)
    # host side: oracle marginalization at the same states, prior arrays handed to the oracle solve
    w["para_pose"][...] = st["para_pose"]; w["para_speedbias"][...] = st["para_speedbias"]; w["inv_depth"][...] = lam
    P = w2["prior_J0"].shape[1]
    for i in range(n):
        ref = oracle.marginalize_window(w, i, opts, mode=0)
        nn = ref["n"]; nb = len(ref["blocks"])
        w2["prior_rows"][i] = nn; w2["prior_nblocks"][i] = nb
        w2["prior_J0"][i] = 0; w2["prior_J0"][i, :nn, :nn] = ref["J0"]; w2["prior_r0"][i] = 0; w2["prior_r0"][i, :nn] = ref["r0"]
        w2["prior_blocks"][i, :nb] = ref["blocks"]
    oracle.solve_batch(w2, opts)
    scale = np.abs(w2["para_pose"]).max()
    assert np.abs(got["para_pose"] - w2["para_pose"]).max() <= 1e-4 * scale
    s.close()


def test_margin_second_new_matches_oracle(gf2, oracle, synth, marg_eig):
    n = 3
    w = synth.make_windows(n, n_landmarks=100, prior="dense")
    opts = gf2.abi.default_opts(); opts.marg_eig = marg_eig
    oracle.imu_preintegrate(w)
    w["para_pose"][:, :, :3] += 0.01
    s = gf2.Solver(n, w["n_frames"], w["max_landmarks"], w["max_obs"])
    s.upload(w, preintegrate="records")
    status, m = s.marginalize(opts, mode=1)
    assert (status == 0).all() and (m == 6).all()
    got = s.get_prior(n)
    for i in range(n):
        ref = oracle.marginalize_window(w, i, opts, mode=1)
        _check_prior(oracle, _prior_of(got, i), ref, w["n_frames"])
    # an anchor prior does not hold the second-newest pose: nothing happens (estimator.cpp:3599)
    w2 = synth.make_windows(n, n_landmarks=100, prior="anchor")
    oracle.imu_preintegrate(w2)
    s.upload(w2, preintegrate="records")
    before = s.get_prior(n)
    status, _ = s.marginalize(opts, mode=1)
    assert (status == gf2.abi.MARG_UNCHANGED).all()
    after = s.get_prior(n)
    assert np.array_equal(before["prior_J0"], after["prior_J0"]) and np.array_equal(before["prior_rows"], after["prior_rows"])
    s.close()


def test_margin_old_edge_cases(gf2, oracle, synth, marg_eig):
    """No landmark hosted in frame 0; no prior; no IMU -> m == 0 -> invalid prior (valid = false, marginalization_factor.cpp:205)."""
    n = 2
    w = synth.make_windows(n, n_landmarks=120, prior="anchor")
    oracle.imu_preintegrate(w)
    opts = gf2.abi.default_opts(); opts.marg_eig = marg_eig
    # drop the frame-0 landmarks by compacting the landmark arrays
    for i in range(n):
        nl = int(w["n_landmarks"][i]); sel = np.where(w["start_frame"][i, :nl] != 0)[0]
        obeg = np.concatenate([[0], np.cumsum(w["track_len"][i, :nl])])
        obs = np.concatenate([w["obs"][i, obeg[l]:obeg[l + 1]] for l in sel])
        w["obs"][i, :len(obs)] = obs
        for name in ("start_frame", "track_len", "inv_depth", "fixed"):
            w[name][i, :len(sel)] = w[name][i, sel]
        w["n_landmarks"][i] = len(sel)
    s = gf2.Solver(n, w["n_frames"], w["max_landmarks"], w["max_obs"])
    s.upload(w, preintegrate="records")
    status, m = s.marginalize(opts, mode=0)
    assert (status == 0).all() and (m == 15).all()
    got = s.get_prior(n)
    from test_oracle_marg import assemble_margin_old
    for i in range(n):
        ref = oracle.marginalize_window(w, i, opts, mode=0)
        assert ref["m"] == 15 and ref["n"] == got["prior_rows"][i] == 15
        # Here the kept information (~1) is what is left of the IMU factor's (~1e8) after the Schur complement: both the
        # device (Cholesky) and the oracle (eigen pseudo-inverse) lose ~cond(Amm) * 1e-16 of it. Judge both against the same
        # elimination carried out in extended precision.
        A, b, mm, rr, T = assemble_margin_old(w, i, oracle, opts)
        L = np.linalg.cholesky(A[np.ix_(mm, mm)]).astype(np.longdouble)
        Amm = A[np.ix_(mm, mm)].astype(np.longdouble)
        for _ in range(3):   # refine the factor in long double (Cholesky in place)
            Lx = np.zeros_like(Amm)
            for j in range(len(mm)):
                d = Amm[j, j] - (Lx[j, :j] ** 2).sum()
                Lx[j, j] = np.sqrt(d)
                for r in range(j + 1, len(mm)):
                    Lx[r, j] = (Amm[r, j] - (Lx[r, :j] * Lx[j, :j]).sum()) / Lx[j, j]
            L = Lx
        X = A[np.ix_(mm, rr)].astype(np.longdouble)
        Z = np.zeros_like(X)
        for j in range(len(mm)):
            Z[j] = (X[j] - (L[j, :j, None] * Z[:j]).sum(axis=0)) / L[j, j]
        Htrue = (A[np.ix_(rr, rr)].astype(np.longdouble) - Z.T @ Z).astype(np.float64)
        new = np.where(rr < 15 * w["n_frames"], rr - 15, rr)
        Hg, _, _ = oracle.prior_information(_prior_of(got, i), w["n_frames"])
        Hr, _, _ = oracle.prior_information(ref, w["n_frames"])
        err_dev = np.abs(Hg[np.ix_(new, new)] - Htrue).max(); err_orc = np.abs(Hr[np.ix_(new, new)] - Htrue).max()
        assert err_dev <= max(2.0 * err_orc, 1e-8 * np.abs(Htrue).max()), (err_dev, err_orc)
        assert err_dev <= 1e-9 * np.abs(A[np.ix_(rr, rr)]).max()
    # now also without prior and without IMU: m == 0
    w["prior_rows"][...] = 0; w["prior_nblocks"][...] = 0
    w["imu"]["valid"][...] = 0
    s.upload(w, preintegrate="records")
    status, m = s.marginalize(opts, mode=0)
    assert (status == gf2.abi.MARG_INVALID).all() and (m == 0).all()
    assert (s.get_prior(n)["prior_rows"] == 0).all()
    for i in range(n):
        assert oracle.marginalize_window(w, i, opts, mode=0)["status"] == -1
    s.close()


def test_margin_old_truncates_uninformative_landmarks_like_the_pseudo_inverse(gf2, oracle, synth, marg_eig):
    """A robot turning on the spot: zero baseline, d r / d lambda = 0 for every landmark, so their eigenvalues of Amm are below eps = 1e-8 and
    MarginalizationInfo::marginalize() truncates them in its pseudo-inverse (marginalization_factor.cpp:278-283). The device drops exactly
    those landmarks' Schur terms (status 0, not GF2_MARG_DEGENERATE) and must agree with the restated eigen-decomposition pseudo-inverse.
    Window 1 is left untouched (informative landmarks) as the control."""
    n = 2
    w = synth.make_windows(n, n_landmarks=160, prior="anchor")
    oracle.imu_preintegrate(w)
    w["para_pose"][0, :, :3] = w["para_pose"][0, 0, :3]   # all camera centres coincide: no translation between the frames ...
    w["ex_pose"][0, :3] = 0.0                             # ... and no lever arm
    opts = gf2.abi.default_opts(); opts.marg_eig = marg_eig
    s = gf2.Solver(n, w["n_frames"], w["max_landmarks"], w["max_obs"])
    s.upload(w, preintegrate="records")
    status, m = s.marginalize(opts, mode=0)
    assert (status == 0).all(), status
    got = s.get_prior(n)
    for i in range(n):
        ref = oracle.marginalize_window(w, i, opts, mode=0)
        assert ref["status"] == 0 and ref["m"] == m[i]
        _check_prior(oracle, _prior_of(got, i), ref, w["n_frames"])
    s.close()
