"""Pins the restated oracle against the REFERENCE's own factor code.

(1) live: where /root/reference is present (the build container) oracle/_ref/libgf2_ref.so holds the reference's factor sources compiled
    unmodified (oracle/Makefile; Eigen / Ceres / ROS / Sophus headers replaced by the stand-ins of oracle/shim); the restated oracle
    must reproduce its residuals, Jacobians, preintegration records and marginalization priors on the same inputs.
(2) golden: the outputs of that library on seeded inputs are committed as tests/golden/ref_golden.npz (made by
    tests/golden/make_ref_golden.py) and checked everywhere, GPU box included.
Tolerances: 1e-12 relative for closed-form factors; 1e-9 where the reference inverts a 15x15 / 6x6 covariance (sqrt_info =
LLT(cov^-1), the inverse amplifies rounding by the covariance's condition number); marginalization priors compared in the
order-independent form J0^T J0 / J0^T r0 (the reference keys its block tables by address in an unordered_map)."""
import importlib
import os

import numpy as np
import pytest

import ref_cases as rc

GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_golden.npz"))


def _ref():
    import gf2_ref
    if not gf2_ref.available():
        pytest.skip("oracle/_ref is only built where /root/reference is present")
    return gf2_ref


def _close(a, b, rel, what):
    scale = max(np.abs(b).max(), 1e-300)
    err = np.abs(np.asarray(a) - np.asarray(b)).max() / scale
    assert err <= rel, f"{what}: relative error {err:.3e} > {rel:.1e}"


def _eval_both(kind, consts, params, extra, oracle, other, rel, what):
    r, J = oracle.factor_eval(kind, consts, params, extra)
    r2, J2 = other
    _close(r, r2, rel, what + " residual")
    for b, (a, c) in enumerate(zip(J, J2)):
        scale = max(np.abs(c).max(), np.abs(np.concatenate([j.ravel() for j in J2])).max() * 1e-6)
        assert np.abs(a - c).max() <= rel * scale, f"{what} jacobian block {b}: {np.abs(a - c).max() / scale:.3e}"


# ------------------------------------------------------------------------------------------------ (2) golden vectors of the reference code
@pytest.mark.parametrize("s", range(4))
def test_projection_factor_matches_reference_golden(oracle, s):
    c, p = rc.projection_case(s)
    _eval_both(0, c, p, None, oracle, (GOLD[f"proj{s}_res"], [GOLD[f"proj{s}_J{b}"] for b in range(5)]), 1e-12, "projection")


@pytest.mark.parametrize("s", range(3))
def test_imu_preintegration_and_factor_match_reference_golden(oracle, gf2, s):
    abi = gf2.abi
    smp, first, lb = rc.imu_samples(abi, s)
    rec = np.zeros(1, abi.IMU_PREINT)
    oracle.lib.gf2o_imu_preintegrate(oracle._p(smp), len(smp), oracle._p(first), oracle._p(lb), oracle._p(rc.IMU_NOISE), oracle._p(rec))
    for f in ("sum_dt", "delta_p", "delta_q", "delta_v", "lin_ba", "lin_bg", "jacobian", "covariance"):
        _close(rec[f][0], GOLD[f"imu{s}_{f}"], 1e-12, "IntegrationBase." + f)
    _eval_both(1, rec, rc.imu_params(s, lb), [9.7944], oracle, (GOLD[f"imu{s}_res"], [GOLD[f"imu{s}_J{b}"] for b in range(4)]), 1e-9, "IMUFactor")


@pytest.mark.parametrize("s", range(3))
def test_wheel_preintegration_and_factor_match_reference_golden(oracle, gf2, s):
    abi = gf2.abi
    smp, first, lin = rc.wheel_samples(abi, s)
    rec = np.zeros(1, abi.WHEEL_PREINT)
    oracle.lib.gf2o_wheel_preintegrate(oracle._p(smp), len(smp), oracle._p(first), oracle._p(lin), oracle._p(rc.WHEEL_NOISE), oracle._p(rec))
    for f in ("sum_dt", "delta_p", "delta_q", "jacobian", "covariance", "vel_1", "gyr_1", "lin_vel", "lin_gyr"):
        _close(rec[f][0], GOLD[f"wheel{s}_{f}"], 1e-12, "WheelIntegrationBase." + f)
    for tag, dtd in (("a", 0.0), ("b", 0.004)):
        _eval_both(2, rec, rc.wheel_params(s, lin, dtd), None, oracle, (GOLD[f"wheel{s}{tag}_res"], [GOLD[f"wheel{s}{tag}_J{b}"] for b in range(7)]), 1e-9, "WheelFactor " + tag)


@pytest.mark.parametrize("s", range(3))
@pytest.mark.parametrize("ct", [0, 1])
def test_lidar_plane_factors_match_reference_golden(oracle, s, ct):
    c, p = rc.plane_case(s, ct)
    _eval_both(3 + ct, c, p, None, oracle, (GOLD[f"plane{s}_{ct}_res"], [GOLD[f"plane{s}_{ct}_J{b}"] for b in range(2 + 2 * ct)]), 1e-11, "lidar plane")


def _second_generation(abi, w, i, m0):
    w2 = {k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in w.items()}
    P = 96
    w2["prior_rows"][i] = m0["n"]
    J0 = np.zeros((2, P, P)); r0 = np.zeros((2, P)); J0[i, :m0["n"], :m0["n"]] = m0["J0"]; r0[i, :m0["n"]] = m0["r0"]
    blocks = np.zeros((2, 2 * 11 + 8), abi.PRIOR_BLOCK); blocks[i, :len(m0["blocks"])] = m0["blocks"]
    nb = w2["prior_nblocks"].copy(); nb[i] = len(m0["blocks"])
    w2["prior_J0"] = J0; w2["prior_r0"] = r0; w2["prior_blocks"] = blocks; w2["prior_nblocks"] = nb
    return w2


def _marg_windows(oracle, tag):
    synth = importlib.import_module("gf2_b200.synth")
    kw = dict(n_landmarks=120) if tag == "vio" else dict(n_landmarks=100, wheel=True, config_id=4)
    w = synth.make_windows(2, **kw)
    oracle.imu_preintegrate(w)
    if tag == "wheel":
        oracle.wheel_preintegrate(w)
    return w


def _check_prior(oracle, got, H, g, nm, what):
    assert got["status"] == 0 and [got["n"], got["m"]] == list(nm), what
    Hg, gg, _ = oracle.prior_information(got, 11)
    assert np.abs(Hg - H).max() <= 1e-9 * np.abs(H).max(), f"{what}: J0^T J0 differs by {np.abs(Hg - H).max() / np.abs(H).max():.3e}"
    # J0^T r0 in units of the prior's own sigma: |dg_i| / sqrt(H_ii)
    d = np.abs(gg - g) / np.sqrt(np.maximum(np.diag(H), 1e-300))
    assert d[np.diag(H) > 0].max() <= 1e-5, f"{what}: J0^T r0 differs by {d.max():.3e} sigma"


@pytest.mark.parametrize("tag", ["vio", "wheel"])
def test_marginalization_matches_reference_golden(oracle, gf2, tag):
    """MarginalizationInfo::{preMarginalize, marginalize} of the reference (golden) vs the restatement: MARGIN_OLD from the anchor prior, then
    MARGIN_SECOND_NEW and MARGIN_OLD with the first result as last_marginalization_info."""
    opts = gf2.abi.default_opts()
    w = _marg_windows(oracle, tag)
    for i in range(2):
        m0 = oracle.marginalize_window(w, i, opts, mode=0)
        _check_prior(oracle, m0, GOLD[f"marg_{tag}{i}_old_H"], GOLD[f"marg_{tag}{i}_old_g"], GOLD[f"marg_{tag}{i}_old_n"], f"{tag}{i} MARGIN_OLD")
        _, _, x0 = oracle.prior_information(m0, 11)
        for k, v in x0.items():
            assert np.array_equal(v, GOLD[f"marg_{tag}{i}_old_x0_{k[0]}_{k[1]}"]), "linearisation points must be bit-equal"
        w2 = _second_generation(gf2.abi, w, i, m0)
        for mode, name in ((1, "second"), (0, "old2")):
            m1 = oracle.marginalize_window(w2, i, opts, mode=mode)
            _check_prior(oracle, m1, GOLD[f"marg_{tag}{i}_{name}_H"], GOLD[f"marg_{tag}{i}_{name}_g"], GOLD[f"marg_{tag}{i}_{name}_n"], f"{tag}{i} {name}")


# ------------------------------------------------------------------------------------------------ (1) live against oracle/_ref
def test_golden_vectors_are_what_the_reference_code_produces(gf2):
    """The committed golden file equals a fresh evaluation by oracle/_ref (guards against a stale .npz)."""
    ref = _ref()
    for s in range(4):
        c, p = rc.projection_case(s)
        r, J = ref.factor_eval(0, c, p)
        assert np.array_equal(r, GOLD[f"proj{s}_res"]) and all(np.array_equal(j, GOLD[f"proj{s}_J{b}"]) for b, j in enumerate(J))
    x, jac = ref.pose_plus(np.array([0.1, -0.2, 0.3, 0.1, 0.2, -0.3, 0.9273618495495703]), np.array([0.01, 0.02, -0.03, 0.05, -0.04, 0.02]))
    assert np.array_equal(x, GOLD["plus_x"]) and np.array_equal(jac, GOLD["plus_jac"])


@pytest.mark.parametrize("seed", range(10, 18))
def test_factors_match_reference_live(oracle, gf2, seed):
    ref = _ref(); abi = gf2.abi
    c, p = rc.projection_case(seed)
    _eval_both(0, c, p, None, oracle, ref.factor_eval(0, c, p), 1e-12, "projection")
    smp, first, lb = rc.imu_samples(abi, seed, n=7 + seed)
    rec = ref.imu_preintegrate_one(abi, smp, len(smp), first, lb, rc.IMU_NOISE)
    mine = np.zeros(1, abi.IMU_PREINT)
    oracle.lib.gf2o_imu_preintegrate(oracle._p(smp), len(smp), oracle._p(first), oracle._p(lb), oracle._p(rc.IMU_NOISE), oracle._p(mine))
    for f in ("sum_dt", "delta_p", "delta_q", "delta_v", "jacobian", "covariance"):
        _close(mine[f][0], rec[f][0], 1e-12, "IntegrationBase." + f)
    pp = rc.imu_params(seed, lb)
    _eval_both(1, rec, pp, [9.7944], oracle, ref.factor_eval(1, rec, pp, extra=[9.7944]), 1e-9, "IMUFactor")
    smp, first, lin = rc.wheel_samples(abi, seed)
    rec = ref.wheel_preintegrate_one(abi, smp, len(smp), first, lin, rc.WHEEL_NOISE)
    mine = np.zeros(1, abi.WHEEL_PREINT)
    oracle.lib.gf2o_wheel_preintegrate(oracle._p(smp), len(smp), oracle._p(first), oracle._p(lin), oracle._p(rc.WHEEL_NOISE), oracle._p(mine))
    for f in ("sum_dt", "delta_p", "delta_q", "jacobian", "covariance"):
        _close(mine[f][0], rec[f][0], 1e-12, "WheelIntegrationBase." + f)
    pp = rc.wheel_params(seed, lin, 0.003 * (seed % 3))
    _eval_both(2, rec, pp, None, oracle, ref.factor_eval(2, rec, pp), 1e-9, "WheelFactor")
    for ct in (0, 1):
        c, p = rc.plane_case(seed, ct)
        _eval_both(3 + ct, c, p, None, oracle, ref.factor_eval(3 + ct, c, p), 1e-11, "lidar plane")


def test_pose_local_parameterization_matches_reference_live(gf2):
    """PoseLocalParameterization::Plus / ComputeJacobian of the reference vs the retraction the tests' FD helper (and the device) use."""
    ref = _ref()
    from fd_util import plus, random_unit_quat
    rng = np.random.default_rng(5)
    for _ in range(5):
        x = np.concatenate([rng.normal(size=3), random_unit_quat(rng)]); d = rng.normal(size=6) * 0.1
        got, jac = ref.pose_plus(x, d)
        assert np.abs(got - plus(x, d, "pose7")).max() < 1e-15
        assert np.array_equal(jac, np.vstack([np.eye(6), np.zeros((1, 6))]))


@pytest.mark.parametrize("tag", ["vio", "wheel"])
def test_marginalization_matches_reference_live(oracle, gf2, tag):
    ref = _ref()
    opts = gf2.abi.default_opts()
    w = _marg_windows(oracle, tag)
    for i in range(2):
        r0 = ref.marginalize_window(w, i, opts, mode=0)
        H, g, x0 = oracle.prior_information(r0, 11)
        m0 = oracle.marginalize_window(w, i, opts, mode=0)
        _check_prior(oracle, m0, H, g, [r0["n"], r0["m"]], f"{tag}{i} MARGIN_OLD")
        w2 = _second_generation(gf2.abi, w, i, r0)
        for mode in (1, 0):
            r1 = ref.marginalize_window(w2, i, opts, mode=mode)
            H, g, _ = oracle.prior_information(r1, 11)
            _check_prior(oracle, oracle.marginalize_window(w2, i, opts, mode=mode), H, g, [r1["n"], r1["m"]], f"{tag}{i} second generation mode {mode}")
