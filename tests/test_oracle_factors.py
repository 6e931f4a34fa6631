"""CPU tests of the oracle's factor restatements: analytic Jacobians vs finite differences with the reference's own
check() recipe (VE/factor/projectionTwoFrameOneCamFactor.cpp:214-269), and the only known-answer fixture the reference
holds on this path: LIO/apps/test_analytic_factor.cpp:56-66 (tolerance 1e-6 at :134)."""
import numpy as np
import pytest

from fd_util import numeric_jacobians, plus, random_unit_quat


def _check(orc, kind, consts, blocks, kinds, extra=None, tol=1e-6, skip=()):
    res, J = orc.factor_eval(kind, consts, np.concatenate(blocks), extra)

    def f(bl):
        return orc.factor_eval(kind, consts, np.concatenate(bl), extra, want_jac=False)[0]
    Jn = numeric_jacobians(f, blocks, kinds)
    for i, (a, n) in enumerate(zip(J, Jn)):
        if i in skip:
            continue
        err = np.abs(a[:, :n.shape[1]] - n).max() / max(1.0, np.abs(n).max())
        assert err < tol, f"kind {kind} block {i}: rel err {err}"
        if a.shape[1] > n.shape[1]:  # ambient quaternion-w column must be zero (cpp:104-111)
            assert np.abs(a[:, n.shape[1]:]).max() == 0.0
    return res, J


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_projection_factor_jacobians(oracle, seed):
    rng = np.random.default_rng(seed)
    pose_i = np.concatenate([rng.normal(size=3), random_unit_quat(rng)])
    pose_j = pose_i.copy(); pose_j[:3] += rng.normal(size=3) * 0.3; pose_j[3:] = plus(pose_j[3:], rng.normal(size=3) * 0.1, "quat4")
    ex = np.concatenate([rng.normal(size=3) * 0.05, random_unit_quat(rng, 0.1)])
    consts = np.array([0.1, -0.2, 0.3, 0.1, 0.01, 0.12, -0.18, 0.2, -0.3, 0.02, 400.0])
    _check(oracle, 0, consts, [pose_i, pose_j, ex, np.array([0.2]), np.array([0.005])], ["pose7", "pose7", "pose7", "vec", "vec"])


def _imu_record(orc, abi, rng, n=20):
    smp = np.zeros(n, abi.IMU_SAMPLE); smp["dt"] = 0.005
    smp["acc"] = rng.normal(size=(n, 3)) + [0, 0, 9.8]; smp["gyr"] = rng.normal(size=(n, 3)) * 0.3
    rec = np.zeros(1, abi.IMU_PREINT)
    first = np.concatenate([smp["acc"][0], smp["gyr"][0]]); lb = rng.normal(size=6) * 0.01
    noise = np.array([0.012, 0.003, 1.9e-4, 5.4e-5])
    orc.lib.gf2o_imu_preintegrate(orc._p(smp), n, orc._p(first), orc._p(lb), orc._p(noise), orc._p(rec))
    return rec, lb


@pytest.mark.parametrize("seed", [0, 1])
def test_imu_factor_jacobians(oracle, gf2, seed):
    rng = np.random.default_rng(seed)
    rec, lb = _imu_record(oracle, gf2.abi, rng)
    pi = np.concatenate([rng.normal(size=3), random_unit_quat(rng)])
    sbi = np.concatenate([rng.normal(size=3), lb[:3] + rng.normal(size=3) * 0.01, lb[3:] + rng.normal(size=3) * 0.001])
    pj = pi.copy(); pj[:3] += rng.normal(size=3) * 0.1; pj[3:] = plus(pj[3:], rng.normal(size=3) * 0.05, "quat4")
    sbj = sbi + rng.normal(size=9) * 0.01
    _check(oracle, 1, rec, [pi, sbi, pj, sbj], ["pose7", "vec", "pose7", "vec"], extra=[9.8], tol=1e-6)


def test_imu_preintegration_properties(oracle, gf2):
    """covariance symmetric PSD, jacobian block structure (identity diagonal blocks), delta_q unit."""
    rng = np.random.default_rng(3)
    rec, _ = _imu_record(oracle, gf2.abi, rng)
    C = rec["covariance"][0].reshape(15, 15); J = rec["jacobian"][0].reshape(15, 15)
    assert np.abs(C - C.T).max() < 1e-12 * np.abs(C).max()
    assert np.linalg.eigvalsh(C).min() > 0
    assert np.allclose(J[9:12, 9:12], np.eye(3)) and np.allclose(J[12:15, 12:15], np.eye(3)) and np.allclose(J[0:3, 0:3], np.eye(3))
    assert abs(np.linalg.norm(rec["delta_q"][0]) - 1) < 1e-14
    assert abs(rec["sum_dt"][0] - 0.1) < 1e-15


def test_wheel_factor_jacobians(oracle, gf2):
    abi = gf2.abi
    rng = np.random.default_rng(5)
    n = 5; ws = np.zeros(n, abi.WHEEL_SAMPLE); ws["dt"] = 0.02
    ws["vel"] = rng.normal(size=(n, 3)) * 0.1 + [1, 0, 0]; ws["gyr"] = rng.normal(size=(n, 3)) * 0.1 + [0, 0, 0.2]
    wrec = np.zeros(1, abi.WHEEL_PREINT); first = np.concatenate([ws["vel"][0], ws["gyr"][0]])
    oracle.lib.gf2o_wheel_preintegrate(oracle._p(ws), n, oracle._p(first), oracle._p(np.array([1.0, 1.0, 1.0, 0.0])), oracle._p(np.array([0.01, 0.004])), oracle._p(wrec))
    pi = np.concatenate([rng.normal(size=3), random_unit_quat(rng)])
    pj = pi.copy(); pj[:3] += rng.normal(size=3) * 0.1; pj[3:] = plus(pj[3:], rng.normal(size=3) * 0.05, "quat4")
    exw = np.concatenate([rng.normal(size=3) * 0.1, random_unit_quat(rng)])
    kinds = ["pose7", "pose7", "pose7", "vec", "vec", "vec", "vec"]
    # at dtd = 0 and s = linearisation point every block is exact
    _check(oracle, 2, wrec, [pi, pj, exw, np.array([1.0]), np.array([1.0]), np.array([1.0]), np.array([0.0])], kinds)
    # away from the linearisation point the sx/sy/sw/td blocks are the reference's first-order approximations
    # (SURVEY Appendix C): only the pose / extrinsic blocks are checked
    _check(oracle, 2, wrec, [pi, pj, exw, np.array([1.02]), np.array([0.99]), np.array([1.01]), np.array([0.01])], kinds, skip=(3, 4, 5, 6))


def test_lidar_plane_known_answer(oracle):
    """Inputs of LIO/apps/test_analytic_factor.cpp:56-66; the reference accepts |analytic - autodiff| <= 1e-6."""
    nv = np.array([0.3, 1.5, -2.0]); nv /= np.linalg.norm(nv)
    neigh = np.array([1.0, 3.0, 5.0]); off = -nv @ neigh
    q = np.array([0.6, 1.3, -0.9, 0.2]); q /= np.linalg.norm(q)  # Quaterniond(0.2, 0.6, 1.3, -0.9) as [x y z w]
    t = np.array([11.0, 13.0, 15.0]); p = np.array([10.0, 12.0, 14.0])
    consts = np.concatenate([p, nv, [off, 1.0, 1.0]])
    res, J = _check(oracle, 3, consts, [t, q], ["vec", "quat4"], tol=1e-6)
    # independent closed form
    x, y, z, w = q
    R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                  [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                  [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
    assert abs(res[0] - (nv @ (R @ p + t) + off)) < 1e-12
    assert np.abs(J[0][0] - nv).max() < 1e-15
    sk = np.array([[0, -p[2], p[1]], [p[2], 0, -p[0]], [-p[1], p[0], 0]])
    assert np.abs(J[1][0, :3] - (-nv @ R @ sk)).max() < 1e-12 and J[1][0, 3] == 0.0
    # golden values pinned from this fixture
    assert abs(res[0] - 11.27908089) < 1e-7
    assert np.abs(J[1][0] - np.array([-13.45930147, 10.62172794, 0.50944853, 0.0])).max() < 1e-7


def test_plane_pose_factor_matches_lidar_factor(oracle):
    rng = np.random.default_rng(7)
    nv = rng.normal(size=3); nv /= np.linalg.norm(nv)
    consts = np.concatenate([rng.normal(size=3), nv, [0.3, 1.0, 31.6]])
    pose = np.concatenate([rng.normal(size=3), random_unit_quat(rng)])
    _check(oracle, 5, consts, [pose], ["pose7"])


def test_sym_eigen_matches_numpy(oracle):
    rng = np.random.default_rng(11)
    A = rng.normal(size=(40, 40)); A = A @ A.T
    ev = np.zeros(40); V = np.zeros((40, 40))
    oracle.lib.gf2o_sym_eigen(40, oracle._p(np.ascontiguousarray(A)), oracle._p(ev), oracle._p(V))
    assert np.allclose(ev, np.linalg.eigvalsh(A), rtol=1e-10, atol=1e-10 * ev.max())
    assert np.abs(V @ np.diag(ev) @ V.T - A).max() < 1e-9 * np.abs(A).max()


@pytest.mark.parametrize("seed", [0, 1])
def test_ct_lidar_plane_factor_translation_exact_rotation_first_order(oracle, seed):
    """CTLidarPlaneNormFactor (LIO/liw/lidarFactor.cpp:58-123). The reference ships no test for it (test_analytic_factor.cpp covers only
    the non-CT factor). Its translation Jacobians are exact; its rotation Jacobians are first order in the begin-end rotation difference
    (restated verbatim): their error against finite differences is O(|delta|) (1e-6 at 1e-4 rad, < 1 % at 0.05 rad), which this test pins so that a transcription error would show."""
    rng = np.random.default_rng(seed)
    nv = rng.normal(size=3); nv /= np.linalg.norm(nv)
    alpha = rng.uniform(0.1, 0.9)
    consts = np.concatenate([rng.normal(size=3), nv, [0.3, alpha, 0.8, 31.6]])
    tb = rng.normal(size=3); qb = random_unit_quat(rng)
    for delta, tol in ((1e-4, 1e-6), (0.05, 1e-2)):
        te = tb + rng.normal(size=3) * 0.1; qe = plus(qb, rng.normal(size=3) * delta, "quat4")
        blocks = [tb, qb, te, qe]; kinds = ["vec", "quat4", "vec", "quat4"]
        res, J = oracle.factor_eval(4, consts, np.concatenate(blocks))
        Jn = numeric_jacobians(lambda bl: oracle.factor_eval(4, consts, np.concatenate(bl), want_jac=False)[0], blocks, kinds)
        for i in (0, 2):
            assert np.abs(J[i] - Jn[i]).max() < 1e-8 * max(1.0, np.abs(Jn[i]).max())
        for i in (1, 3):
            assert np.abs(J[i][:, :3] - Jn[i]).max() < tol * max(1.0, np.abs(Jn[i]).max()) and J[i][0, 3] == 0.0
        # the two-pose wrapper used in the window composition carries the same numbers in 7-blocks
        r7, J7 = oracle.factor_eval(6, consts, np.concatenate([tb, qb, te, qe]))
        assert r7[0] == res[0]
        assert np.array_equal(J7[0][0, :3], J[0][0]) and np.array_equal(J7[0][0, 3:6], J[1][0, :3]) and J7[0][0, 6] == 0.0
        assert np.array_equal(J7[1][0, :3], J[2][0]) and np.array_equal(J7[1][0, 3:6], J[3][0, :3]) and J7[1][0, 6] == 0.0
        # alpha = 0 / 1 reduce to the begin / end pose
    for a, (t, q) in ((0.0, (tb, qb)), (1.0, (te, qe))):
        c = consts.copy(); c[7] = a
        r_ct = oracle.factor_eval(4, c, np.concatenate([tb, qb, te, qe]), want_jac=False)[0]
        r_pl = oracle.factor_eval(3, np.concatenate([c[:7], [c[8], c[9]]]), np.concatenate([t, q]), want_jac=False)[0]
        assert abs(r_ct[0] - r_pl[0]) < 1e-12 * max(1.0, abs(r_pl[0]))
