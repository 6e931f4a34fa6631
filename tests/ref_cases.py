"""Seeded inputs shared by tests/golden/make_ref_golden.py (which evaluates them with the REFERENCE's own factor code, oracle/_ref) and
tests/test_oracle_vs_ref.py (which evaluates them with the restated oracle, and on the GPU box with the device)."""
import numpy as np

from fd_util import plus, random_unit_quat

IMU_NOISE = np.array([0.012, 0.003, 1.9e-4, 5.4e-5])
WHEEL_NOISE = np.array([0.02, 0.01])


def projection_case(seed):
    rng = np.random.default_rng(100 + seed)
    pose_i = np.concatenate([rng.normal(size=3), random_unit_quat(rng)])
    pose_j = pose_i.copy(); pose_j[:3] += rng.normal(size=3) * 0.3; pose_j[3:] = plus(pose_j[3:], rng.normal(size=3) * 0.1, "quat4")
    ex = np.concatenate([rng.normal(size=3) * 0.05, random_unit_quat(rng, 0.1)])
    f32 = lambda v: float(np.float32(v))
    consts = np.array([f32(rng.normal() * 0.3), f32(rng.normal() * 0.3), f32(rng.normal() * 0.2), f32(rng.normal() * 0.2), 0.01 * rng.normal(),
                       f32(rng.normal() * 0.3), f32(rng.normal() * 0.3), f32(rng.normal() * 0.2), f32(rng.normal() * 0.2), 0.01 * rng.normal(), 400.0])
    params = np.concatenate([pose_i, pose_j, ex, [0.1 + abs(rng.normal()) * 0.3], [0.005 * rng.normal()]])
    return consts, params


def imu_samples(abi, seed, n=20):
    rng = np.random.default_rng(200 + seed)
    smp = np.zeros(n, abi.IMU_SAMPLE); smp["dt"] = 0.005
    smp["acc"] = rng.normal(size=(n, 3)) + [0, 0, 9.8]; smp["gyr"] = rng.normal(size=(n, 3)) * 0.3
    first = np.concatenate([smp["acc"][0] + rng.normal(size=3) * 0.1, smp["gyr"][0]]); lb = rng.normal(size=6) * 0.01
    return smp, first, lb


def imu_params(seed, lb):
    rng = np.random.default_rng(300 + seed)
    pi = np.concatenate([rng.normal(size=3), random_unit_quat(rng)])
    sbi = np.concatenate([rng.normal(size=3), lb[:3] + rng.normal(size=3) * 0.01, lb[3:] + rng.normal(size=3) * 0.001])
    pj = pi.copy(); pj[:3] += rng.normal(size=3) * 0.1; pj[3:] = plus(pj[3:], rng.normal(size=3) * 0.05, "quat4")
    sbj = sbi + rng.normal(size=9) * 0.01
    return np.concatenate([pi, sbi, pj, sbj])


def wheel_samples(abi, seed, n=5):
    rng = np.random.default_rng(400 + seed)
    smp = np.zeros(n, abi.WHEEL_SAMPLE); smp["dt"] = 0.02
    smp["vel"] = rng.normal(size=(n, 3)) * 0.05 + [1.0, 0, 0]; smp["gyr"] = rng.normal(size=(n, 3)) * 0.05 + [0, 0, 0.2]
    first = np.concatenate([smp["vel"][0] + rng.normal(size=3) * 0.01, smp["gyr"][0]])
    lin = np.array([1.0 + 0.02 * rng.normal(), 1.0 + 0.02 * rng.normal(), 1.0 + 0.02 * rng.normal(), 0.0])
    return smp, first, lin


def wheel_params(seed, lin, dtd=0.0):
    rng = np.random.default_rng(500 + seed)
    pi = np.concatenate([rng.normal(size=3), random_unit_quat(rng)])
    pj = pi.copy(); pj[:3] += rng.normal(size=3) * 0.1; pj[3:] = plus(pj[3:], rng.normal(size=3) * 0.05, "quat4")
    exw = np.concatenate([rng.normal(size=3) * 0.1, random_unit_quat(rng, 0.2)])
    return np.concatenate([pi, pj, exw, [lin[0] + 0.01 * rng.normal()], [lin[1] + 0.01 * rng.normal()], [lin[2] + 0.01 * rng.normal()], [lin[3] + dtd]])


def plane_case(seed, ct):
    rng = np.random.default_rng(600 + seed)
    nrm = rng.normal(size=3); nrm /= np.linalg.norm(nrm)
    p = rng.normal(size=3) * 3
    if not ct:
        consts = np.concatenate([p, nrm, [rng.normal()], [0.3 + rng.random()], [31.622776601683793]])
        params = np.concatenate([rng.normal(size=3), random_unit_quat(rng)])
    else:
        consts = np.concatenate([p, nrm, [rng.normal()], [rng.random()], [0.3 + rng.random()], [31.622776601683793]])
        q0 = random_unit_quat(rng)
        params = np.concatenate([rng.normal(size=3), q0, rng.normal(size=3), plus(q0, rng.normal(size=3) * 0.02, "quat4")])
    return consts, params
