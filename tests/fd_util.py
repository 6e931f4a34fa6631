"""Finite-difference helpers following the reference's own check() recipe
(VE/factor/projectionTwoFrameOneCamFactor.cpp:214-269): eps = 1e-6, additive on positions/scalars,
right-multiplicative Q * deltaQ(delta) on rotations. Central differences are used for a tighter bound."""
import numpy as np


def quat_mul(a, b):  # [x y z w]
    ax, ay, az, aw = a; bx, by, bz, bw = b
    return np.array([aw * bx + ax * bw + ay * bz - az * by,
                     aw * by + ay * bw + az * bx - ax * bz,
                     aw * bz + az * bw + ax * by - ay * bx,
                     aw * bw - ax * bx - ay * by - az * bz])


def delta_q(theta):  # Utility::deltaQ
    q = np.array([theta[0] / 2, theta[1] / 2, theta[2] / 2, 1.0])
    return q / np.linalg.norm(q)


def plus(block, delta, kind):
    """kind: 'pose7' [p, qx qy qz qw] | 'vec' | 'quat4' [qx qy qz qw] alone"""
    b = np.array(block, dtype=float)
    if kind == "pose7":
        b[:3] += delta[:3]
        q = quat_mul(b[3:7], delta_q(delta[3:6])); b[3:7] = q / np.linalg.norm(q)
    elif kind == "quat4":
        q = quat_mul(b, delta_q(delta)); b = q / np.linalg.norm(q)
    else:
        b = b + delta
    return b


def numeric_jacobians(f, blocks, kinds, eps=1e-6):
    """f(list_of_blocks) -> residual vector. Returns per-block Jacobians in LOCAL coordinates."""
    r0 = f(blocks)
    out = []
    for i, (b, k) in enumerate(zip(blocks, kinds)):
        nloc = 6 if k == "pose7" else (3 if k == "quat4" else len(b))
        J = np.zeros((len(r0), nloc))
        for c in range(nloc):
            d = np.zeros(nloc); d[c] = eps
            bp = list(blocks); bp[i] = plus(b, d, k)
            bm = list(blocks); bm[i] = plus(b, -d, k)
            J[:, c] = (f(bp) - f(bm)) / (2 * eps)
        out.append(J)
    return out


def random_unit_quat(rng, max_angle=None):
    if max_angle is None:
        q = rng.normal(size=4)
    else:
        ax = rng.normal(size=3); ax /= np.linalg.norm(ax); ang = rng.uniform(0, max_angle)
        q = np.concatenate([np.sin(ang / 2) * ax, [np.cos(ang / 2)]])
    return q / np.linalg.norm(q)
