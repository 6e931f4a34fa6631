"""Synthetic sliding windows in the reference's packed wire layout.

Pure numpy, deterministic (seed = 20250925 + config_id*1000 + window_id). Produces RAW sensor samples and
observations; IMU/wheel preintegration records are made from them either by the product (device kernel, bench)
or by the oracle (CPU tests) — this module contains no solver arithmetic.

Camera: GF/config/realsense/color.yaml (fx 607.80, fy 607.84, cx 328.80, cy 245.53, 640x480, no distortion);
extrinsic body_T_cam0 and IMU noise: GF/config/realsense/m3dgr.yaml:44-51,113-117; wheel noise :121-123.
"""
import numpy as np
from . import abi

FX, FY, CX, CY = 607.79772949218, 607.83526613281, 328.79772949218, 245.53321838378
W_IMG, H_IMG = 640, 480
BODY_T_CAM0 = np.array([[0.99957087, 0.00215313, 0.02921355, 0.03668114],
                        [-0.00192891, 0.99996848, -0.00770122, -0.00477653],
                        [-0.02922921, 0.00764156, 0.99954353, 0.0316039],
                        [0., 0., 0., 1.]])
ACC_N, GYR_N, ACC_W, GYR_W, G_NORM = 1.2374091609523514e-02, 3.0032654435730201e-03, 1.9218003442176448e-04, 5.4692100664858005e-05, 9.7944
WHEEL_VEL_N, WHEEL_GYR_N = 0.01, 0.004
BASE_SEED = 20250925


def quat_from_R(R):
    """Eigen Quaterniond(Matrix3d) — returns [x y z w]; batched over leading dims."""
    R = np.asarray(R)
    t = np.trace(R, axis1=-2, axis2=-1)
    q = np.zeros(R.shape[:-2] + (4,))
    flat_R = R.reshape(-1, 3, 3); flat_q = q.reshape(-1, 4); flat_t = np.atleast_1d(t).reshape(-1)
    for n in range(flat_R.shape[0]):
        m = flat_R[n]; tt = flat_t[n]
        if tt > 0:
            s = np.sqrt(tt + 1.0); w = 0.5 * s; s = 0.5 / s
            flat_q[n] = [(m[2, 1] - m[1, 2]) * s, (m[0, 2] - m[2, 0]) * s, (m[1, 0] - m[0, 1]) * s, w]
        else:
            i = 0
            if m[1, 1] > m[0, 0]: i = 1
            if m[2, 2] > m[i, i]: i = 2
            j = (i + 1) % 3; k = (j + 1) % 3
            s = np.sqrt(m[i, i] - m[j, j] - m[k, k] + 1.0)
            v = np.zeros(3); v[i] = 0.5 * s; s = 0.5 / s
            w = (m[k, j] - m[j, k]) * s; v[j] = (m[j, i] + m[i, j]) * s; v[k] = (m[k, i] + m[i, k]) * s
            flat_q[n] = [v[0], v[1], v[2], w]
    return q


def R_from_quat(q):
    """[x y z w] -> rotation matrix (Eigen toRotationMatrix), batched."""
    q = np.asarray(q); x, y, z, w = q[..., 0], q[..., 1], q[..., 2], q[..., 3]
    R = np.empty(q.shape[:-1] + (3, 3))
    R[..., 0, 0] = 1 - 2 * (y * y + z * z); R[..., 0, 1] = 2 * (x * y - z * w); R[..., 0, 2] = 2 * (x * z + y * w)
    R[..., 1, 0] = 2 * (x * y + z * w); R[..., 1, 1] = 1 - 2 * (x * x + z * z); R[..., 1, 2] = 2 * (y * z - x * w)
    R[..., 2, 0] = 2 * (x * z - y * w); R[..., 2, 1] = 2 * (y * z + x * w); R[..., 2, 2] = 1 - 2 * (x * x + y * y)
    return R


def so3_exp(v):
    v = np.asarray(v); th = np.linalg.norm(v, axis=-1)[..., None, None]
    K = np.zeros(v.shape[:-1] + (3, 3))
    K[..., 0, 1] = -v[..., 2]; K[..., 0, 2] = v[..., 1]; K[..., 1, 0] = v[..., 2]
    K[..., 1, 2] = -v[..., 0]; K[..., 2, 0] = -v[..., 1]; K[..., 2, 1] = v[..., 0]
    th_safe = np.where(th < 1e-12, 1.0, th)
    a = np.where(th < 1e-12, 1.0, np.sin(th_safe) / th_safe)
    b = np.where(th < 1e-12, 0.5, (1 - np.cos(th_safe)) / th_safe ** 2)
    return np.eye(3) + a * K + b * (K @ K)


def _heading_R(psi):
    """Body frame of a camera-like IMU: x right, y down, z forward; world z up. Columns = body axes in world."""
    c, s = np.cos(psi), np.sin(psi)
    R = np.zeros(np.shape(psi) + (3, 3))
    R[..., 0, 0] = s; R[..., 1, 0] = -c          # right
    R[..., 2, 1] = -1.0                          # down
    R[..., 0, 2] = c; R[..., 1, 2] = s           # forward
    return R


def make_windows(n_windows, config_id=2, n_landmarks=1000, n_frames=11, wheel=False, n_planes=0, ct_fraction=0.0, first_window=0,
                 sorted_landmarks=True, prior="anchor", speed=1.0, yaw_rate=0.2, frame_dt=0.1, imu_hz=200, wheel_hz=50,
                 pixel_noise=0.5, max_landmarks=None, max_obs=None, prior_weight=1.0, prior_stride=None):
    """Return a dict of window-major arrays for `n_windows` windows (W10-F1000 when n_landmarks = 1000).

    Landmark l starts in frame s_l = l mod 8 (track length n_frames - s_l, observed in every later frame). With
    sorted_landmarks the table is ordered by start frame (the reference's f_manager.feature list order is
    non-decreasing in start_frame because features are appended when first seen, feature_manager.cpp:67-88).
    """
    F = n_frames
    L = n_landmarks
    start = (np.arange(L) % 8).astype(np.int32)
    if F < 11:
        start = np.minimum(start, max(F - 4, 0)).astype(np.int32)
    if sorted_landmarks:
        start = np.sort(start, kind="stable")
    tlen = (F - start).astype(np.int32)
    n_obs = int(tlen.sum())
    Lmax = max_landmarks or L
    Omax = max_obs or n_obs
    n_imu = int(round(frame_dt * imu_hz))
    n_whl = int(round(frame_dt * wheel_hz))
    Ric = BODY_T_CAM0[:3, :3]; tic = BODY_T_CAM0[:3, 3]
    qic = quat_from_R(Ric)

    out = {
        "n_frames": F, "max_landmarks": Lmax, "max_obs": Omax, "n_imu_samples": n_imu, "n_wheel_samples": n_whl,
        "para_pose": np.zeros((n_windows, F, 7)), "para_speedbias": np.zeros((n_windows, F, 9)),
        "ex_pose": np.tile(np.concatenate([tic, qic]), (n_windows, 1)), "td": np.zeros(n_windows),
        "n_landmarks": np.full(n_windows, L, np.int32), "inv_depth": np.zeros((n_windows, Lmax)),
        "start_frame": np.zeros((n_windows, Lmax), np.int32), "track_len": np.zeros((n_windows, Lmax), np.int32),
        "fixed": np.zeros((n_windows, Lmax), np.uint8), "obs": np.zeros((n_windows, Omax), abi.OBS),
        "frame_td": np.zeros((n_windows, F)),
        "imu_samples": np.zeros((n_windows, F - 1, n_imu), abi.IMU_SAMPLE), "imu_n": np.full((n_windows, F - 1), n_imu, np.int32),
        "imu_first": np.zeros((n_windows, F - 1, 6)), "imu_lin_bias": np.zeros((n_windows, F - 1, 6)),
        "imu_noise": np.array([ACC_N, GYR_N, ACC_W, GYR_W]),
        "gt_pose": np.zeros((n_windows, F, 7)), "gt_speedbias": np.zeros((n_windows, F, 9)), "gt_inv_depth": np.zeros((n_windows, Lmax)),
        "use_wheel": bool(wheel), "max_planes": int(n_planes),
    }
    out["start_frame"][:, :L] = start; out["track_len"][:, :L] = tlen
    if wheel:
        # wheel frame == body frame rotated so that wheel x is forward: body_T_wheel of m3dgr.yaml:66-76 is specific to
        # that robot; the synthetic robot uses R_io mapping wheel (x fwd, y left, z up) into the body (x right, y down, z fwd)
        Rio = np.array([[0., -1., 0.], [0., 0., -1.], [1., 0., 0.]]); tio = np.array([0.0, 0.05, -0.1])
        out["ex_pose_wheel"] = np.tile(np.concatenate([tio, quat_from_R(Rio)]), (n_windows, 1))
        out["sxsysw"] = np.ones((n_windows, 3)); out["td_wheel"] = np.zeros(n_windows)
        out["wheel_samples"] = np.zeros((n_windows, F - 1, n_whl), abi.WHEEL_SAMPLE)
        out["wheel_n"] = np.full((n_windows, F - 1), n_whl, np.int32)
        out["wheel_first"] = np.zeros((n_windows, F - 1, 6)); out["wheel_lin"] = np.tile(np.array([1., 1., 1., 0.]), (n_windows, F - 1, 1))
        out["wheel_noise"] = np.array([WHEEL_VEL_N, WHEEL_GYR_N])
    if n_planes:
        out["n_planes"] = np.full(n_windows, n_planes, np.int32)
        out["planes"] = np.zeros((n_windows, n_planes), abi.PLANE)
        if ct_fraction > 0:   # CTLidarPlaneNormFactor records (ct == 1): alpha_time per plane
            out["plane_alpha"] = np.zeros((n_windows, n_planes))
    P = prior_stride or abi.MAX_PRIOR_DIM
    out["prior_stride"] = P
    out["prior_rows"] = np.zeros(n_windows, np.int32); out["prior_nblocks"] = np.zeros(n_windows, np.int32)
    out["prior_J0"] = np.zeros((n_windows, P, P)); out["prior_r0"] = np.zeros((n_windows, P))
    out["prior_blocks"] = np.zeros((n_windows, 2 * F + 8), abi.PRIOR_BLOCK)

    obs_begin = np.concatenate([[0], np.cumsum(tlen)[:-1]])
    lm_of_obs = np.repeat(np.arange(L), tlen)
    k_of_obs = np.arange(n_obs) - obs_begin[lm_of_obs]
    for wi in range(n_windows):
        rng = np.random.Generator(np.random.PCG64(BASE_SEED + config_id * 1000 + first_window + wi))
        psi0 = rng.uniform(-np.pi, np.pi)
        p0 = rng.uniform(-5, 5, 3); p0[2] = rng.uniform(0.2, 0.6)
        ba = np.array([0.02, -0.01, 0.015]) * rng.uniform(0.5, 1.5); bg = np.array([0.002, -0.001, 0.0015]) * rng.uniform(0.5, 1.5)
        # --- ground-truth trajectory at IMU rate, frames every n_imu samples; one extra frame before frame 0 for velocities
        dt = 1.0 / imu_hz
        t_imu = np.arange(-n_imu, (F - 1) * n_imu + 1) * dt
        psi = psi0 + yaw_rate * t_imu
        if abs(yaw_rate) > 1e-12:
            px = p0[0] + speed / yaw_rate * (np.sin(psi) - np.sin(psi0)); py = p0[1] - speed / yaw_rate * (np.cos(psi) - np.cos(psi0))
        else:
            px = p0[0] + speed * np.cos(psi0) * t_imu; py = p0[1] + speed * np.sin(psi0) * t_imu
        pos = np.stack([px, py, np.full_like(px, p0[2])], -1)
        vel = np.stack([speed * np.cos(psi), speed * np.sin(psi), np.zeros_like(psi)], -1)
        acc_w = np.stack([-speed * yaw_rate * np.sin(psi), speed * yaw_rate * np.cos(psi), np.zeros_like(psi)], -1)
        Rwb = _heading_R(psi)
        acc_b = np.einsum("nji,nj->ni", Rwb, acc_w + np.array([0, 0, G_NORM])) + ba + rng.normal(0, ACC_N, acc_w.shape)
        gyr_b = np.tile(np.array([0.0, -yaw_rate, 0.0]), (len(psi), 1)) + bg + rng.normal(0, GYR_N, acc_w.shape)
        fidx = n_imu + np.arange(F) * n_imu  # index of frame k in the IMU timeline
        for k in range(F - 1):
            s0 = fidx[k]
            smp = out["imu_samples"][wi, k]
            smp["dt"] = dt; smp["acc"] = acc_b[s0 + 1:s0 + 1 + n_imu]; smp["gyr"] = gyr_b[s0 + 1:s0 + 1 + n_imu]
            out["imu_first"][wi, k, :3] = acc_b[s0]; out["imu_first"][wi, k, 3:] = gyr_b[s0]
        gtR = Rwb[fidx]; gtp = pos[fidx]; gtv = vel[fidx]
        out["gt_pose"][wi, :, :3] = gtp; out["gt_pose"][wi, :, 3:] = quat_from_R(gtR)
        out["gt_speedbias"][wi, :, :3] = gtv; out["gt_speedbias"][wi, :, 3:6] = ba; out["gt_speedbias"][wi, :, 6:] = bg
        # camera poses for frames -1..F-1 (frame -1 only for the velocity of frame-0 observations)
        fall = np.concatenate([[0], fidx])
        Rwc = Rwb[fall] @ Ric; twc = pos[fall] + np.einsum("nij,j->ni", Rwb[fall], tic)
        # --- landmarks: uniform pixel x depth in the host camera frustum, visible in every frame of the track
        Xw = np.zeros((L, 3)); need = np.arange(L)
        for _ in range(200):
            if len(need) == 0: break
            u = rng.uniform(20, W_IMG - 20, len(need)); v = rng.uniform(20, H_IMG - 20, len(need)); d = rng.uniform(3.0, 15.0, len(need))
            pc = np.stack([(u - CX) / FX * d, (v - CY) / FY * d, d], -1)
            s = start[need] + 1
            cand = np.einsum("nij,nj->ni", Rwc[s], pc) + twc[s]
            ok = np.ones(len(need), bool)
            for f in range(1, F + 1):
                pcf = np.einsum("ji,nj->ni", Rwc[f], cand - twc[f])
                uu = FX * pcf[:, 0] / pcf[:, 2] + CX; vv = FY * pcf[:, 1] / pcf[:, 2] + CY
                vis = (pcf[:, 2] > 0.5) & (uu > 2) & (uu < W_IMG - 2) & (vv > 2) & (vv < H_IMG - 2)
                ok &= vis | (f < s)
            Xw[need[ok]] = cand[ok]; need = need[~ok]
        assert len(need) == 0, "landmark sampling failed"
        # --- observations (normalised plane) in every frame incl. the pre-window frame, + pixel noise, fp32
        pcs = np.einsum("fji,flj->fli", Rwc, Xw[None] - twc[:, None])          # [F+1][L][3]
        xn = pcs[..., :2] / pcs[..., 2:3]
        xn = xn + rng.normal(0, pixel_noise, xn.shape) / np.array([FX, FY])
        xn32 = xn.astype(np.float32)
        o = out["obs"][wi]
        fr = start[lm_of_obs] + k_of_obs
        cur = xn32[fr + 1, lm_of_obs]; prv = xn32[fr, lm_of_obs]
        vel32 = np.where((k_of_obs > 0)[:, None], (cur - prv) / np.float32(frame_dt), np.float32(0)).astype(np.float32)
        o["x"][:n_obs] = cur[:, 0]; o["y"][:n_obs] = cur[:, 1]   # new points get velocity 0 (feature_tracker.cpp:840-846)
        o["vx"][:n_obs] = vel32[:, 0]; o["vy"][:n_obs] = vel32[:, 1]
        true_depth = pcs[start + 1, np.arange(L), 2]
        out["gt_inv_depth"][wi, :L] = 1.0 / true_depth
        out["inv_depth"][wi, :L] = (1.0 / true_depth) * (1.0 + rng.normal(0, 0.1, L))
        # --- initial states: ground truth perturbed (5 cm, 1 deg, 0.05 m/s); biases start at 0 = linearisation point
        dR = so3_exp(rng.normal(0, np.deg2rad(1.0), (F, 3)))
        out["para_pose"][wi, :, :3] = gtp + rng.normal(0, 0.05, (F, 3))
        out["para_pose"][wi, :, 3:] = quat_from_R(gtR @ dR)
        out["para_speedbias"][wi, :, :3] = gtv + rng.normal(0, 0.05, (F, 3))
        # --- wheel odometry samples (wheel frame velocity along wheel x, yaw rate about wheel z)
        if wheel:
            n_sub = imu_hz // wheel_hz
            for k in range(F - 1):
                s0 = fidx[k]
                ws = out["wheel_samples"][wi, k]
                ws["dt"] = 1.0 / wheel_hz
                vel_o = np.zeros((n_whl + 1, 3)); vel_o[:, 0] = speed; gyr_o = np.zeros((n_whl + 1, 3)); gyr_o[:, 2] = yaw_rate
                vel_o += rng.normal(0, WHEEL_VEL_N, vel_o.shape); gyr_o += rng.normal(0, WHEEL_GYR_N, gyr_o.shape)
                ws["vel"] = vel_o[1:]; ws["gyr"] = gyr_o[1:]
                out["wheel_first"][wi, k, :3] = vel_o[0]; out["wheel_first"][wi, k, 3:] = gyr_o[0]
        # --- LiDAR plane factors: points on 6 random planes seen from each frame 1..F-1 (config 4)
        if n_planes:
            per = n_planes // (F - 1)
            pl = out["planes"][wi]
            nrm = rng.normal(0, 1, (6, 3)); nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
            off = -np.einsum("ij,ij->i", nrm, gtp.mean(0) + nrm * rng.uniform(3, 10, (6, 1)))  # plane through a point 3..10 m away
            idx = 0
            for f in range(1, F):
                cnt = per if f < F - 1 else n_planes - per * (F - 2)
                which = rng.integers(0, 6, cnt)
                # random world points, projected onto their plane, + 2 cm noise along the normal, into the body frame
                pw = gtp[f] + rng.uniform(-8, 8, (cnt, 3))
                dist = np.einsum("ij,ij->i", nrm[which], pw) + off[which]
                pw = pw - dist[:, None] * nrm[which] + rng.normal(0, 0.02, (cnt, 1)) * nrm[which]
                pb = np.einsum("ji,nj->ni", gtR[f], pw - gtp[f])
                if ct_fraction > 0 and f < F - 1:
                    # continuous-time factors: the point was measured at the pose interpolated between frames f and f + 1
                    from scipy.spatial.transform import Rotation, Slerp
                    is_ct = rng.random(cnt) < ct_fraction
                    alpha = np.where(is_ct, rng.random(cnt), 0.0)
                    sl = Slerp([0.0, 1.0], Rotation.from_matrix(np.stack([gtR[f], gtR[f + 1]])))
                    Ri = sl(alpha).as_matrix(); ti = gtp[f][None] * (1 - alpha[:, None]) + gtp[f + 1][None] * alpha[:, None]
                    pb_ct = np.einsum("nji,nj->ni", Ri, pw - ti)
                    pb = np.where(is_ct[:, None], pb_ct, pb)
                    pl["ct"][idx:idx + cnt] = is_ct.astype(np.int32); out["plane_alpha"][wi, idx:idx + cnt] = alpha
                pl["p_body"][idx:idx + cnt] = pb; pl["normal"][idx:idx + cnt] = nrm[which]; pl["offset"][idx:idx + cnt] = off[which]
                pl["weight"][idx:idx + cnt] = 1.0; pl["frame"][idx:idx + cnt] = f
                idx += cnt
        # --- prior
        if prior == "anchor":  # identity-information anchor on frame 0's pose (6 rows), SURVEY 8(d) config 2
            out["prior_rows"][wi] = 6; out["prior_nblocks"][wi] = 1
            out["prior_J0"][wi, :6, :6] = np.eye(6) * prior_weight
            blk = out["prior_blocks"][wi, 0]
            blk["kind"] = abi.BLK_POSE; blk["index"] = 0; blk["offset"] = 0; blk["x0"][:7] = out["para_pose"][wi, 0]
        elif prior == "dense":  # marginalization-shaped prior over poses 0..F-2 and speed-bias 0 (n = 6(F-1)+9)
            n = 6 * (F - 1) + 9
            A = rng.normal(0, 1, (n, n)) * 0.3 + np.diag(rng.uniform(5, 50, n))
            out["prior_rows"][wi] = n; out["prior_nblocks"][wi] = F
            out["prior_J0"][wi, :n, :n] = np.triu(A)
            out["prior_r0"][wi, :n] = rng.normal(0, 0.05, n)
            off_c = 0
            for b in range(F - 1):
                blk = out["prior_blocks"][wi, b]
                blk["kind"] = abi.BLK_POSE; blk["index"] = b; blk["offset"] = off_c; blk["x0"][:7] = out["para_pose"][wi, b]
                off_c += 6
                if b == 0:
                    blk = out["prior_blocks"][wi, F - 1]
                    blk["kind"] = abi.BLK_SPEEDBIAS; blk["index"] = 0; blk["offset"] = 6 * (F - 1); blk["x0"][:9] = out["para_speedbias"][wi, 0]
            # keep block list ordered: poses first then speed-bias (any order is valid)
    return out


# ---------------------------------------------------------------------------------------------------- front-end images
def image_pair(seed=0, w=640, h=480, shift=(3.3, -2.1), n_pts=300):
    """Band-limited random texture (blurred white noise, contrast normalised) and a smoothly warped second frame + noise;
    corners from a coarse grid of local texture maxima (no cv2 needed)."""
    rng = np.random.default_rng(seed)
    big = rng.normal(size=(h + 64, w + 64))
    k = np.exp(-0.5 * (np.arange(-6, 7) / 2.0) ** 2); k /= k.sum()
    big = np.apply_along_axis(lambda r: np.convolve(r, k, mode="same"), 1, big)
    big = np.apply_along_axis(lambda c: np.convolve(c, k, mode="same"), 0, big)
    big = (big - big.mean()) / big.std()
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float64)

    def sample(dx, dy, rot):
        cx, cy = w / 2, h / 2
        xs = cx + (xx - cx) * np.cos(rot) - (yy - cy) * np.sin(rot) + dx + 32
        ys = cy + (xx - cx) * np.sin(rot) + (yy - cy) * np.cos(rot) + dy + 32
        x0 = np.floor(xs).astype(int); y0 = np.floor(ys).astype(int); ax = xs - x0; ay = ys - y0
        x0 = np.clip(x0, 0, w + 62); y0 = np.clip(y0, 0, h + 62)
        v = big[y0, x0] * (1 - ax) * (1 - ay) + big[y0, x0 + 1] * ax * (1 - ay) + big[y0 + 1, x0] * (1 - ax) * ay + big[y0 + 1, x0 + 1] * ax * ay
        return v
    a = sample(0, 0, 0.0); b = sample(-shift[0], -shift[1], 0.004)
    to8 = lambda v: np.clip(128 + 45 * v + rng.normal(0, 2, v.shape), 0, 255).astype(np.uint8)
    prev, cur = to8(a), to8(b)
    # corners: strongest gradient-energy pixel of each cell of a coarse grid, away from the border
    gy, gx = np.gradient(prev.astype(np.float64))
    e = gx * gx + gy * gy
    pts = []
    step = int(np.sqrt(w * h / n_pts))
    for y in range(12, h - 12 - step, step):
        for x in range(12, w - 12 - step, step):
            c = e[y:y + step, x:x + step]
            iy, ix = np.unravel_index(np.argmax(c), c.shape)
            pts.append((x + ix + 0.25 * ((x + y) % 3), y + iy + 0.125 * ((x * 7 + y) % 5)))
    pts = np.array(pts[:n_pts], np.float32)
    return prev, cur, pts


def _lio_surface(rng, n):
    """Points on a room: ground, four walls, a box top."""
    which = rng.integers(0, 6, n)
    u, v = rng.uniform(-1, 1, n), rng.uniform(-1, 1, n)
    p = np.zeros((n, 3))
    for k in range(n):
        w_ = which[k]
        if w_ == 0: p[k] = (6 * u[k], 5 * v[k], 0.0)                       # ground
        elif w_ == 1: p[k] = (6.0, 5 * u[k], 1.5 + 1.5 * v[k])             # walls
        elif w_ == 2: p[k] = (-6.0, 5 * u[k], 1.5 + 1.5 * v[k])
        elif w_ == 3: p[k] = (6 * u[k], 5.0, 1.5 + 1.5 * v[k])
        elif w_ == 4: p[k] = (6 * u[k], -5.0, 1.5 + 1.5 * v[k])
        else: p[k] = (2.0 + 0.5 * u[k], 1.0 + 0.5 * v[k], 1.0)             # box top
    return p


def lio_scan(seed, n, noise=0.01):
    """n world points of one synthetic scan (surface samples + noise), in scan order."""
    rng = np.random.default_rng(9000 + seed)
    return _lio_surface(rng, n) + rng.normal(0, noise, (n, 3))


def voxel_map_insert(vox, pts, size_voxel_map=0.2, max_points_per_voxel=20, min_distance_points=0.05, min_num_points=0):
    """lidarodom::addPointToMap (LIO/liw/lio/lidarodom.cpp:1167-1213) for the points of a scan in order; vox: dict voxel key -> list of
    points (insertion order). Used to generate map snapshots and as the sequential statement the device insertion is compared with."""
    for p in pts:
        key = tuple(int(c / size_voxel_map) for c in p)    # short(point / voxel_size): truncation toward zero
        blk = vox.get(key)
        if blk is None:
            if min_num_points <= 0:
                vox[key] = [p.copy()]
        elif len(blk) < max_points_per_voxel:
            d2 = min(10 * size_voxel_map * size_voxel_map, min(float(((q - p) ** 2).sum()) for q in blk))
            if d2 > min_distance_points * min_distance_points and (min_num_points <= 0 or len(blk) >= min_num_points):
                blk.append(p.copy())
    return vox


def voxel_map_arrays(vox, max_points_per_voxel=20, order=None):
    """(keys [n, 3] int16, n_points [n], points [n, M, 3]) of a voxel dict, in the given key order (default: dict order)."""
    order = list(vox) if order is None else order
    keys = np.array(order, np.int16).reshape(-1, 3)
    n_points = np.array([len(vox[k]) for k in order], np.int32)
    points = np.zeros((len(order), max_points_per_voxel, 3))
    for j, k in enumerate(order):
        points[j, :n_points[j]] = np.array(vox[k])
    return keys, n_points, points


def lio_scene(seed=0, n_map_points=60000, n_keypoints=3000, size_voxel_map=0.2, max_points_per_voxel=20, min_distance_points=0.05):
    """Synthetic LIO scene: a room (ground, four walls, a box) sampled with 1 cm noise and inserted into a voxel map with the rules
    of lidarodom::addPointToMap (LIO/liw/lio/lidarodom.cpp:1167-1213: at most max_points_per_voxel per voxel, new points at least
    min_distance_points from those already there); keypoints of a new scan on the same surfaces (2 cm noise), some off-surface
    (rejected by the point-to-plane gate) and some in unmapped space (too few neighbours). Returns a dict with the map snapshot
    (keys, n_points, points), the keypoints (abi.LIO_KEYPOINT) and the frame state used to express them."""
    rng = np.random.default_rng(4000 + seed)
    surface = lambda n: _lio_surface(rng, n)
    pts = surface(n_map_points) + rng.normal(0, 0.01, (n_map_points, 3))
    vox = voxel_map_insert({}, pts, size_voxel_map, max_points_per_voxel, min_distance_points, 0)
    order = list(vox)
    perm = rng.permutation(len(order))     # the hash map has no meaningful order: hand the voxels over shuffled
    keys, n_points, points = voxel_map_arrays(vox, max_points_per_voxel, [order[i] for i in perm])
    # frame state: the sensor 1.2 m above the ground, slightly rotated
    yaw = 0.3 + 0.1 * seed
    q = np.array([0.0, 0.0, np.sin(yaw / 2), np.cos(yaw / 2)])
    R = np.array([[np.cos(yaw), -np.sin(yaw), 0], [np.sin(yaw), np.cos(yaw), 0], [0, 0, 1.0]])
    t = np.array([0.5, -0.3, 1.2])
    world = surface(n_keypoints) + rng.normal(0, 0.02, (n_keypoints, 3))
    off = rng.random(n_keypoints) < 0.10
    world[off] += rng.normal(0, 0.35, (int(off.sum()), 3))                     # off-surface points
    far = rng.random(n_keypoints) < 0.03
    world[far] = rng.uniform(-3, 3, (int(far.sum()), 3)) + np.array([0, 0, 12.0])   # unmapped space
    kp = np.zeros(n_keypoints, abi.LIO_KEYPOINT)
    kp["point"] = world
    kp["raw_point"] = (world - t) @ R          # R^T (p - t)
    kp["alpha_time"] = rng.random(n_keypoints)
    return {"keys": keys, "n_points": n_points, "points": points, "keypoints": kp, "rotation": q, "translation": t,
            "translation_begin": t - np.array([0.05, 0.0, 0.0]), "size_voxel_map": size_voxel_map, "max_points_per_voxel": max_points_per_voxel}


def _arc_trajectory(rng, n_frames, speed, yaw_rate, frame_dt, imu_hz, pause, wheel_hz=0):
    """Ground truth of the replay streams: planar arc with an optional stop, IMU samples with the m3dgr noise / bias."""
    Ric = BODY_T_CAM0[:3, :3]; tic = BODY_T_CAM0[:3, 3]
    n_imu = int(round(frame_dt * imu_hz)); dt = 1.0 / imu_hz
    psi0 = 0.3; p0 = np.array([1.0, -2.0, 0.4])
    ba = np.array([0.02, -0.01, 0.015]); bg = np.array([0.002, -0.001, 0.0015])
    t_imu = np.arange(0, (n_frames - 1) * n_imu + 1) * dt
    # speed profile v(t) = speed * g(t) along a circular arc of curvature kappa = yaw_rate / speed (arc length s = integral of v)
    g = np.ones_like(t_imu); gd = np.zeros_like(t_imu)
    if pause is not None:
        ramp = 3 * frame_dt; t1, t2 = pause[0] * frame_dt, pause[1] * frame_dt
        down = (t_imu >= t1 - ramp) & (t_imu < t1); up = (t_imu > t2) & (t_imu <= t2 + ramp)
        g[down] = 0.5 * (1 + np.cos(np.pi * (t_imu[down] - (t1 - ramp)) / ramp)); gd[down] = -0.5 * np.pi / ramp * np.sin(np.pi * (t_imu[down] - (t1 - ramp)) / ramp)
        g[up] = 0.5 * (1 - np.cos(np.pi * (t_imu[up] - t2) / ramp)); gd[up] = 0.5 * np.pi / ramp * np.sin(np.pi * (t_imu[up] - t2) / ramp)
        g[(t_imu >= t1) & (t_imu <= t2)] = 0.0
    kappa = yaw_rate / speed
    v_t = speed * g
    s_arc = np.concatenate([[0.0], np.cumsum(0.5 * (v_t[1:] + v_t[:-1]) * dt)])
    psi = psi0 + kappa * s_arc
    px = p0[0] + (np.sin(psi) - np.sin(psi0)) / kappa; py = p0[1] - (np.cos(psi) - np.cos(psi0)) / kappa
    pos = np.stack([px, py, np.full_like(px, p0[2])], -1)
    tang = np.stack([np.cos(psi), np.sin(psi), np.zeros_like(psi)], -1); nrm_ = np.stack([-np.sin(psi), np.cos(psi), np.zeros_like(psi)], -1)
    vel = v_t[:, None] * tang
    acc_w = (speed * gd)[:, None] * tang + (v_t * v_t * kappa)[:, None] * nrm_
    Rwb = _heading_R(psi)
    acc_b = np.einsum("nji,nj->ni", Rwb, acc_w + np.array([0, 0, G_NORM])) + ba + rng.normal(0, ACC_N, acc_w.shape)
    gyr_b = np.stack([np.zeros_like(psi), -kappa * v_t, np.zeros_like(psi)], -1) + bg + rng.normal(0, GYR_N, acc_w.shape)
    fidx = np.arange(n_frames) * n_imu
    gtR = Rwb[fidx]; gtp = pos[fidx]; gtv = vel[fidx]
    Rwc = gtR @ Ric; twc = gtp + np.einsum("nij,j->ni", gtR, tic)
    imu = []
    for k in range(n_frames - 1):
        s0 = fidx[k]
        smp = np.zeros(n_imu, abi.IMU_SAMPLE); smp["dt"] = dt; smp["acc"] = acc_b[s0 + 1:s0 + 1 + n_imu]; smp["gyr"] = gyr_b[s0 + 1:s0 + 1 + n_imu]
        imu.append({"first": np.concatenate([acc_b[s0], gyr_b[s0]]), "samples": smp})
    out = {"gtp": gtp, "gtR": gtR, "gtv": gtv, "Rwc": Rwc, "twc": twc, "imu": imu, "ba": ba, "bg": bg, "Ric": Ric, "tic": tic}
    if wheel_hz:   # wheel odometer (drawn after everything else so that the IMU / image streams do not depend on it). The odometer frame is
        # ALIGNED with the body frame (R_io = I, a small lever arm): Estimator::processWheel dead-reckons the newest frame with the raw wheel
        # velocity / yaw rate as if they were body quantities (estimator.cpp:871-876, the RIO factor is commented out there)
        n_sub = imu_hz // wheel_hz; n_whl = n_imu // n_sub
        out["Rio"] = np.eye(3); out["tio"] = np.array([0.0, 0.05, -0.1])
        wheel = []
        for k in range(n_frames - 1):
            idx = fidx[k] + n_sub * np.arange(n_whl + 1)
            vel_o = np.zeros((n_whl + 1, 3)); vel_o[:, 2] = v_t[idx]; gyr_o = np.zeros((n_whl + 1, 3)); gyr_o[:, 1] = -kappa * v_t[idx]   # body z forward, yaw about -y
            vel_o += rng.normal(0, WHEEL_VEL_N, vel_o.shape); gyr_o += rng.normal(0, WHEEL_GYR_N, gyr_o.shape)
            smp = np.zeros(n_whl, abi.WHEEL_SAMPLE); smp["dt"] = 1.0 / wheel_hz; smp["vel"] = vel_o[1:]; smp["gyr"] = gyr_o[1:]
            wheel.append({"first": np.concatenate([vel_o[0], gyr_o[0]]), "samples": smp})
        out["wheel"] = wheel
    return out


def image_stream(seed=0, n_frames=300, w=640, h=480, max_flow=4.0):
    """BASELINE.json config 3 stream: one band-limited random texture seen through a smoothly varying similarity warp (translation <= max_flow
    px per frame, slow rotation / zoom) + N(0, 2^2) intensity noise, 30 Hz. Uses cv2.warpAffine when cv2 is importable (1 ms per frame), the
    numpy bilinear sampler otherwise. Returns uint8 [n_frames, h, w]."""
    rng = np.random.default_rng(BASE_SEED + 9000 + seed)
    pad = 256
    big = rng.normal(size=(h + 2 * pad, w + 2 * pad)).astype(np.float32)
    k = np.exp(-0.5 * (np.arange(-6, 7) / 2.0) ** 2).astype(np.float32); k /= k.sum()
    big = np.apply_along_axis(lambda r: np.convolve(r, k, mode="same"), 1, big)
    big = np.apply_along_axis(lambda c: np.convolve(c, k, mode="same"), 0, big)
    big = ((big - big.mean()) / big.std()).astype(np.float32)
    t = np.arange(n_frames)
    dx = pad + 60.0 * np.sin(2 * np.pi * t / 97.0) * (max_flow / 3.9); dy = pad + 45.0 * np.sin(2 * np.pi * t / 71.0 + 0.7) * (max_flow / 4.0)
    rot = 0.05 * np.sin(2 * np.pi * t / 151.0); zoom = 1.0 + 0.03 * np.sin(2 * np.pi * t / 113.0)
    out = np.zeros((n_frames, h, w), np.uint8)
    try:
        import cv2
    except ImportError:
        cv2 = None
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    for f in range(n_frames):
        c, s_ = zoom[f] * np.cos(rot[f]), zoom[f] * np.sin(rot[f])
        M = np.array([[c, -s_, dx[f] + w / 2 - c * w / 2 + s_ * h / 2], [s_, c, dy[f] + h / 2 - s_ * w / 2 - c * h / 2]], np.float32)   # output pixel -> texture coordinate
        if cv2 is not None:
            v = cv2.warpAffine(big, M, (w, h), flags=cv2.INTER_LINEAR | cv2.WARP_INVERSE_MAP)
        else:
            xs = M[0, 0] * xx + M[0, 1] * yy + M[0, 2]; ys = M[1, 0] * xx + M[1, 1] * yy + M[1, 2]
            x0 = np.floor(xs).astype(int); y0 = np.floor(ys).astype(int); ax = xs - x0; ay = ys - y0
            v = big[y0, x0] * (1 - ax) * (1 - ay) + big[y0, x0 + 1] * ax * (1 - ay) + big[y0 + 1, x0] * (1 - ax) * ay + big[y0 + 1, x0 + 1] * ax * ay
        out[f] = np.clip(128 + 45 * v + rng.normal(0, 2, v.shape), 0, 255).astype(np.uint8)
    return out


def lidar_scans(gt_p, gt_R, seed=0, lines=32, az_step_deg=0.6, fov_deg=(-16.0, 15.0), noise=0.01, margin=6.0, height=(0.0, 3.0)):
    """A 32-line spinning LiDAR (BASELINE.json config 5: 32 lines, 10 Hz) at the frame poses of a replay stream, ray-cast against an axis-aligned
    box room that encloses the trajectory with `margin` metres to spare (floor / ceiling at `height`): per frame the hit points in the BODY frame
    (LiDAR frame == body frame, R_IL = I) in scan order (line-major), 1 cm range noise. World z is up; the robot moves in the plane."""
    rng = np.random.Generator(np.random.PCG64(BASE_SEED + 8000 + seed))
    lo = np.array([gt_p[:, 0].min() - margin, gt_p[:, 1].min() - margin, height[0]]); hi = np.array([gt_p[:, 0].max() + margin, gt_p[:, 1].max() + margin, height[1]])
    az = np.deg2rad(np.arange(0.0, 360.0, az_step_deg)); el = np.deg2rad(np.linspace(fov_deg[0], fov_deg[1], lines))
    d = np.stack([np.cos(el)[:, None] * np.cos(az)[None], np.cos(el)[:, None] * np.sin(az)[None], np.sin(el)[:, None] * np.ones_like(az)[None]], -1).reshape(-1, 3)
    scans = []
    for k in range(len(gt_p)):
        o = gt_p[k]
        with np.errstate(divide="ignore", invalid="ignore"):
            t = np.where(d > 0, (hi - o) / d, np.where(d < 0, (lo - o) / d, np.inf)).min(axis=1)
        t = t + rng.normal(0, noise, t.shape)
        pw = o + d * t[:, None]
        scans.append(np.ascontiguousarray((pw - o) @ gt_R[k]))      # R_wb^T (p_w - P)
    return {"scans": scans, "room": (lo, hi)}


def feature_stream(seed=0, n_frames=45, target_features=150, speed=1.0, yaw_rate=0.2, frame_dt=0.1, imu_hz=200, pixel_noise=0.5,
                   max_life=30, depth_range=(1.5, 12.0), pause=None, wheel_hz=0):
    """A synthetic stream for the steady-state replay (BASELINE.json config 5 shape, visual-inertial part): a planar arc like
    make_windows, IMU samples at imu_hz with the m3dgr noise / bias, and per frame the feature map processImage receives:
    id -> [x, y, 1, u, v, vx, vy, depth] (float32-representable x, y, velocities by finite difference, RGB-D depth below 4 m, 0 = invalid).
    pause = (first_frame, last_frame): the robot brakes to a stop and starts again (cosine ramps of 3 frames), which produces the
    low-parallax frames that make the estimator marginalize the second newest frame. Landmarks are born in the current frustum whenever fewer than target_features are visible (what the detector would do), live at
    most max_life frames and die when they leave the image. Returns a dict with frames, imu intervals and ground truth."""
    rng = np.random.Generator(np.random.PCG64(BASE_SEED + 5000 + seed))
    tr = _arc_trajectory(rng, n_frames, speed, yaw_rate, frame_dt, imu_hz, pause)
    gtp, gtR, gtv, Rwc, twc, imu, ba, bg, Ric, tic = (tr[k] for k in ("gtp", "gtR", "gtv", "Rwc", "twc", "imu", "ba", "bg", "Ric", "tic"))
    wheel_part = {}
    if wheel_hz:   # a second generator: the feature / IMU streams stay identical with and without the wheel stream
        trw = _arc_trajectory(np.random.Generator(np.random.PCG64(BASE_SEED + 5000 + seed)), n_frames, speed, yaw_rate, frame_dt, imu_hz, pause, wheel_hz)
        wheel_part = {"wheel": trw["wheel"], "rio": trw["Rio"], "tio": trw["tio"], "wheel_noise": np.array([WHEEL_VEL_N, WHEEL_GYR_N])}
    lm = {}       # id -> (Xw, birth frame)
    prev_xy = {}
    next_id = 0
    frames = []
    for f in range(n_frames):
        def project(X):
            pc = (X - twc[f]) @ Rwc[f]
            return pc, FX * pc[..., 0] / pc[..., 2] + CX, FY * pc[..., 1] / pc[..., 2] + CY
        for i in list(lm):                                  # deaths: out of the image, too close / behind, too old
            pc, u, v = project(lm[i][0])
            if not (pc[2] > 0.5 and 5 < u < W_IMG - 5 and 5 < v < H_IMG - 5) or f - lm[i][1] >= max_life:
                del lm[i]; prev_xy.pop(i, None)
        while len(lm) < target_features:                    # births in the current frustum
            u = rng.uniform(20, W_IMG - 20); v = rng.uniform(20, H_IMG - 20); d = rng.uniform(*depth_range)
            lm[next_id] = (Rwc[f] @ np.array([(u - CX) / FX * d, (v - CY) / FY * d, d]) + twc[f], f); next_id += 1
        ids = np.array(sorted(lm), np.int32)
        pts = np.zeros((len(ids), 8))
        for k, i in enumerate(ids):
            pc, u, v = project(lm[i][0])
            xn = np.float32(pc[0] / pc[2] + rng.normal(0, pixel_noise) / FX); yn = np.float32(pc[1] / pc[2] + rng.normal(0, pixel_noise) / FY)
            vx, vy = (np.float32((xn - prev_xy[i][0]) / np.float32(frame_dt)), np.float32((yn - prev_xy[i][1]) / np.float32(frame_dt))) if i in prev_xy else (0.0, 0.0)
            depth = pc[2] * (1.0 + rng.normal(0, 0.005)) if pc[2] < 4.0 else 0.0
            pts[k] = [xn, yn, 1.0, FX * xn + CX, FY * yn + CY, vx, vy, depth]
            prev_xy[i] = (xn, yn)
        frames.append({"ids": ids, "pts": pts, "header": 100.0 + f * frame_dt})
    return {"frames": frames, "imu": imu, "gt_p": gtp, "gt_R": gtR, "gt_v": gtv, "ba": ba, "bg": bg, "ric": Ric, "tic": tic,
            "imu_noise": np.array([ACC_N, GYR_N, ACC_W, GYR_W]), "n_frames": n_frames, "landmarks_total": next_id, **wheel_part}


def render_stream(seed=0, n_frames=40, speed=1.0, yaw_rate=0.2, frame_dt=0.1, imu_hz=200, pause=None, room=(9.0, 9.0, 3.0), texel=0.015):
    """RGB-D + IMU stream of a camera moving inside a textured box room (BASELINE.json config 5 shape without wheel / LiDAR / GNSS): the
    same arc and IMU model as feature_stream; every frame is ray-cast against the six walls (band-limited random texture, bilinear
    lookup) into a 640x480 uint8 image and a uint16 depth image in millimetres (z-depth, as an RGB-D camera reports it). The scene is
    geometrically consistent, so features tracked in the images are real 3-D points. Returns a dict with images, depths, imu, ground truth."""
    rng = np.random.Generator(np.random.PCG64(BASE_SEED + 7000 + seed))
    tr = _arc_trajectory(rng, n_frames, speed, yaw_rate, frame_dt, imu_hz, pause)
    hx, hy, hz = room                                   # walls at x = +-hx, y = +-hy, floor z = 0... the camera height is 0.4 m: floor z = -0.6, ceiling hz
    lo = np.array([-hx, -hy, -0.6]); hi = np.array([hx, hy, hz])
    n_u = int(2 * max(hx, hy) / texel) + 64; n_v = int(2 * max(hx, hy) / texel) + 64
    tex = rng.normal(size=(n_v, n_u)).astype(np.float32)
    k = np.exp(-0.5 * (np.arange(-7, 8) / 2.5) ** 2).astype(np.float32); k /= k.sum()
    tex = np.apply_along_axis(lambda r: np.convolve(r, k, mode="same"), 1, tex)
    tex = np.apply_along_axis(lambda c: np.convolve(c, k, mode="same"), 0, tex)
    tex = ((tex - tex.mean()) / tex.std()).astype(np.float32)
    vv, uu = np.mgrid[0:H_IMG, 0:W_IMG].astype(np.float64)
    dc = np.stack([(uu - CX) / FX, (vv - CY) / FY, np.ones_like(uu)], -1)          # camera rays, z = 1
    images, depths = [], []
    for f in range(n_frames):
        dw = dc @ tr["Rwc"][f].T                                                    # world rays
        o = tr["twc"][f]
        best_t = np.full(uu.shape, np.inf); best_a = np.zeros(uu.shape, np.int32); best_side = np.zeros(uu.shape, np.int32)
        for axis in range(3):
            for side, plane in enumerate((lo[axis], hi[axis])):
                with np.errstate(divide="ignore", invalid="ignore"):
                    t = (plane - o[axis]) / dw[..., axis]
                ok = (t > 1e-6) & (t < best_t)
                best_t = np.where(ok, t, best_t); best_a = np.where(ok, axis, best_a); best_side = np.where(ok, side, best_side)
        hit = o + dw * best_t[..., None]
        # texture coordinates: the two axes other than the plane's normal, shifted per wall so that walls do not repeat each other
        a1 = np.where(best_a == 0, 1, 0); a2 = np.where(best_a == 2, 1, 2)
        c1 = np.take_along_axis(hit, a1[..., None], -1)[..., 0]; c2 = np.take_along_axis(hit, a2[..., None], -1)[..., 0]
        tu = (c1 + max(hx, hy)) / texel + 17.0 * (2 * best_a + best_side); tv = (c2 + max(hx, hy)) / texel + 11.0 * (2 * best_a + best_side)
        tu = np.clip(tu, 0, n_u - 2.001); tv = np.clip(tv, 0, n_v - 2.001)
        u0 = np.floor(tu).astype(np.int64); v0 = np.floor(tv).astype(np.int64); au = tu - u0; av = tv - v0
        val = tex[v0, u0] * (1 - au) * (1 - av) + tex[v0, u0 + 1] * au * (1 - av) + tex[v0 + 1, u0] * (1 - au) * av + tex[v0 + 1, u0 + 1] * au * av
        img = np.clip(128 + 45 * val + rng.normal(0, 1.5, val.shape), 0, 255).astype(np.uint8)
        images.append(img)
        depths.append(np.clip(np.rint(best_t * 1000.0), 0, 65535).astype(np.uint16))   # rays have z = 1 in the camera frame: t is the z-depth
    return {"images": images, "depths": depths, "imu": tr["imu"], "gt_p": tr["gtp"], "gt_R": tr["gtR"], "gt_v": tr["gtv"], "ba": tr["ba"], "bg": tr["bg"],
            "ric": tr["Ric"], "tic": tr["tic"], "imu_noise": np.array([ACC_N, GYR_N, ACC_W, GYR_W]), "n_frames": n_frames,
            "headers": 100.0 + np.arange(n_frames) * frame_dt, "intrinsics": np.array([FX, FY, CX, CY, 0.0, 0.0, 0.0, 0.0])}
