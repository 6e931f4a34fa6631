"""Steady-state replay (first slice; BASELINE.json config 5 shape, visual-inertial part) through the C++ mirror's
Estimator::processIMU / processImage: keyframe decision, depth initialisation, device solve + marginalization, outlier check, window slide,
for a synthetic feature stream. Every step is checked against the CPU oracle on IDENTICAL inputs (the estimator's capture hook hands over
what it sent to the C ABI and what came back), and the trajectory against ground truth."""
import ctypes as C
import importlib
import time

import numpy as np
import pytest

import host_py as H

pytestmark = pytest.mark.gpu


def _capture(L, e, abi):
    s6 = np.zeros(6, np.int32); L.gf2h_capture_sizes(e, H.p(s6))
    n_lm, n_obs, rows, nb, const_mask, marg_mode = [int(x) for x in s6]
    c = dict(n_lm=n_lm, n_obs=n_obs, rows=rows, nb=nb, const_mask=const_mask, marg_mode=marg_mode,
             pose=np.zeros((11, 7)), sb=np.zeros((11, 9)), ex_td=np.zeros(8), frame_td=np.zeros(11), start=np.zeros(max(n_lm, 1), np.int32), len=np.zeros(max(n_lm, 1), np.int32),
             fixed=np.zeros(max(n_lm, 1), np.uint8), invdep=np.zeros(max(n_lm, 1)), obs=np.zeros(max(n_obs, 1), abi.OBS), imu_samples=np.zeros((10, 256), abi.IMU_SAMPLE),
             imu_n=np.zeros(10, np.int32), imu_first=np.zeros((10, 6)), imu_bias=np.zeros((10, 6)), J0=np.zeros((96, 96)), r0=np.zeros(96), blocks=np.zeros(30, abi.PRIOR_BLOCK),
             pose_out=np.zeros((11, 7)), sb_out=np.zeros((11, 9)), invdep_out=np.zeros(max(n_lm, 1)), pose_marg=np.zeros((11, 7)), sb_marg=np.zeros((11, 9)), invdep_marg=np.zeros(max(n_lm, 1)))
    L.gf2h_capture_get(e, *[H.p(c[k]) for k in ("pose", "sb", "ex_td", "frame_td", "start", "len", "fixed", "invdep", "obs", "imu_samples", "imu_n", "imu_first", "imu_bias",
                                                 "J0", "r0", "blocks", "pose_out", "sb_out", "invdep_out", "pose_marg", "sb_marg", "invdep_marg")])
    return c


def _oracle_window(c, noise, abi):
    n, no = c["n_lm"], c["n_obs"]
    Lm, Om = max(n, 1), max(no, 1)
    w = {"n_frames": 11, "max_landmarks": Lm, "max_obs": Om, "para_pose": c["pose"][None].copy(), "para_speedbias": c["sb"][None].copy(), "ex_pose": c["ex_td"][None, :7].copy(),
         "td": c["ex_td"][7:8].copy(), "inv_depth": c["invdep"][None].copy(), "n_landmarks": np.array([n], np.int32), "start_frame": c["start"][None].copy(),
         "track_len": c["len"][None].copy(), "fixed": c["fixed"][None].copy(), "obs": c["obs"][None].copy(), "frame_td": c["frame_td"][None].copy(),
         "imu_samples": c["imu_samples"][None].copy(), "imu_n": c["imu_n"][None].copy(), "imu_first": c["imu_first"][None].copy(), "imu_lin_bias": c["imu_bias"][None].copy(),
         "imu_noise": noise, "prior_rows": np.array([c["rows"]], np.int32), "prior_J0": c["J0"][None].copy(), "prior_r0": c["r0"][None].copy(),
         "prior_nblocks": np.array([c["nb"]], np.int32), "prior_blocks": c["blocks"][None].copy()}
    return w


def test_steady_state_replay_matches_oracle_step_by_step(gf2, oracle):
    synth = importlib.import_module("gf2_b200.synth")
    abi = gf2.abi
    L = H.lib()
    st = synth.feature_stream(0, n_frames=38, pause=(20, 24))     # the robot stops for a few frames: MARGIN_SECOND_NEW steps
    e = C.c_void_p(L.gf2h_estimator_create())
    L.gf2h_set_extrinsic(e, H.p(st["tic"].copy()), H.p(st["ric"].copy()), C.c_double(0.0), C.c_double(synth.G_NORM), H.p(st["imu_noise"]))
    L.gf2h_set_flags(e, 1, 0, 1, 0)                       # IMU, no wheel, RGB-D depth initialisation, moving-consistency check after the solve
    L.gf2h_set_min_parallax(e, C.c_double(10.0 / 460.0))
    rng = np.random.default_rng(1)
    P = st["gt_p"][:11].copy() + rng.normal(0, 0.01, (11, 3)); R = st["gt_R"][:11].copy(); V = st["gt_v"][:11].copy()
    P[10] = P[9]; R[10] = R[9]; V[10] = V[9]              # what slideWindow leaves in the newest slot (estimator.cpp:3746-3754)
    L.gf2h_set_frame_states(e, H.p(H.frame_states(P, R, V, np.zeros((11, 3)), np.zeros((11, 3)))))
    for f in range(10):
        fr = st["frames"][f]
        L.gf2h_add_image(e, f, len(fr["ids"]), H.p(fr["ids"]), H.p(fr["pts"]), C.c_double(0.0))
    for j in range(1, 10):
        iv = st["imu"][j - 1]
        L.gf2h_new_interval(e, j, H.p(iv["first"][:3].copy()), H.p(iv["first"][3:].copy()), H.p(np.zeros(3)), H.p(np.zeros(3)))
        for s in iv["samples"]:
            L.gf2h_push_imu(e, j, C.c_double(s["dt"]), H.p(s["acc"].copy()), H.p(s["gyr"].copy()))
    L.gf2h_set_imu0(e, H.p(st["imu"][9]["first"][:3].copy()), H.p(st["imu"][9]["first"][3:].copy()))
    pose = np.zeros((11, 7)); sbv = np.zeros((11, 9)); exv = np.zeros(7)
    L.gf2h_vector2double(e, H.p(pose), H.p(sbv), H.p(exv))
    blk = np.zeros(1, abi.PRIOR_BLOCK); blk["kind"] = abi.BLK_POSE; blk["x0"][0, :7] = pose[0]
    L.gf2h_set_prior(e, 6, H.p(np.eye(6) * 100.0), H.p(np.zeros(6)), 1, H.p(blk))      # anchor on the oldest pose until the first marginalization
    L.gf2h_set_capture(e, 1)
    flags, errs, worst = [], [], 0.0
    t_steps = []
    for k in range(10, st["n_frames"]):
        for s in st["imu"][k - 1]["samples"]:
            L.gf2h_process_imu(e, C.c_double(0.0), C.c_double(s["dt"]), H.p(s["acc"].copy()), H.p(s["gyr"].copy()))
        fr = st["frames"][k]
        t0 = time.perf_counter()
        flag = L.gf2h_process_image(e, len(fr["ids"]), H.p(fr["ids"]), H.p(fr["pts"]), C.c_double(fr["header"]))
        t_steps.append(time.perf_counter() - t0)
        assert flag >= 0, L.gf2h_last_error(e)
        flags.append(flag)
        # ---- this step against the oracle on identical inputs
        c = _capture(L, e, abi)
        assert c["n_lm"] > 50 and c["marg_mode"] == flag
        opts = abi.default_opts(const_mask=c["const_mask"])
        w = _oracle_window(c, st["imu_noise"], abi)
        oracle.imu_preintegrate(w)
        oracle.solve_batch(w, opts)
        scale = np.abs(w["para_pose"][0, :, :3]).max()
        d_pos = np.abs(c["pose_out"][:, :3] - w["para_pose"][0, :, :3]).max() / scale
        d_rot = np.abs(c["pose_out"][:, 3:] - w["para_pose"][0, :, 3:]).max()
        worst = max(worst, d_pos, d_rot)
        assert d_pos <= 1e-4 and d_rot <= 1e-4, (k, d_pos, d_rot)                      # BASELINE.json tolerance on pose states
        assert np.abs(c["sb_out"] - w["para_speedbias"][0]).max() <= 1e-4 * max(1.0, np.abs(w["para_speedbias"]).max())
        # marginalization at the states the estimator used
        wm = _oracle_window(c, st["imu_noise"], abi)
        wm["para_pose"][0] = c["pose_marg"]; wm["para_speedbias"][0] = c["sb_marg"]; wm["inv_depth"][0, :c["n_lm"]] = c["invdep_marg"][:c["n_lm"]]
        oracle.imu_preintegrate(wm)
        ref = oracle.marginalize_window(wm, 0, opts, mode=flag)
        nn = C.c_int(0); nb = C.c_int(0); stt = C.c_int(0)
        J0 = np.zeros(96 * 96); r0 = np.zeros(96); blocks = np.zeros(30, abi.PRIOR_BLOCK)
        L.gf2h_get_prior(e, C.byref(nn), H.p(J0), H.p(r0), C.byref(nb), H.p(blocks), C.byref(stt))
        assert stt.value == ref["status"], (k, stt.value, ref["status"])
        if ref["status"] == 0:
            got = {"n": nn.value, "J0": J0[:nn.value ** 2].reshape(nn.value, nn.value), "r0": r0[:nn.value], "blocks": blocks[:nb.value]}
            Hg, gg, _ = oracle.prior_information(got, 11); Hr, gr, _ = oracle.prior_information(ref, 11)
            assert got["n"] == ref["n"] and np.abs(Hg - Hr).max() <= 1e-7 * np.abs(Hr).max(), (k, np.abs(Hg - Hr).max() / np.abs(Hr).max())
        # ---- trajectory against ground truth: the newest frame sits at index 9 after the slide
        out = np.zeros((11, 21)); L.gf2h_get_frame_states(e, H.p(out))
        errs.append(np.linalg.norm(out[9, :3] - st["gt_p"][k]))
    assert 0 in flags and 1 in flags                                                    # both MARGIN_OLD and MARGIN_SECOND_NEW steps happened
    assert max(errs) < 0.15 and np.mean(errs[-10:]) < 0.10, (max(errs), errs[-10:])     # no drift blow-up over the replay (1 m/s, 0.5 px noise)
    print(f"replay: {len(flags)} frames, {flags.count(0)} keyframes, first processImage {t_steps[0] * 1e3:.1f} ms (handle creation), then median {np.median(t_steps[1:]) * 1e3:.2f} ms "
          f"= {1.0 / np.median(t_steps[1:]):.0f} frames/s, max position error {max(errs):.3f} m, worst oracle deviation {worst:.2e}")
    L.gf2h_estimator_destroy(e)


def test_rgbd_imu_stream_end_to_end_images_in_poses_out(gf2, oracle):
    """The whole mirrored pipeline on a geometrically consistent synthetic RGB-D + IMU stream (a camera moving inside a textured room,
    ray-cast per frame): FeatureTracker::trackImage (CLAHE, LK + reverse check, detector on the device; depth lookup) feeds
    Estimator::processImage (device solve + marginalization, window glue). Checked: every solve against the oracle on identical inputs,
    feature bookkeeping sanity, and the estimated trajectory against ground truth."""
    synth = importlib.import_module("gf2_b200.synth")
    abi = gf2.abi
    L = H.lib()
    st = synth.render_stream(0, n_frames=36, pause=(22, 25))
    t = C.c_void_p(L.gf2h_tracker_create(480, 640, 150, 30, H.p(st["intrinsics"])))
    L.gf2h_tracker_set_equalize(t, 1)
    e = C.c_void_p(L.gf2h_estimator_create())
    L.gf2h_set_extrinsic(e, H.p(st["tic"].copy()), H.p(st["ric"].copy()), C.c_double(0.0), C.c_double(synth.G_NORM), H.p(st["imu_noise"]))
    L.gf2h_set_flags(e, 1, 0, 1, 0)
    L.gf2h_set_min_parallax(e, C.c_double(10.0 / 460.0))
    rng = np.random.default_rng(2)
    P = st["gt_p"][:11].copy() + rng.normal(0, 0.01, (11, 3)); R = st["gt_R"][:11].copy(); V = st["gt_v"][:11].copy()
    P[10] = P[9]; R[10] = R[9]; V[10] = V[9]
    L.gf2h_set_frame_states(e, H.p(H.frame_states(P, R, V, np.zeros((11, 3)), np.zeros((11, 3)))))

    def track(k):
        out = np.zeros((200, 10))
        n = L.gf2h_tracker_track(t, C.c_double(st["headers"][k]), H.p(st["images"][k]), H.p(st["depths"][k]), 200, H.p(out))
        assert n > 60, (k, n, L.gf2h_tracker_last_error(t))
        order = np.argsort(out[:n, 0])
        return out[:n, 0][order].astype(np.int32), np.ascontiguousarray(out[:n, 1:9][order]), out[:n, 9][order]

    for k in range(10):
        ids, pts, cnt = track(k)
        L.gf2h_add_image(e, k, len(ids), H.p(ids), H.p(pts), C.c_double(0.0))
    for j in range(1, 10):
        iv = st["imu"][j - 1]
        L.gf2h_new_interval(e, j, H.p(iv["first"][:3].copy()), H.p(iv["first"][3:].copy()), H.p(np.zeros(3)), H.p(np.zeros(3)))
        for s in iv["samples"]:
            L.gf2h_push_imu(e, j, C.c_double(s["dt"]), H.p(s["acc"].copy()), H.p(s["gyr"].copy()))
    L.gf2h_set_imu0(e, H.p(st["imu"][9]["first"][:3].copy()), H.p(st["imu"][9]["first"][3:].copy()))
    pose = np.zeros((11, 7)); sbv = np.zeros((11, 9)); exv = np.zeros(7)
    L.gf2h_vector2double(e, H.p(pose), H.p(sbv), H.p(exv))
    blk = np.zeros(1, abi.PRIOR_BLOCK); blk["kind"] = abi.BLK_POSE; blk["x0"][0, :7] = pose[0]
    L.gf2h_set_prior(e, 6, H.p(np.eye(6) * 100.0), H.p(np.zeros(6)), 1, H.p(blk))
    L.gf2h_set_capture(e, 1)
    flags, errs, n_lm, long_tracks, worst = [], [], [], [], 0.0
    for k in range(10, st["n_frames"]):
        for s in st["imu"][k - 1]["samples"]:
            L.gf2h_process_imu(e, C.c_double(0.0), C.c_double(s["dt"]), H.p(s["acc"].copy()), H.p(s["gyr"].copy()))
        ids, pts, cnt = track(k)
        long_tracks.append(int((cnt >= 4).sum()))
        flag = L.gf2h_process_image(e, len(ids), H.p(ids), H.p(pts), C.c_double(st["headers"][k]))
        assert flag >= 0, L.gf2h_last_error(e)
        flags.append(flag)
        c = _capture(L, e, abi)
        n_lm.append(c["n_lm"])
        opts = abi.default_opts(const_mask=c["const_mask"])
        w = _oracle_window(c, st["imu_noise"], abi)
        oracle.imu_preintegrate(w)
        oracle.solve_batch(w, opts)
        scale = np.abs(w["para_pose"][0, :, :3]).max()
        d = max(np.abs(c["pose_out"][:, :3] - w["para_pose"][0, :, :3]).max() / scale, np.abs(c["pose_out"][:, 3:] - w["para_pose"][0, :, 3:]).max())
        worst = max(worst, d)
        assert d <= 1e-4, (k, d)
        out = np.zeros((11, 21)); L.gf2h_get_frame_states(e, H.p(out))
        errs.append(np.linalg.norm(out[9, :3] - st["gt_p"][k]))
    assert min(n_lm) > 40 and min(long_tracks) > 40                                    # the front end keeps enough long tracks alive
    assert 0 in flags
    assert max(errs) < 0.15, (max(errs), errs)
    print(f"rgbd+imu replay: {len(flags)} frames, {flags.count(0)} keyframes, landmarks in the solve {min(n_lm)}..{max(n_lm)}, max position error {max(errs):.3f} m, "
          f"worst oracle deviation {worst:.2e}")
    L.gf2h_tracker_destroy(t); L.gf2h_estimator_destroy(e)


def test_public_single_thread_api_input_imu_input_image(gf2, oracle):
    """The reference's single-threaded public API (MULTIPLE_THREAD == 0): Estimator::inputIMU + Estimator::inputImage only (estimator.cpp:213-242,
    324-352): inputImage tracks on the device, queues the feature frame and runs processMeasurements (interval extraction with the cut first / last
    dt, processIMU, processImage); after each solve the tracker receives removeOutliers and the constant-velocity prediction (:1185-1189), so the
    hasPrediction branch of trackImage is live. Same stream and checks as the test above."""
    synth = importlib.import_module("gf2_b200.synth")
    abi = gf2.abi
    L = H.lib()
    st = synth.render_stream(0, n_frames=32, pause=(22, 25))
    e = C.c_void_p(L.gf2h_estimator_create())
    L.gf2h_set_extrinsic(e, H.p(st["tic"].copy()), H.p(st["ric"].copy()), C.c_double(0.0), C.c_double(synth.G_NORM), H.p(st["imu_noise"]))
    L.gf2h_set_flags(e, 1, 0, 1, 0)
    L.gf2h_set_min_parallax(e, C.c_double(10.0 / 460.0))
    L.gf2h_set_tracker_parameters(e, 480, 640, 150, 30, 1, H.p(st["intrinsics"]))
    L.gf2h_set_flags(e, 1, 0, 1, 0)
    rng = np.random.default_rng(2)
    P = st["gt_p"][:11].copy() + rng.normal(0, 0.01, (11, 3)); R = st["gt_R"][:11].copy(); V = st["gt_v"][:11].copy()
    P[10] = P[9]; R[10] = R[9]; V[10] = V[9]
    L.gf2h_set_frame_states(e, H.p(H.frame_states(P, R, V, np.zeros((11, 3)), np.zeros((11, 3)))))
    for k in range(10):                                    # the first window is filled outside the steady-state path (initialisation is out of scope)
        out = np.zeros((200, 10))
        n = L.gf2h_estimator_track_only(e, C.c_double(st["headers"][k]), H.p(st["images"][k]), H.p(st["depths"][k]), 200, H.p(out))
        assert n > 60
        order = np.argsort(out[:n, 0]); ids = out[:n, 0][order].astype(np.int32); pts = np.ascontiguousarray(out[:n, 1:9][order])
        L.gf2h_add_image(e, k, len(ids), H.p(ids), H.p(pts), C.c_double(0.0))
    for j in range(1, 10):
        iv = st["imu"][j - 1]
        L.gf2h_new_interval(e, j, H.p(iv["first"][:3].copy()), H.p(iv["first"][3:].copy()), H.p(np.zeros(3)), H.p(np.zeros(3)))
        for s in iv["samples"]:
            L.gf2h_push_imu(e, j, C.c_double(s["dt"]), H.p(s["acc"].copy()), H.p(s["gyr"].copy()))
    L.gf2h_set_imu0(e, H.p(st["imu"][9]["first"][:3].copy()), H.p(st["imu"][9]["first"][3:].copy()))
    L.gf2h_set_prev_time(e, C.c_double(st["headers"][9]), C.c_double(st["headers"][9]))
    pose = np.zeros((11, 7)); sbv = np.zeros((11, 9)); exv = np.zeros(7)
    L.gf2h_vector2double(e, H.p(pose), H.p(sbv), H.p(exv))
    blk = np.zeros(1, abi.PRIOR_BLOCK); blk["kind"] = abi.BLK_POSE; blk["x0"][0, :7] = pose[0]
    L.gf2h_set_prior(e, 6, H.p(np.eye(6) * 100.0), H.p(np.zeros(6)), 1, H.p(blk))
    L.gf2h_set_capture(e, 1)
    flags, errs, worst = [], [], 0.0
    for k in range(10, st["n_frames"]):
        t_prev = st["headers"][k - 1]
        for i, s in enumerate(st["imu"][k - 1]["samples"]):
            L.gf2h_input_imu(e, C.c_double(t_prev + (i + 1) * float(s["dt"])), H.p(s["acc"].copy()), H.p(s["gyr"].copy()))
        flag = L.gf2h_input_image(e, C.c_double(st["headers"][k]), H.p(st["images"][k]), H.p(st["depths"][k]))
        assert flag >= 0, L.gf2h_last_error(e)
        flags.append(flag)
        c = _capture(L, e, abi)
        assert c["imu_n"][9] == 20 and abs(c["imu_samples"][9]["dt"][:20].sum() - 0.1) < 1e-9     # the newest interval: exactly the samples between the two images
        opts = abi.default_opts(const_mask=c["const_mask"])
        w = _oracle_window(c, st["imu_noise"], abi)
        oracle.imu_preintegrate(w)
        oracle.solve_batch(w, opts)
        scale = np.abs(w["para_pose"][0, :, :3]).max()
        d = max(np.abs(c["pose_out"][:, :3] - w["para_pose"][0, :, :3]).max() / scale, np.abs(c["pose_out"][:, 3:] - w["para_pose"][0, :, 3:]).max())
        worst = max(worst, d)
        assert d <= 1e-4, (k, d)
        out = np.zeros((11, 21)); L.gf2h_get_frame_states(e, H.p(out))
        errs.append(np.linalg.norm(out[9, :3] - st["gt_p"][k]))
    q = np.zeros(3, np.int32); L.gf2h_queue_sizes(e, H.p(q))
    assert q[2] == 0 and q[0] == 1                          # every image consumed; the IMU sample at the last image time stays queued
    assert 0 in flags and max(errs) < 0.25, (flags, max(errs))     # observed 0.12 m: with the prediction the tracks live longer, fewer keyframes
    print(f"inputIMU/inputImage replay: {len(flags)} frames, {flags.count(0)} keyframes, max position error {max(errs):.3f} m, worst oracle deviation {worst:.2e}")
    L.gf2h_estimator_destroy(e)


def test_replay_with_wheel_odometer_and_free_wheel_extrinsic(gf2, oracle):
    """wheel: 1, estimate_wheel_extrinsic: 1 (gc_test / groundchallenge / idc_rs / m2dgrp .yaml) in the loop: processWheel buffers the 50 Hz odometer
    and dead-reckons the newest frame, the solve frees body_T_wheel (openExWheelEstimation), the marginalization keeps the wheel calibration blocks;
    every step against the oracle on identical inputs (incl. the wheel samples), trajectory against ground truth."""
    synth = importlib.import_module("gf2_b200.synth")
    abi = gf2.abi
    L = H.lib()
    st = synth.feature_stream(0, n_frames=30, pause=(20, 23), wheel_hz=50)
    e = C.c_void_p(L.gf2h_estimator_create())
    L.gf2h_set_extrinsic(e, H.p(st["tic"].copy()), H.p(st["ric"].copy()), C.c_double(0.0), C.c_double(synth.G_NORM), H.p(st["imu_noise"]))
    calib = np.concatenate([st["tio"], st["rio"].ravel(), [1.0, 1.0, 1.0, 0.0]])
    L.gf2h_set_wheel_parameters(e, H.p(calib), H.p(np.array([1.0, 1.0, 0.0, 0.0, st["wheel_noise"][0], st["wheel_noise"][1]])))
    L.gf2h_set_flags(e, 1, 1, 1, 0)
    L.gf2h_set_min_parallax(e, C.c_double(10.0 / 460.0))
    rng = np.random.default_rng(1)
    P = st["gt_p"][:11].copy() + rng.normal(0, 0.01, (11, 3)); R = st["gt_R"][:11].copy(); V = st["gt_v"][:11].copy()
    P[10] = P[9]; R[10] = R[9]; V[10] = V[9]
    L.gf2h_set_frame_states(e, H.p(H.frame_states(P, R, V, np.zeros((11, 3)), np.zeros((11, 3)))))
    for f in range(10):
        fr = st["frames"][f]
        L.gf2h_add_image(e, f, len(fr["ids"]), H.p(fr["ids"]), H.p(fr["pts"]), C.c_double(0.0))
    for j in range(1, 10):
        iv = st["imu"][j - 1]
        L.gf2h_new_interval(e, j, H.p(iv["first"][:3].copy()), H.p(iv["first"][3:].copy()), H.p(np.zeros(3)), H.p(np.zeros(3)))
        for s in iv["samples"]:
            L.gf2h_push_imu(e, j, C.c_double(s["dt"]), H.p(s["acc"].copy()), H.p(s["gyr"].copy()))
        wv = st["wheel"][j - 1]
        L.gf2h_new_wheel_interval(e, j, H.p(wv["first"][:3].copy()), H.p(wv["first"][3:].copy()))
        for s in wv["samples"]:
            L.gf2h_push_wheel(e, j, C.c_double(s["dt"]), H.p(s["vel"].copy()), H.p(s["gyr"].copy()))
    L.gf2h_set_imu0(e, H.p(st["imu"][9]["first"][:3].copy()), H.p(st["imu"][9]["first"][3:].copy()))
    w0 = st["wheel"][9]["first"]
    L.gf2h_process_wheel(e, C.c_double(0.0), C.c_double(0.0), H.p(w0[:3].copy()), H.p(w0[3:].copy()))   # the sample at the image time: sets vel_0 / gyr_0 of the next interval, dt = 0
    pose = np.zeros((11, 7)); sbv = np.zeros((11, 9)); exv = np.zeros(7)
    L.gf2h_vector2double(e, H.p(pose), H.p(sbv), H.p(exv))
    blk = np.zeros(1, abi.PRIOR_BLOCK); blk["kind"] = abi.BLK_POSE; blk["x0"][0, :7] = pose[0]
    L.gf2h_set_prior(e, 6, H.p(np.eye(6) * 100.0), H.p(np.zeros(6)), 1, H.p(blk))
    L.gf2h_set_capture(e, 1)
    flags, errs, worst, worst_cal = [], [], 0.0, 0.0
    for k in range(10, st["n_frames"]):
        for s in st["imu"][k - 1]["samples"]:
            L.gf2h_process_imu(e, C.c_double(0.0), C.c_double(s["dt"]), H.p(s["acc"].copy()), H.p(s["gyr"].copy()))
        for s in st["wheel"][k - 1]["samples"]:
            L.gf2h_process_wheel(e, C.c_double(0.0), C.c_double(s["dt"]), H.p(s["vel"].copy()), H.p(s["gyr"].copy()))
        fr = st["frames"][k]
        flag = L.gf2h_process_image(e, len(fr["ids"]), H.p(fr["ids"]), H.p(fr["pts"]), C.c_double(fr["header"]))
        assert flag >= 0, L.gf2h_last_error(e)
        flags.append(flag)
        c = _capture(L, e, abi)
        ws = np.zeros((10, 64), abi.WHEEL_SAMPLE); wn = np.zeros(10, np.int32); wf = np.zeros((10, 6)); wl = np.zeros((10, 4)); cal = np.zeros(12); exw_out = np.zeros(7)
        assert L.gf2h_capture_get_wheel(e, H.p(ws), H.p(wn), H.p(wf), H.p(wl), H.p(cal), H.p(exw_out)) == 1
        assert not (c["const_mask"] & abi.CONST_EX_WHEEL) and (c["const_mask"] & abi.CONST_WHEEL_INTRINSIC)     # extrinsic free (|Vs[0]| > 0.2 latched), intrinsics fixed
        opts = abi.default_opts(const_mask=c["const_mask"])
        w = _oracle_window(c, st["imu_noise"], abi)
        w.update(use_wheel=True, ex_pose_wheel=cal[None, :7].copy(), sxsysw=cal[None, 7:10].copy(), td_wheel=cal[10:11].copy(), wheel_samples=ws[None].copy(), wheel_n=wn[None].copy(),
                 wheel_first=wf[None].copy(), wheel_lin=wl[None].copy(), wheel_noise=st["wheel_noise"])
        oracle.imu_preintegrate(w); oracle.wheel_preintegrate(w)
        start_cal = w["ex_pose_wheel"].copy()
        oracle.solve_batch(w, opts)
        scale = np.abs(w["para_pose"][0, :, :3]).max()
        d = max(np.abs(c["pose_out"][:, :3] - w["para_pose"][0, :, :3]).max() / scale, np.abs(c["pose_out"][:, 3:] - w["para_pose"][0, :, 3:]).max())
        worst = max(worst, d)
        assert d <= 1e-4, (k, d)
        moved = np.abs(w["ex_pose_wheel"] - start_cal).max()
        dc = np.abs(exw_out - w["ex_pose_wheel"][0]).max()
        worst_cal = max(worst_cal, dc / max(moved, 1e-12))
        assert dc <= 1e-3 * moved + 1e-9, (k, dc, moved)
        out = np.zeros((11, 21)); L.gf2h_get_frame_states(e, H.p(out))
        errs.append(np.linalg.norm(out[9, :3] - st["gt_p"][k]))
    cal16 = np.zeros(16); of = np.zeros(2, np.int32); L.gf2h_get_wheel_states(e, H.p(cal16), H.p(of))
    assert of.tolist() == [1, 0] and 0 in flags and 1 in flags
    assert max(errs) < 0.15, (max(errs), errs)
    nn = C.c_int(0); nb = C.c_int(0); stt = C.c_int(0)
    J0 = np.zeros(96 * 96); r0 = np.zeros(96); blocks = np.zeros(30, abi.PRIOR_BLOCK)
    L.gf2h_get_prior(e, C.byref(nn), H.p(J0), H.p(r0), C.byref(nb), H.p(blocks), C.byref(stt))
    assert abi.BLK_EX_WHEEL in set(int(b["kind"]) for b in blocks[:nb.value])                                 # the wheel calibration lives on in the prior
    print(f"wheel replay: {len(flags)} frames, {flags.count(0)} keyframes, max position error {max(errs):.3f} m, worst oracle deviation {worst:.2e}, "
          f"calibration deviation / motion {worst_cal:.2e}, |tio - truth| = {np.abs(cal16[:3] - st['tio']).max():.3f} m")
    L.gf2h_estimator_destroy(e)


def test_full_fusion_replay_lidar_planes_wheel_features(gf2, oracle):
    """BASELINE.json config 5 shape: features + IMU + 50 Hz wheel odometer + a 32-line 10 Hz LiDAR. Per frame the scan's keypoints go through the
    device's addSurfCostFactor (gf2_lio_build_factors) against the device-resident voxel map at the dead-reckoned pose, the point-to-plane factors
    ride in the window on that frame's pose (Estimator::inputLidarPlanes -> gf2_set_planes, slid with the frames), processImage solves and
    marginalizes, the scan is inserted into the map at the solved pose. Every solve against the oracle on identical inputs (planes included)."""
    synth = importlib.import_module("gf2_b200.synth")
    abi = gf2.abi
    L = H.lib()
    st = synth.feature_stream(3, n_frames=26, pause=(18, 20), wheel_hz=50)
    lid = synth.lidar_scans(st["gt_p"], st["gt_R"], seed=3)
    e = C.c_void_p(L.gf2h_estimator_create())
    L.gf2h_set_extrinsic(e, H.p(st["tic"].copy()), H.p(st["ric"].copy()), C.c_double(0.0), C.c_double(synth.G_NORM), H.p(st["imu_noise"]))
    calib = np.concatenate([st["tio"], st["rio"].ravel(), [1.0, 1.0, 1.0, 0.0]])
    L.gf2h_set_wheel_parameters(e, H.p(calib), H.p(np.array([1.0, 0.0, 0.0, 0.0, st["wheel_noise"][0], st["wheel_noise"][1]])))
    L.gf2h_set_flags(e, 1, 1, 1, 0)
    L.gf2h_set_min_parallax(e, C.c_double(10.0 / 460.0))
    rng = np.random.default_rng(4)
    P = st["gt_p"][:11].copy() + rng.normal(0, 0.01, (11, 3)); R = st["gt_R"][:11].copy(); V = st["gt_v"][:11].copy()
    P[10] = P[9]; R[10] = R[9]; V[10] = V[9]
    L.gf2h_set_frame_states(e, H.p(H.frame_states(P, R, V, np.zeros((11, 3)), np.zeros((11, 3)))))
    lio = gf2.Lio(max_voxels=200000, max_keypoints=4096)
    for f in range(10):
        fr = st["frames"][f]
        L.gf2h_add_image(e, f, len(fr["ids"]), H.p(fr["ids"]), H.p(fr["pts"]), C.c_double(0.0))
        lio.add_points(lid["scans"][f] @ st["gt_R"][f].T + st["gt_p"][f])
    for j in range(1, 10):
        iv = st["imu"][j - 1]
        L.gf2h_new_interval(e, j, H.p(iv["first"][:3].copy()), H.p(iv["first"][3:].copy()), H.p(np.zeros(3)), H.p(np.zeros(3)))
        for s in iv["samples"]:
            L.gf2h_push_imu(e, j, C.c_double(s["dt"]), H.p(s["acc"].copy()), H.p(s["gyr"].copy()))
        wv = st["wheel"][j - 1]
        L.gf2h_new_wheel_interval(e, j, H.p(wv["first"][:3].copy()), H.p(wv["first"][3:].copy()))
        for s in wv["samples"]:
            L.gf2h_push_wheel(e, j, C.c_double(s["dt"]), H.p(s["vel"].copy()), H.p(s["gyr"].copy()))
    L.gf2h_set_imu0(e, H.p(st["imu"][9]["first"][:3].copy()), H.p(st["imu"][9]["first"][3:].copy()))
    w0 = st["wheel"][9]["first"]
    L.gf2h_process_wheel(e, C.c_double(0.0), C.c_double(0.0), H.p(w0[:3].copy()), H.p(w0[3:].copy()))
    pose = np.zeros((11, 7)); sbv = np.zeros((11, 9)); exv = np.zeros(7)
    L.gf2h_vector2double(e, H.p(pose), H.p(sbv), H.p(exv))
    blk = np.zeros(1, abi.PRIOR_BLOCK); blk["kind"] = abi.BLK_POSE; blk["x0"][0, :7] = pose[0]
    L.gf2h_set_prior(e, 6, H.p(np.eye(6) * 100.0), H.p(np.zeros(6)), 1, H.p(blk))
    L.gf2h_set_capture(e, 1)
    L.gf2h_capture_get_planes.restype = C.c_int
    flags, errs, worst, n_planes_seen = [], [], 0.0, []
    for k in range(10, st["n_frames"]):
        for s in st["imu"][k - 1]["samples"]:
            L.gf2h_process_imu(e, C.c_double(0.0), C.c_double(s["dt"]), H.p(s["acc"].copy()), H.p(s["gyr"].copy()))
        for s in st["wheel"][k - 1]["samples"]:
            L.gf2h_process_wheel(e, C.c_double(0.0), C.c_double(s["dt"]), H.p(s["vel"].copy()), H.p(s["gyr"].copy()))
        out = np.zeros((11, 21)); L.gf2h_get_frame_states(e, H.p(out))
        Rp = out[10, 3:12].reshape(3, 3); Pn = out[10, :3]
        assert np.linalg.norm(Pn - st["gt_p"][k]) < 0.2                                 # processWheel dead-reckoned the newest frame
        scan = lid["scans"][k]
        kp = np.zeros(len(scan[::6]), abi.LIO_KEYPOINT); kp["raw_point"] = scan[::6]; kp["point"] = scan[::6] @ Rp.T + Pn
        o = abi.default_lio_opts(icp_model=abi.ICP_POINT_TO_PLANE, max_num_residuals=480, rotation=synth.quat_from_R(Rp), translation=Pn, translation_begin=Pn)
        fac, _, _, _ = lio.build_factors(kp, o)
        assert len(fac) > 200, len(fac)
        L.gf2h_input_lidar_planes(e, len(fac), H.p(fac), C.c_double(31.622776601683793))
        fr = st["frames"][k]
        flag = L.gf2h_process_image(e, len(fr["ids"]), H.p(fr["ids"]), H.p(fr["pts"]), C.c_double(fr["header"]))
        assert flag >= 0, L.gf2h_last_error(e)
        flags.append(flag)
        # ---- this solve against the oracle on identical inputs: landmarks, IMU + wheel samples, prior, planes
        c = _capture(L, e, abi)
        ws = np.zeros((10, 64), abi.WHEEL_SAMPLE); wn = np.zeros(10, np.int32); wf = np.zeros((10, 6)); wl = np.zeros((10, 4)); cal = np.zeros(12); exw_out = np.zeros(7)
        assert L.gf2h_capture_get_wheel(e, H.p(ws), H.p(wn), H.p(wf), H.p(wl), H.p(cal), H.p(exw_out)) == 1
        pl = np.zeros(11 * 480, abi.PLANE); npl = L.gf2h_capture_get_planes(e, H.p(pl), len(pl))
        n_planes_seen.append(npl)
        assert npl >= len(fac) and set(np.unique(pl["frame"][:npl])) <= set(range(11)) and (pl["frame"][:npl] == 10).sum() == len(fac)
        opts = abi.default_opts(const_mask=c["const_mask"], lidar_sqrt_info=31.622776601683793)
        w = _oracle_window(c, st["imu_noise"], abi)
        w.update(use_wheel=True, ex_pose_wheel=cal[None, :7].copy(), sxsysw=cal[None, 7:10].copy(), td_wheel=cal[10:11].copy(), wheel_samples=ws[None].copy(), wheel_n=wn[None].copy(),
                 wheel_first=wf[None].copy(), wheel_lin=wl[None].copy(), wheel_noise=st["wheel_noise"],
                 max_planes=max(npl, 1), n_planes=np.array([npl], np.int32), planes=pl[None, :max(npl, 1)].copy())
        oracle.imu_preintegrate(w); oracle.wheel_preintegrate(w)
        oracle.solve_batch(w, opts)
        scale = np.abs(w["para_pose"][0, :, :3]).max()
        d = max(np.abs(c["pose_out"][:, :3] - w["para_pose"][0, :, :3]).max() / scale, np.abs(c["pose_out"][:, 3:] - w["para_pose"][0, :, 3:]).max())
        worst = max(worst, d)
        assert d <= 1e-4, (k, d)
        L.gf2h_get_frame_states(e, H.p(out))
        lio.add_points(scan @ out[9, 3:12].reshape(3, 3).T + out[9, :3])
        errs.append(np.linalg.norm(out[9, :3] - st["gt_p"][k]))
    assert 0 in flags and 1 in flags
    assert max(n_planes_seen) > 3000                                                    # the window fills up with the scans' factors as it slides
    assert max(errs) < 0.10, (max(errs), errs)
    print(f"full-fusion replay: {len(flags)} frames, {flags.count(0)} keyframes, up to {max(n_planes_seen)} plane factors per window, "
          f"max position error {max(errs):.3f} m, worst oracle deviation {worst:.2e}")
    L.gf2h_estimator_destroy(e); lio.close()
