#!/usr/bin/env python
"""bench.py — sliding-window solves/sec on the W10-F1000 window (BASELINE.json config 2).

One "step" = one pass of the hot path over one batch of B synthetic windows per GPU: device IMU preintegration
(200 Hz, 20 samples/interval) + a full 8-iteration trust-region solve (Jacobian sweep, Schur, reduced solve,
back-substitution, candidate evaluation) of every window. `value` is measured with the inputs resident in HBM; `e2e`
goes through the public C ABI with pinned HOST buffers (H2D of every input, D2H of states + inverse depths) each step.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--windows B] [--impl gf2|reference]

Under torchrun (N > 1) every rank solves its own B windows (window-level data parallelism, no collective: weak scaling).
`--impl reference` times the restated-reference CPU path (oracle/, Ceres is not installable here) on the host cores.
"""
import argparse
import ctypes
import importlib
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "oracle")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

# algorithmic bytes of one Jacobian sweep of one W10-F1000 window
N_OBS, N_LM, N_FRAMES, N_IMU, D_RED = 6500, 1000, 11, 10, 165
PRIOR_STRIDE = 8   # row stride of the prior arrays (the anchor prior of this workload has 6 rows)
WORKLOAD = "W10-F1000 (11 frames, 1000 landmarks, 6500 projection + 10 IMU factors + anchor prior), 200 Hz IMU preintegration, 8 trust-region iterations, time cap off"
BYTES_SWEEP = N_OBS * 20 + N_LM * 32 + N_FRAMES * 136 + 64 + N_IMU * 1456 + (D_RED * (D_RED + 1) // 2 + D_RED) * 8  # = 289,000


FP64_SM_CYCLES_PER_WINDOW = (39000 * 2 + 9500 * 16) / 4   # k_linearize: fp64 datapath time of one W10-F1000 window


def fp64_bound_ms(B, clocks, n_sm=148):
    """k_linearize's launch time if the SMs' fp64 datapath never idled: windows / SMs x cycles per window / SM clock."""
    mhz = (clocks or {}).get("sm_mhz") or 1965.0
    return 1e3 * B / n_sm * FP64_SM_CYCLES_PER_WINDOW / (mhz * 1e6)


def ncu_summary():
    """Per-kernel numbers of the latest `ncu --set full` capture of this build, written by scripts/ncu_summarize.py into profiles/ (dated, with
    the capture's command): DRAM bytes per launch, fp64 tensor sub-pipe activity, warps active. None when no summary is committed."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_summary_r2.json")) as f:
            return json.load(f)
    except Exception:
        return None


def bench_config(B, distinct, world, h2d_mb=None):
    """The `config` object of the JSON line: identical for the gf2 arm and the reference arm (same workload, same batch definition)."""
    return {"workload": WORKLOAD, "windows_per_gpu_per_step": B, "distinct_windows": distinct, "parallelism": f"window-dp{world} (no collective)",
            "l2": "inputs larger than L2 (631 MB of window data per GPU at 4096 windows)"}


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(gpu_index)],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush(); self.f.seek(0)
        sm, mx, reasons = [], [], set()
        for line in self.f.read().strip().splitlines():
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1])); mx.append(float(parts[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), parts[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if sm:
            out = {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        return out


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def make_batch(gf2, synth, B, distinct, first_window=0, pinned=True, prior_stride=PRIOR_STRIDE):
    """B windows tiled from `distinct` generated ones, in pinned host arrays."""
    base = synth.make_windows(distinct, n_landmarks=N_LM, first_window=first_window, prior_stride=prior_stride)
    reps = (B + distinct - 1) // distinct
    w = {}
    for k, v in base.items():
        if isinstance(v, np.ndarray) and v.ndim >= 1 and v.shape[0] == distinct and k != "imu_noise":
            t = np.concatenate([v] * reps)[:B]
            if pinned:
                a = gf2.pinned_empty(t.shape, t.dtype); a[...] = t; t = a
            w[k] = t
        else:
            w[k] = v
    return w


def run_lk(gf2, synth, streams=64, steps=10, cv2_seconds=2.0):
    """FeatureTracker::trackImage's LK stage (forward 4 levels + backward 2 levels + consistency check) on `streams`
    independent 640x480 frame pairs per launch; images cross PCIe every call (the tracker API takes host images)."""
    base = [synth.image_pair(s % 8, shift=(2.0 + 0.3 * (s % 8), -1.0), n_pts=400) for s in range(min(streams, 8))]
    npts = min(300, min(len(b[2]) for b in base))      # BASELINE config 3: 300 corners
    prev = gf2.pinned_empty((streams, 480, 640), np.uint8); cur = gf2.pinned_empty((streams, 480, 640), np.uint8)   # host images in pinned memory: the upload runs at PCIe speed
    for s in range(streams):
        prev[s] = base[s % len(base)][0]; cur[s] = base[s % len(base)][1]
    pts = np.stack([base[s % len(base)][2][:npts] for s in range(streams)])
    t = gf2.Tracker(640, 480, max_pts=npts, max_streams=streams)
    for _ in range(3):
        t.track_fb(prev, cur, pts)
    lk_ms = tot_ms = 0.0
    t0 = time.perf_counter()
    for _ in range(steps):
        out, ok = t.track_fb(prev, cur, pts)
        tm = t.last_timing(); lk_ms += tm["lk_ms"]; tot_ms += tm["total_ms"]
    wall = time.perf_counter() - t0
    t1 = gf2.Tracker(640, 480, max_pts=npts, max_streams=1)
    for _ in range(3):
        t1.track_fb(prev[:1], cur[:1], pts[:1])
    one = []
    for _ in range(20):
        t1.track_fb(prev[:1], cur[:1], pts[:1]); one.append(t1.last_timing()["total_ms"])
    # algorithmic bytes of one frame pair (image + its pyramid once, patch + search window per level pass)
    bytes_lk = 640 * 480 * (1 + 2 * (0.25 + 0.0625 + 0.015625)) + npts * 6 * ((21 + 2) ** 2 + (21 + 8) ** 2)
    peaks, _ = measured_peaks()
    ach = bytes_lk * streams / (lk_ms / steps / 1e3) / 1e9
    line = {"metric": "LK frame pairs/sec (640x480, 4 levels, fwd+bwd)", "value": streams * steps / wall, "unit": "frame pairs/s", "streams": streams, "points": int(npts),
            "device_ms_per_batch": tot_ms / steps, "lk_kernel_ms_per_batch": lk_ms / steps, "single_stream_ms_per_pair": float(np.median(one)), "tracked_fraction": float(ok.mean()),
            "roofline": {"bound": "hbm", "kernel": "k_lk", "achieved": ach, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": ach / peaks["hbm_gbs"], "algorithmic_bytes_per_pair": bytes_lk}}
    if cv2_seconds > 0:
        try:
            import cv2
            crit = (cv2.TERM_CRITERIA_COUNT + cv2.TERM_CRITERIA_EPS, 30, 0.01)
            cv2.setNumThreads(0)
            t0 = time.perf_counter(); n = 0
            while time.perf_counter() - t0 < cv2_seconds:
                s = n % len(base)
                c, st, _ = cv2.calcOpticalFlowPyrLK(base[s][0], base[s][1], base[s][2][:npts].reshape(-1, 1, 2), None, winSize=(21, 21), maxLevel=3, criteria=crit)
                cv2.calcOpticalFlowPyrLK(base[s][1], base[s][0], c, base[s][2][:npts].reshape(-1, 1, 2).copy(), winSize=(21, 21), maxLevel=1, criteria=crit, flags=cv2.OPTFLOW_USE_INITIAL_FLOW)
                n += 1
            line["cpu_baseline"] = {"value": n / (time.perf_counter() - t0), "unit": "frame pairs/s", "kind": "reference", "sample": f"cv2 {cv2.__version__} calcOpticalFlowPyrLK x2, all threads, one stream, {cv2_seconds:.0f} s"}
        except ImportError:
            pass
    # detector stage of trackImage (CLAHE of the node + goodFeaturesToTrack, feature_tracker.cpp:198) on the same streams: steady state,
    # mask = MIN_DIST circles around the tracked points, 40 new corners wanted per stream
    mask = np.full((streams, 480, 640), 255, np.uint8)
    yy, xx = np.mgrid[:480, :640]
    for s in range(min(streams, len(base))):
        for p_ in base[s][2][:npts:2]:
            mask[s][(xx - int(p_[0])) ** 2 + (yy - int(p_[1])) ** 2 <= 900] = 0
    for s in range(len(base), streams):
        mask[s] = mask[s % len(base)]
    t.set_equalize(40.0, (8, 8))
    for _ in range(2):
        t.detect(cur, 40, mask=mask)
    det_ms = 0.0; t0 = time.perf_counter()
    for _ in range(steps):
        corners = t.detect(cur, 40, mask=mask); det_ms += t.last_timing()["detect_ms"]
    wall_d = time.perf_counter() - t0
    t1.set_equalize(40.0, (8, 8))
    one_d = []
    for _ in range(10):
        t0 = time.perf_counter(); t1.detect(cur[0], 40, mask=mask[0]); one_d.append((time.perf_counter() - t0) * 1e3)
    # algorithmic bytes per frame: image read twice (CLAHE histogram + apply), written once, mask read, score map written + read once
    bytes_det = 640 * 480 * (3 + 1 + 8)
    line["detect"] = {"metric": "CLAHE + goodFeaturesToTrack frames/sec (640x480, 8x8 tiles, 40 new corners, mask)", "value": streams * steps / wall_d, "unit": "frames/s",
                      "device_ms_per_batch": det_ms / steps, "single_stream_ms_per_frame": float(np.median(one_d)), "corners_per_frame": float(np.mean([len(c) for c in corners])),
                      "roofline": {"bound": "hbm", "kernel": "k_gftt_eig + k_gftt_nms + k_clahe_*", "achieved": bytes_det * streams / (det_ms / steps / 1e3) / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                                   "frac": bytes_det * streams / (det_ms / steps / 1e3) / 1e9 / peaks["hbm_gbs"], "algorithmic_bytes_per_frame": bytes_det,
                                   "note": "device_ms includes the H2D upload of image + mask; the column running sum of k_gftt_eig is sequential by construction (cv's rounding history)"}}
    if cv2_seconds > 0:
        try:
            import cv2
            cv2.setNumThreads(0)
            cl = cv2.createCLAHE()
            t0 = time.perf_counter(); n = 0
            while time.perf_counter() - t0 < cv2_seconds:
                s = n % len(base)
                cv2.goodFeaturesToTrack(cl.apply(base[s][1]), 40, 0.01, 30, mask=mask[s]); n += 1
            line["detect"]["cpu_baseline"] = {"value": n / (time.perf_counter() - t0), "unit": "frames/s", "kind": "reference", "sample": f"cv2 {cv2.__version__} CLAHE + goodFeaturesToTrack, all threads, one stream, {cv2_seconds:.0f} s"}
        except ImportError:
            pass
    t.close(); t1.close()
    line["stream"] = run_lk_stream(gf2, synth)
    return line


def run_lk_stream(gf2, synth, n_frames=300):
    """BASELINE.json config 3 as a STREAM: 300 frames of a 640x480 30 Hz synthetic sequence through FeatureTracker::trackImage of the C++ mirror
    (CLAHE off; forward 4-level LK + backward check on the device, border / status filtering, setMask, goodFeaturesToTrack top-up to
    max_cnt = 300 corners with min_dist = 20 on the device, undistortion + velocities on the host): one stream, frame after frame — latency."""
    import ctypes as C
    L = C.CDLL(os.path.join(ROOT, "ground-fusion2_b200", "libgf2_host.so"))
    L.gf2h_tracker_create.restype = C.c_void_p
    P_ = lambda a: np.ascontiguousarray(a).ctypes.data_as(C.c_void_p)
    imgs = synth.image_stream(0, n_frames)
    K = np.array([607.8, 607.8, 328.8, 245.5, 0.0, 0.0, 0.0, 0.0])
    t = C.c_void_p(L.gf2h_tracker_create(480, 640, 300, 20, P_(K)))
    depth = np.zeros((480, 640), np.uint16)
    out = np.zeros((400, 10)); ms = []; counts = []; ages = {}
    for k in range(n_frames):
        t0 = time.perf_counter()
        n = L.gf2h_tracker_track(t, C.c_double(k / 30.0), P_(imgs[k]), P_(depth), 400, P_(out))
        ms.append((time.perf_counter() - t0) * 1e3)
        if n < 0:
            raise RuntimeError("trackImage failed on the config-3 stream")
        counts.append(n)
        ids = out[:n, 0].astype(int)
        ages = {i: ages.get(i, 0) + 1 for i in ids}
    L.gf2h_tracker_destroy(t)
    med = float(np.median(ms[5:]))
    return {"metric": "FeatureTracker::trackImage frames/sec, one 640x480 stream of 300 frames, 300 corners (BASELINE config 3)", "value": 1e3 / med, "unit": "frames/s",
            "frames": n_frames, "median_ms_per_frame": med, "p95_ms_per_frame": float(np.percentile(ms[5:], 95)), "real_time_factor_at_30hz": (1e3 / med) / 30.0,
            "mean_features_per_frame": float(np.mean(counts)), "mean_track_age_frames": float(np.mean(list(ages.values()))) if ages else 0.0}


def run_lio(gf2, synth, steps=10, with_cpu=True):
    """LIO factor construction (lidarodom::addSurfCostFactor: voxel-hash kNN + PCA normals + residual gate) for one scan of 3000
    keypoints against a 100k-point map snapshot; the map is resident (set once), keypoints cross PCIe every call."""
    scene = synth.lio_scene(7, n_map_points=100000, n_keypoints=3000)
    o = gf2.abi.default_lio_opts(translation_begin=scene["translation_begin"], rotation=scene["rotation"], translation=scene["translation"])
    h = gf2.Lio(max_voxels=len(scene["keys"]), max_keypoints=len(scene["keypoints"]), max_points_per_voxel=scene["max_points_per_voxel"])
    h.set_map(scene["keys"], scene["n_points"], scene["points"])
    for _ in range(3):
        fac, _, _, _ = h.build_factors(scene["keypoints"], o)
    k_ms = 0.0; t0 = time.perf_counter()
    for _ in range(steps):
        fac, _, _, _ = h.build_factors(scene["keypoints"], o); k_ms += h.last_timing()["kernel_ms"]
    wall = (time.perf_counter() - t0) / steps
    nk = len(scene["keypoints"])
    # algorithmic bytes: every keypoint reads the <= 27 x 20 candidate points of its voxel neighbourhood once (24 B each) + its own record
    cand = 27 * float(scene["n_points"].mean()) * 24
    peaks, _ = measured_peaks()
    ach = nk * (cand + 56) / (k_ms / steps / 1e3) / 1e9
    line = {"metric": "LIO scans/sec (3000 keypoints, 27-voxel kNN-20 + PCA normal + gate)", "value": 1.0 / wall, "unit": "scans/s", "keypoints": nk, "voxels": int(len(scene["keys"])),
            "residuals": int(len(fac)), "kernel_ms": k_ms / steps, "call_ms": wall * 1e3,
            "roofline": {"bound": "hbm", "kernel": "k_lio_factors", "achieved": ach, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": ach / peaks["hbm_gbs"],
                         "note": "latency bound: 750 warps per scan, 20 dependent arg-min rounds each; the candidate points are L2 resident"}}
    # map maintenance on the device-resident map: addPointToMap for 30k-point scans (the third scan lands in a populated map)
    hm = gf2.Lio(max_voxels=60000, max_keypoints=8)
    scans = [synth.lio_scan(s_, 30000) for s_ in range(4)]
    for sc in scans[:3]:
        hm.add_points(sc)
    t0 = time.perf_counter(); hm.add_points(scans[3]); add_wall = time.perf_counter() - t0
    tm = hm.last_timing()
    line["map_insert"] = {"metric": "addPointToMap points/sec (30k-point scan into a populated device-resident map)", "value": 30000 / add_wall, "unit": "points/s",
                          "device_ms": tm["kernel_ms"], "call_ms": add_wall * 1e3, "voxels": tm["voxels"], "new_voxels": tm["new_voxels"]}
    if with_cpu:
        vox = {}
        synth.voxel_map_insert(vox, scans[0][:3000])
        t0 = time.perf_counter(); synth.voxel_map_insert(vox, scans[1][:3000]); dt = time.perf_counter() - t0
        line["map_insert"]["cpu_baseline"] = {"value": 3000 / dt, "unit": "points/s", "kind": "port", "cores": 1, "sample": "Python statement of addPointToMap, 3000 points (an interpreted baseline: indicative only)"}
    hm.close()
    if with_cpu:
        import gf2_oracle as orc
        t0 = time.perf_counter(); n = 0; loop_s = 0.0
        while time.perf_counter() - t0 < 2.0:
            tm = {}; orc.lio_build_factors(scene, o, timing=tm); loop_s += tm["loop_seconds"]; n += 1
        line["cpu_baseline"] = {"value": n / loop_s, "unit": "scans/s", "cores": 1, "kind": "port",
                                "sample": "restated addSurfCostFactor keypoint loop (std::unordered_map voxel map resident, its build excluded), 1 thread, 2 s"}
    h.close()
    return line


def run_replay_rgbd(gf2, synth):
    """The same loop fed by IMAGES: a ray-cast RGB-D + IMU stream through FeatureTracker::trackImage (CLAHE, LK + reverse check, detector on the
    device) and Estimator::processImage — BASELINE.json config 5 shape without wheel / LiDAR / GNSS; per-frame latency of one robot."""
    import ctypes as C
    L = C.CDLL(os.path.join(ROOT, "ground-fusion2_b200", "libgf2_host.so"))
    L.gf2h_estimator_create.restype = C.c_void_p; L.gf2h_tracker_create.restype = C.c_void_p; L.gf2h_last_error.restype = C.c_char_p
    P_ = lambda a: np.ascontiguousarray(a).ctypes.data_as(C.c_void_p)
    st = synth.render_stream(2, n_frames=50, pause=(30, 33))
    t = C.c_void_p(L.gf2h_tracker_create(480, 640, 150, 30, P_(st["intrinsics"]))); L.gf2h_tracker_set_equalize(t, 1)
    e = C.c_void_p(L.gf2h_estimator_create())
    L.gf2h_set_extrinsic(e, P_(st["tic"].copy()), P_(st["ric"].copy()), C.c_double(0.0), C.c_double(synth.G_NORM), P_(st["imu_noise"]))
    L.gf2h_set_flags(e, 1, 0, 1, 0); L.gf2h_set_min_parallax(e, C.c_double(10.0 / 460.0))
    Pp = st["gt_p"][:11].copy(); R = st["gt_R"][:11].copy(); V = st["gt_v"][:11].copy(); Pp[10] = Pp[9]; R[10] = R[9]; V[10] = V[9]
    fs = np.zeros((11, 21)); fs[:, 0:3] = Pp; fs[:, 3:12] = R.reshape(11, 9); fs[:, 12:15] = V
    L.gf2h_set_frame_states(e, P_(fs))

    def track(k):
        out = np.zeros((200, 10))
        n = L.gf2h_tracker_track(t, C.c_double(st["headers"][k]), P_(st["images"][k]), P_(st["depths"][k]), 200, P_(out))
        if n < 0:
            raise RuntimeError("tracker failed")
        order = np.argsort(out[:n, 0])
        return out[:n, 0][order].astype(np.int32), np.ascontiguousarray(out[:n, 1:9][order])
    for k in range(10):
        ids, pts = track(k); L.gf2h_add_image(e, k, len(ids), P_(ids), P_(pts), C.c_double(0.0))
    for j in range(1, 10):
        iv = st["imu"][j - 1]
        L.gf2h_new_interval(e, j, P_(iv["first"][:3].copy()), P_(iv["first"][3:].copy()), P_(np.zeros(3)), P_(np.zeros(3)))
        for s_ in iv["samples"]:
            L.gf2h_push_imu(e, j, C.c_double(s_["dt"]), P_(s_["acc"].copy()), P_(s_["gyr"].copy()))
    L.gf2h_set_imu0(e, P_(st["imu"][9]["first"][:3].copy()), P_(st["imu"][9]["first"][3:].copy()))
    pose = np.zeros((11, 7)); sbv = np.zeros((11, 9)); exv = np.zeros(7); L.gf2h_vector2double(e, P_(pose), P_(sbv), P_(exv))
    blk = np.zeros(1, gf2.abi.PRIOR_BLOCK); blk["kind"] = gf2.abi.BLK_POSE; blk["x0"][0, :7] = pose[0]
    L.gf2h_set_prior(e, 6, P_(np.eye(6) * 100.0), P_(np.zeros(6)), 1, P_(blk))
    t_track, t_proc, flags, errs = [], [], [], []
    for k in range(10, st["n_frames"]):
        for s_ in st["imu"][k - 1]["samples"]:
            L.gf2h_process_imu(e, C.c_double(0.0), C.c_double(s_["dt"]), P_(s_["acc"].copy()), P_(s_["gyr"].copy()))
        t0 = time.perf_counter(); ids, pts = track(k); t1 = time.perf_counter()
        flag = L.gf2h_process_image(e, len(ids), P_(ids), P_(pts), C.c_double(st["headers"][k])); t2 = time.perf_counter()
        if flag < 0:
            raise RuntimeError("replay: " + L.gf2h_last_error(e).decode())
        t_track.append(t1 - t0); t_proc.append(t2 - t1); flags.append(flag)
        out = np.zeros((11, 21)); L.gf2h_get_frame_states(e, P_(out)); errs.append(float(np.linalg.norm(out[9, :3] - st["gt_p"][k])))
    L.gf2h_tracker_destroy(t); L.gf2h_estimator_destroy(e)
    med = float(np.median(np.array(t_track[1:]) + np.array(t_proc[1:])))
    return {"metric": "RGB-D + IMU replay frames/sec, images in -> poses out (trackImage + processImage, one robot)", "value": 1.0 / med, "unit": "frames/s", "frames": len(flags),
            "keyframes": flags.count(0), "keyframes_per_s": flags.count(0) / float(np.sum(t_track[1:]) + np.sum(t_proc[1:])) if len(flags) > 1 else None,
            "median_track_ms": float(np.median(t_track[1:])) * 1e3, "median_process_ms": float(np.median(t_proc[1:])) * 1e3, "max_position_error_m": max(errs)}


def run_replay_full(gf2, synth):
    """BASELINE.json config 5: the full-fusion replay — features (RGB-D shaped) + 200 Hz IMU + 50 Hz wheel odometer + a 32-line 10 Hz LiDAR
    (GNSS is off in the shipped configs, gnss_enable: 0). Per frame: the scan's keypoints go through lidarodom::addSurfCostFactor on the device
    (gf2_lio_build_factors: voxel kNN + PCA normals + gate) against the device-resident map at the predicted pose; the resulting point-to-plane
    factors ride in the visual-inertial-wheel window on that frame's pose (Estimator::inputLidarPlanes -> gf2_set_planes, the config-4
    composition); Estimator::processImage solves + marginalizes; the scan is added to the map at the solved pose (gf2_lio_add_points).
    One robot, one window at a time: a latency figure."""
    import ctypes as C
    L = C.CDLL(os.path.join(ROOT, "ground-fusion2_b200", "libgf2_host.so"))
    L.gf2h_estimator_create.restype = C.c_void_p; L.gf2h_last_error.restype = C.c_char_p
    P_ = lambda a: np.ascontiguousarray(a).ctypes.data_as(C.c_void_p)
    abi = gf2.abi
    st = synth.feature_stream(3, n_frames=70, pause=(40, 43), wheel_hz=50)
    lid = synth.lidar_scans(st["gt_p"], st["gt_R"], seed=3)
    e = C.c_void_p(L.gf2h_estimator_create())
    L.gf2h_set_extrinsic(e, P_(st["tic"].copy()), P_(st["ric"].copy()), C.c_double(0.0), C.c_double(synth.G_NORM), P_(st["imu_noise"]))
    calib = np.concatenate([st["tio"], st["rio"].ravel(), [1.0, 1.0, 1.0, 0.0]])
    L.gf2h_set_wheel_parameters(e, P_(calib), P_(np.array([1.0, 0.0, 0.0, 0.0, st["wheel_noise"][0], st["wheel_noise"][1]])))
    L.gf2h_set_flags(e, 1, 1, 1, 0); L.gf2h_set_min_parallax(e, C.c_double(10.0 / 460.0))
    Pp = st["gt_p"][:11].copy(); R = st["gt_R"][:11].copy(); V = st["gt_v"][:11].copy(); Pp[10] = Pp[9]; R[10] = R[9]; V[10] = V[9]
    fs = np.zeros((11, 21)); fs[:, 0:3] = Pp; fs[:, 3:12] = R.reshape(11, 9); fs[:, 12:15] = V
    L.gf2h_set_frame_states(e, P_(fs))
    lio = gf2.Lio(max_voxels=200000, max_keypoints=4096)
    for f in range(10):
        fr = st["frames"][f]; L.gf2h_add_image(e, f, len(fr["ids"]), P_(fr["ids"]), P_(fr["pts"]), C.c_double(0.0))
        lio.add_points(lid["scans"][f] @ st["gt_R"][f].T + st["gt_p"][f])          # map bootstrap at the initial window's poses
    for j in range(1, 10):
        iv = st["imu"][j - 1]
        L.gf2h_new_interval(e, j, P_(iv["first"][:3].copy()), P_(iv["first"][3:].copy()), P_(np.zeros(3)), P_(np.zeros(3)))
        for s_ in iv["samples"]:
            L.gf2h_push_imu(e, j, C.c_double(s_["dt"]), P_(s_["acc"].copy()), P_(s_["gyr"].copy()))
        wv = st["wheel"][j - 1]
        L.gf2h_new_wheel_interval(e, j, P_(wv["first"][:3].copy()), P_(wv["first"][3:].copy()))
        for s_ in wv["samples"]:
            L.gf2h_push_wheel(e, j, C.c_double(s_["dt"]), P_(s_["vel"].copy()), P_(s_["gyr"].copy()))
    L.gf2h_set_imu0(e, P_(st["imu"][9]["first"][:3].copy()), P_(st["imu"][9]["first"][3:].copy()))
    w0 = st["wheel"][9]["first"]
    L.gf2h_process_wheel(e, C.c_double(0.0), C.c_double(0.0), P_(w0[:3].copy()), P_(w0[3:].copy()))
    pose = np.zeros((11, 7)); sbv = np.zeros((11, 9)); exv = np.zeros(7); L.gf2h_vector2double(e, P_(pose), P_(sbv), P_(exv))
    blk = np.zeros(1, abi.PRIOR_BLOCK); blk["kind"] = abi.BLK_POSE; blk["x0"][0, :7] = pose[0]
    L.gf2h_set_prior(e, 6, P_(np.eye(6) * 100.0), P_(np.zeros(6)), 1, P_(blk))
    t_lio, t_proc, t_map, flags, errs, nres = [], [], [], [], [], []
    for k in range(10, st["n_frames"]):
        for s_ in st["imu"][k - 1]["samples"]:
            L.gf2h_process_imu(e, C.c_double(0.0), C.c_double(s_["dt"]), P_(s_["acc"].copy()), P_(s_["gyr"].copy()))
        for s_ in st["wheel"][k - 1]["samples"]:
            L.gf2h_process_wheel(e, C.c_double(0.0), C.c_double(s_["dt"]), P_(s_["vel"].copy()), P_(s_["gyr"].copy()))
        t0 = time.perf_counter()
        out = np.zeros((11, 21)); L.gf2h_get_frame_states(e, P_(out))              # slot 10 = the newest frame, dead-reckoned by processWheel
        Rp = out[10, 3:12].reshape(3, 3); Pn = out[10, :3]
        scan = lid["scans"][k]
        kp = np.zeros(len(scan[::6]), abi.LIO_KEYPOINT); kp["raw_point"] = scan[::6]; kp["point"] = scan[::6] @ Rp.T + Pn
        q = synth.quat_from_R(Rp)
        o = abi.default_lio_opts(icp_model=abi.ICP_POINT_TO_PLANE, max_num_residuals=480, rotation=q, translation=Pn, translation_begin=Pn)
        fac, _, _, _ = lio.build_factors(kp, o)
        L.gf2h_input_lidar_planes(e, len(fac), P_(fac), C.c_double(31.622776601683793))
        t1 = time.perf_counter()
        fr = st["frames"][k]
        flag = L.gf2h_process_image(e, len(fr["ids"]), P_(fr["ids"]), P_(fr["pts"]), C.c_double(fr["header"]))
        t2 = time.perf_counter()
        if flag < 0:
            raise RuntimeError("full replay: " + L.gf2h_last_error(e).decode())
        L.gf2h_get_frame_states(e, P_(out))                                        # the newest frame sits in slot 9 after the slide
        Rs_ = out[9, 3:12].reshape(3, 3); Ps_ = out[9, :3]
        lio.add_points(scan @ Rs_.T + Ps_)
        t3 = time.perf_counter()
        t_lio.append(t1 - t0); t_proc.append(t2 - t1); t_map.append(t3 - t2); flags.append(flag); nres.append(len(fac))
        errs.append(float(np.linalg.norm(Ps_ - st["gt_p"][k])))
    L.gf2h_estimator_destroy(e); lio.close()
    tot = np.array(t_lio[1:]) + np.array(t_proc[1:]) + np.array(t_map[1:])
    kf = sum(1 for f_ in flags[1:] if f_ == 0)
    return {"metric": "full-fusion replay keyframes/sec (features + IMU 200 Hz + wheel 50 Hz + 32-line LiDAR 10 Hz; one robot)", "value": kf / float(tot.sum()), "unit": "keyframes/s",
            "frames": len(flags), "keyframes": flags.count(0), "frames_per_s": 1.0 / float(np.median(tot)), "median_ms_per_frame": float(np.median(tot)) * 1e3,
            "median_lio_factor_ms": float(np.median(t_lio[1:])) * 1e3, "median_process_image_ms": float(np.median(t_proc[1:])) * 1e3, "median_map_insert_ms": float(np.median(t_map[1:])) * 1e3,
            "lidar_points_per_scan": int(len(lid["scans"][0])), "lidar_residuals_per_frame": float(np.mean(nres)), "max_position_error_m": max(errs),
            "note": "GNSS is disabled in every shipped config (gnss_enable: 0) and is not part of the stream"}


def run_replay(gf2, synth, with_cpu=True):
    """BASELINE.json config 5 shape, visual-inertial part: a synthetic feature + IMU stream through the C++ mirror's
    Estimator::processIMU / processImage (keyframe decision, depth initialisation, device solve + marginalization, outlier check, slide),
    one window at a time as the reference's node runs it: a LATENCY figure (frames/s of one robot), not a batched throughput."""
    import ctypes as C
    L = C.CDLL(os.path.join(ROOT, "ground-fusion2_b200", "libgf2_host.so"))
    L.gf2h_estimator_create.restype = C.c_void_p; L.gf2h_last_error.restype = C.c_char_p
    P_ = lambda a: np.ascontiguousarray(a).ctypes.data_as(C.c_void_p)
    st = synth.feature_stream(1, n_frames=70, pause=(30, 34))
    e = C.c_void_p(L.gf2h_estimator_create())
    L.gf2h_set_extrinsic(e, P_(st["tic"].copy()), P_(st["ric"].copy()), C.c_double(0.0), C.c_double(synth.G_NORM), P_(st["imu_noise"]))
    L.gf2h_set_flags(e, 1, 0, 1, 0); L.gf2h_set_min_parallax(e, C.c_double(10.0 / 460.0))
    Pp = st["gt_p"][:11].copy(); R = st["gt_R"][:11].copy(); V = st["gt_v"][:11].copy(); Pp[10] = Pp[9]; R[10] = R[9]; V[10] = V[9]
    fs = np.zeros((11, 21)); fs[:, 0:3] = Pp; fs[:, 3:12] = R.reshape(11, 9); fs[:, 12:15] = V
    L.gf2h_set_frame_states(e, P_(fs))
    for f in range(10):
        fr = st["frames"][f]; L.gf2h_add_image(e, f, len(fr["ids"]), P_(fr["ids"]), P_(fr["pts"]), C.c_double(0.0))
    for j in range(1, 10):
        iv = st["imu"][j - 1]
        L.gf2h_new_interval(e, j, P_(iv["first"][:3].copy()), P_(iv["first"][3:].copy()), P_(np.zeros(3)), P_(np.zeros(3)))
        for s_ in iv["samples"]:
            L.gf2h_push_imu(e, j, C.c_double(s_["dt"]), P_(s_["acc"].copy()), P_(s_["gyr"].copy()))
    L.gf2h_set_imu0(e, P_(st["imu"][9]["first"][:3].copy()), P_(st["imu"][9]["first"][3:].copy()))
    pose = np.zeros((11, 7)); sbv = np.zeros((11, 9)); exv = np.zeros(7); L.gf2h_vector2double(e, P_(pose), P_(sbv), P_(exv))
    blk = np.zeros(1, gf2.abi.PRIOR_BLOCK); blk["kind"] = gf2.abi.BLK_POSE; blk["x0"][0, :7] = pose[0]
    L.gf2h_set_prior(e, 6, P_(np.eye(6) * 100.0), P_(np.zeros(6)), 1, P_(blk))
    steps, flags, errs, gaps = [], [], [], []
    for k in range(10, st["n_frames"]):
        for s_ in st["imu"][k - 1]["samples"]:
            L.gf2h_process_imu(e, C.c_double(0.0), C.c_double(s_["dt"]), P_(s_["acc"].copy()), P_(s_["gyr"].copy()))
        fr = st["frames"][k]
        # the marginalization of the previous frame runs asynchronously on the device; a live 10 Hz stream leaves it 100 ms before the next
        # image: wait for it OUTSIDE the latency clock (its duration is reported as background_marginalization_ms and counted in ms_per_frame)
        tg = time.perf_counter(); L.gf2h_finish_marginalization(e); gaps.append(time.perf_counter() - tg)
        t0 = time.perf_counter()
        flag = L.gf2h_process_image(e, len(fr["ids"]), P_(fr["ids"]), P_(fr["pts"]), C.c_double(fr["header"]))
        steps.append(time.perf_counter() - t0)
        if flag < 0:
            raise RuntimeError("replay: " + L.gf2h_last_error(e).decode())
        flags.append(flag)
        out = np.zeros((11, 21)); L.gf2h_get_frame_states(e, P_(out)); errs.append(float(np.linalg.norm(out[9, :3] - st["gt_p"][k])))
    L.gf2h_estimator_destroy(e)
    med = float(np.median(np.array(steps[1:]) + np.array(gaps[2:] + [gaps[-1]])))   # device-busy time per frame: processImage + the marginalization that follows it
    lat = float(np.median(steps[1:]))
    line = {"metric": "replay frames/sec through Estimator::processImage (one window at a time: single-robot latency)", "value": 1.0 / med, "unit": "frames/s",
            "frames": len(flags), "keyframes": flags.count(0), "median_ms_per_frame": med * 1e3, "median_process_image_latency_ms": lat * 1e3,
            "background_marginalization_ms": float(np.median(gaps[2:])) * 1e3,
            "latency_note": "processImage returns after the solve; the marginalization runs asynchronously on the device (the prior stays resident) and is complete long before the next image of a 10 Hz stream; median_ms_per_frame = latency + that background time (back-to-back throughput)",
            "first_frame_ms": steps[0] * 1e3,
            "max_position_error_m": max(errs), "features_per_frame": 150}
    if with_cpu:
        import gf2_oracle as orc
        w = synth.make_windows(1, n_landmarks=150)
        orc.imu_preintegrate(w)
        t0 = time.perf_counter(); orc.solve_batch(w, gf2.abi.default_opts()); orc.marginalize_window(w, 0, gf2.abi.default_opts(), mode=0); dt_ = time.perf_counter() - t0
        line["cpu_baseline"] = {"value": 1.0 / dt_, "unit": "frames/s", "cores": 1, "kind": "port", "sample": "restated solve + marginalization of one 150-landmark window (the glue around them is negligible)"}
    return line


def run_config4(gf2, synth, torch, dist, rank, world, local, B, steps, warmup):
    """BASELINE.json config 4: W10-F1000 + 10 wheel factors + 5,000 point-to-plane factors per window (IMU and wheel preintegrated on the
    device). world == 1: the windows solved on one GPU. world > 1: the factor-sharded mode of SURVEY 8(e) — every rank holds all frame
    states + IMU / wheel / prior factors of ALL B windows and the landmarks l mod N / planes k mod N; per linearisation ONE NCCL all-reduce
    reduce-scatter of the windows' 36.6 KB records, the reduced solve sharded by WINDOW (each rank factorises B / N systems), one all-gather
    of the steps (+ two small all-reduces per iteration); strong scaling over a fixed batch."""
    shard = importlib.import_module("gf2_b200.shard")
    abi = gf2.abi
    distinct = min(B, 8)
    base = synth.make_windows(distinct, config_id=4, n_landmarks=N_LM, wheel=True, n_planes=5000)
    tile = lambda d: {k: (np.concatenate([v] * ((B + distinct - 1) // distinct))[:B] if isinstance(v, np.ndarray) and v.ndim >= 1 and v.shape[0] == distinct
                          and k not in ("imu_noise", "wheel_noise") else v) for k, v in d.items()}
    mk = lambda d: gf2.Solver(B, d["n_frames"], d["max_landmarks"], d["max_obs"], max_planes=d["max_planes"], max_imu_samples=d["n_imu_samples"],
                              use_wheel=True, max_wheel_samples=d["n_wheel_samples"], device=local, max_prior_rows=d.get("prior_stride", 0))
    opts = abi.default_opts()
    stream = torch.cuda.current_stream()

    def timed(s, n_steps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        nccl_ms = 0.0; iters = 0
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0.record(stream)
        for _ in range(n_steps):
            s.restore(B); summ = s.solve(opts, B)
            nccl_ms += s.last_timing()["nccl_ms"]; iters += int(summ["iterations"].max())
        e1.record(stream); torch.cuda.synchronize()
        return e0.elapsed_time(e1), nccl_ms, iters, summ

    mine = tile(shard.shard_windows(base, rank, world)) if world > 1 else tile(base)
    s = mk(mine); s.set_stream(stream.cuda_stream)
    if world > 1:
        uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            uid.copy_(torch.from_numpy(gf2.Solver.comm_unique_id()).cuda())
        dist.broadcast(uid, 0)
        # NCCL prints its version banner on stdout at the first communicator it creates; stdout carries the one JSON line, so send the banner to stderr
        sys.stdout.flush(); saved = os.dup(1); os.dup2(2, 1)
        try:
            s.comm_init(rank, world, uid.cpu().numpy())
        finally:
            os.dup2(saved, 1); os.close(saved)
    s.upload(mine, preintegrate="device"); s.snapshot(B)
    for _ in range(warmup):
        s.restore(B); s.solve(opts, B)
    ms, nccl_ms, iters, summ = timed(s, steps)
    tt = torch.tensor([ms], device="cuda")
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    ms = float(tt.item())
    st = s.get_states(B)
    s.close()
    line = {"metric": "solves/sec, BASELINE config 4 (W10-F1000 + 10 wheel + 5,000 point-to-plane factors, 8 iterations)", "value": B * steps / (ms / 1e3), "unit": "solves/s",
            "windows": B, "distinct_windows": distinct, "n_gpus": world, "ms_per_step": ms / steps, "iterations_per_step": iters / steps,
            "mode": "single GPU" if world == 1 else f"factor-sharded over {world} GPUs (landmarks l mod {world}, planes k mod {world}; frame states, IMU / wheel / prior replicated)",
            "timing": "CUDA events on the launching stream, max over ranks; step = restore + solve, preintegration records resident"}
    if world > 1:
        it = max(iters / steps, 1.0)
        line.update({"scaling": "strong",
                     "nccl_ms_per_step_rank0": nccl_ms / steps, "nccl_share_rank0": nccl_ms / ms if rank == 0 else None,
                     "allreduce_bytes_per_iteration": int(B * (4576 * 8 + 8 + 64 + 32)),
                     "collectives_per_iteration": "after the sweep: reduce-scatter (SUM) of the B x 36,608 B window records + reduce-scatter (MAX) of B x 8 B, grouped; each rank then solves its B / N windows; all-gather of the steps and trust-region states (B x (2 x 1320 + 304) B); all-reduce of B x 64 B after back-substitution and of B x 32 B after the candidate",
                     "nccl_us_per_iteration_rank0": 1e3 * nccl_ms / steps / it})
        if rank == 0:   # the same windows unsharded on rank 0's GPU alone: parity and the strong-scaling reference
            full = tile(base)
            ref = mk(full); ref.set_stream(stream.cuda_stream); ref.upload(full, preintegrate="device"); ref.snapshot(B)
            for _ in range(2):
                ref.restore(B); ref.solve(opts, B)
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(steps):
                ref.restore(B); rsum = ref.solve(opts, B)
            e1.record(stream); torch.cuda.synchronize()
            ms1 = e0.elapsed_time(e1)
            rst = ref.get_states(B); ref.close()
            line["single_gpu"] = {"value": B * steps / (ms1 / 1e3), "ms_per_step": ms1 / steps}
            line["speedup_vs_single_gpu"] = ms1 / ms
            line["parity_vs_single_gpu"] = {"pose_max_abs_diff": float(np.abs(st["para_pose"] - rst["para_pose"]).max()),
                                            "iterations_equal": bool((summ["iterations"] == rsum["iterations"]).all()),
                                            "termination_equal": bool((summ["termination"] == rsum["termination"]).all())}
        dist.barrier()
    return line


def run_single_window(gf2, synth, device, reps=12):
    """Latency of ONE window per call (one robot): the 8-iteration solve with the sweep kernel forced to the batch kernel (k_linearize: 4 warps per
    window) and to the window kernel (k_linearize_ws: 16 warp-specialised warps per window, what GF2_SWEEP_AUTO picks at this size)."""
    abi = gf2.abi
    out = {"metric": "ms per 8-iteration solve of one window (device time, CUDA events), by sweep kernel", "unit": "ms"}
    for nl in (1000, 150):
        w = synth.make_windows(1, n_landmarks=nl, prior_stride=PRIOR_STRIDE)
        rec = {}
        for name, sweep in (("batch_kernel", abi.SWEEP_BATCH), ("window_kernel", abi.SWEEP_WINDOW)):
            s = gf2.Solver(1, N_FRAMES, w["max_landmarks"], w["max_obs"], max_imu_samples=w["n_imu_samples"], device=device, max_prior_rows=PRIOR_STRIDE, sweep=sweep)
            s.upload(w, preintegrate="device"); s.snapshot(1)
            opts = abi.default_opts()
            tot, lin = [], []
            for _ in range(reps):
                s.restore(1); s.solve(opts, 1)
                t = s.last_timing(); tot.append(t["total_ms"]); lin.append(t["linearize_ms"] / max(t["linearize_launches"], 1))
            s.close()
            rec[name] = {"solve_ms": float(np.median(tot[2:])), "sweep_ms_per_iteration": float(np.median(lin[2:]))}
        out[f"{nl}_landmarks"] = rec
    return out


def run_reference(args, rank, world):
    """Restated-reference CPU baseline: oracle solve (same algorithm as ceres::Solve with the reference's options) on
    all host threads, each step a bounded sample of the same workload."""
    if rank != 0:
        return
    from gf2_loader import load
    gf2 = load()
    synth = importlib.import_module("gf2_b200.synth")
    import gf2_oracle as orc
    T = host_threads()
    n = max(T, 8) * 24  # windows per step: a bounded sample of the workload, ~0.1 s per window per thread => ~2.5 s per step
    w = make_batch(gf2, synth, n, min(args.distinct, n), pinned=False)
    orc.imu_preintegrate(w)
    opts = gf2.abi.default_opts()
    init = {k: w[k].copy() for k in ("para_pose", "para_speedbias", "inv_depth")}

    def step():
        for k, v in init.items():
            w[k][...] = v
        orc.imu_preintegrate(w)  # the reference preintegrates on the host too (processIMU)
        orc.solve_batch(w, opts, n_threads=T)
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    val = n * args.steps / dt
    line = {"impl": "reference", "metric": "sliding-window solves/sec (10-frame, 1k-feat)", "value": val, "unit": "solves/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": bench_config(args.windows, min(args.distinct, args.windows), max(world, 1)),
            "cpu_baseline": {"value": val, "unit": "solves/s", "cores": T, "kind": "port",
                             "sample": f"bounded sample of the workload: {n} windows/step x {args.steps} steps, restated-reference CPU baseline (Ceres unavailable; the restatement is pinned "
                                       f"to the reference's own factor code, oracle/_ref), {T} threads over independent windows"},
            "e2e": {"value": val, "unit": "solves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--windows", type=int, default=4096, help="windows per GPU per step (batch B)")
    ap.add_argument("--distinct", type=int, default=64, help="distinct generated windows tiled to B")
    ap.add_argument("--e2e-chunks", type=int, default=4, help="chunks (handle + stream + host thread each) the e2e leg pipelines the batch over")
    ap.add_argument("--e2e-obs", default="xy", choices=["xy", "full"], help="observation records of the e2e leg: positions only (8 B) or with velocities (16 B)")
    ap.add_argument("--impl", default="gf2", choices=["gf2", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-marginalize", action="store_true", help="skip the marginalization leg")
    ap.add_argument("--no-lk", action="store_true", help="skip the front-end (LK) leg")
    ap.add_argument("--no-config4", action="store_true", help="skip the config-4 leg (wheel + LiDAR planes; factor-sharded over the ranks when N > 1)")
    ap.add_argument("--config4-windows", type=int, default=1184, help="windows of the config-4 leg (fixed total: strong scaling when N > 1)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "gf2" else args.warmup
    rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); local = int(os.environ.get("LOCAL_RANK", 0))
    if args.impl == "reference":
        return run_reference(args, rank, world)

    import torch
    import torch.distributed as dist
    from gf2_loader import load
    gf2 = load()
    synth = importlib.import_module("gf2_b200.synth")
    abi = gf2.abi
    if gf2.device_count() < 1:
        raise SystemExit("bench.py needs a CUDA device: the hot path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    B = args.windows
    w = make_batch(gf2, synth, B, min(args.distinct, B), first_window=rank * 100000)
    s = gf2.Solver(B, N_FRAMES, w["max_landmarks"], w["max_obs"], max_imu_samples=w["n_imu_samples"], device=local, max_prior_rows=w["prior_stride"])
    stream = torch.cuda.Stream()  # a real (non-default) stream: the library launches on it and the timing events sit on it
    torch.cuda.set_stream(stream)
    s.set_stream(stream.cuda_stream)
    opts = abi.default_opts()
    noise = w["imu_noise"]
    summ = gf2.pinned_empty((B,), abi.SUMMARY)
    out_states = {"para_pose": gf2.pinned_empty((B, N_FRAMES, 7), "f8"), "para_speedbias": gf2.pinned_empty((B, N_FRAMES, 9), "f8"),
                  "ex_pose": gf2.pinned_empty((B, 7), "f8"), "td": gf2.pinned_empty((B,), "f8")}
    out_lam = gf2.pinned_empty((B, w["max_landmarks"]), "f8")

    # ------------------------------------------------------------ resident run (value)
    s.upload(w, preintegrate="device")
    s.snapshot(B)
    torch.cuda.synchronize()

    def step_resident():
        s.restore(B)
        s.imu_preintegrate_resident(noise, B)
        s.solve(opts, B, summaries=summ)
    lin_ms = solve_ms = step_ms = total_ms = 0.0
    launches = 0
    for _ in range(args.warmup):
        step_resident()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    sampler = ClockSampler(local) if rank == 0 else None
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        step_resident()
        t = s.last_timing()
        lin_ms += t["linearize_ms"]; solve_ms += t["solve_ms"]; step_ms += t["step_ms"]; total_ms += t["total_ms"]
        launches += t["launches"] + 1
        lin_launches = t["linearize_launches"]
    e1.record(stream)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    clocks = sampler.stop() if sampler else None
    ms = e0.elapsed_time(e1)
    tt = torch.tensor([ms], device="cuda")
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    ms_max = float(tt.item())
    value = world * B * args.steps / (ms_max / 1e3)
    ok_frac = float((summ["final_cost"] < 1e-3 * summ["initial_cost"]).mean())

    # ------------------------------------------------------------ end-to-end run through the ABI with host buffers
    # Observations travel as positions only (gf2_set_observations_xy, 8 B instead of 16): ESTIMATE_TD = 0 and td == cur_td of every frame in
    # this workload, so the velocities never enter the residual and the solve is bit-equal (tests/test_gpu_solver.py); --e2e-obs full sends the records.
    xy_only = args.e2e_obs == "xy" and bool(np.all(w["td"][:, None] == w["frame_td"]))
    if xy_only:
        obs_xy = gf2.pinned_empty((B, w["max_obs"], 2), np.float32)
        obs_xy[..., 0] = w["obs"]["x"]; obs_xy[..., 1] = w["obs"]["y"]
    h2d = sum(w[k].nbytes for k in ("para_pose", "para_speedbias", "ex_pose", "td", "n_landmarks", "inv_depth", "start_frame", "track_len",
                                    "fixed", "frame_td", "imu_samples", "imu_n", "imu_first", "imu_lin_bias", "prior_rows",
                                    "prior_J0", "prior_r0", "prior_nblocks", "prior_blocks")) + (obs_xy.nbytes if xy_only else w["obs"].nbytes)
    d2h = sum(v.nbytes for v in out_states.values()) + out_lam.nbytes + summ.nbytes

    # The batch is cut into `--e2e-chunks` chunks; each chunk has its own handle + stream and is driven by its own host
    # thread (ctypes releases the GIL), so the H2D copy of one chunk overlaps the kernels of another and the D2H of a third.
    s.close()
    nch = max(1, min(args.e2e_chunks, B))
    bounds = [B * c // nch for c in range(nch + 1)]
    chunks = []
    for c in range(nch):
        lo, hi = bounds[c], bounds[c + 1]
        cs = gf2.Solver(hi - lo, N_FRAMES, w["max_landmarks"], w["max_obs"], max_imu_samples=w["n_imu_samples"], device=local,
                        max_prior_rows=w["prior_stride"])
        cw = {k: (v[lo:hi] if isinstance(v, np.ndarray) and v.ndim >= 1 and v.shape[0] == B and k != "imu_noise" else v) for k, v in w.items()}
        if xy_only:
            del cw["obs"]; cw["obs_xy"] = obs_xy[lo:hi]
        co = {k: v[lo:hi] for k, v in out_states.items()}
        chunks.append((cs, cw, co, out_lam[lo:hi], summ[lo:hi], hi - lo))
    from concurrent.futures import ThreadPoolExecutor
    pool = ThreadPoolExecutor(nch)

    def chunk_step(c):
        cs, cw, co, cl, csum, n = chunks[c]
        cs.upload(cw, preintegrate="device")
        cs.solve(opts, n, summaries=csum)
        cs.get_states(n, out=co)
        cs.get_landmarks(n, out=cl)

    def run_e2e(steps):
        # every step passes all B windows through upload -> solve -> download; the chunk threads are not re-synchronised
        # between steps, so chunk c of step k+1 may start while chunk c' of step k is still in flight (a streaming server)
        def loop(c):
            for _ in range(steps):
                chunk_step(c)
        list(pool.map(loop, range(nch)))
    run_e2e(2)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    run_e2e(args.steps)
    torch.cuda.synchronize()
    wall_e2e = time.perf_counter() - t0   # host wall clock around blocking calls + device synchronize: bounds the device time from above
    te = torch.tensor([wall_e2e], device="cuda")
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = world * B * args.steps / float(te.item())

    # ------------------------------------------------------------ BASELINE config 4 (wheel + LiDAR planes): one GPU at N = 1, factor-sharded + NCCL at N > 1
    cfg4 = None
    if not args.no_config4:
        cfg4 = run_config4(gf2, synth, torch, dist, rank, world, local, args.config4_windows, max(2, min(args.steps, 5)), 2)

    # ------------------------------------------------------------ marginalization leg, reported beside the metric
    marg = None
    if rank == 0 and not args.no_marginalize:
        nm = min(B, 1024)
        wm = make_batch(gf2, synth, nm, min(args.distinct, nm), pinned=False, prior_stride=80)
        sm = gf2.Solver(nm, N_FRAMES, wm["max_landmarks"], wm["max_obs"], max_imu_samples=wm["n_imu_samples"], device=local, max_prior_rows=80)
        sm.upload(wm, preintegrate="device")
        sm.solve(opts, nm)
        ms_m = []
        for _ in range(4):
            sm.set_prior(wm)                      # marginalization replaces the resident prior: put the anchor back
            st_m, m_m = sm.marginalize(opts, abi.MARGIN_OLD, nm)
            ms_m.append(sm.last_marginalize_ms())
        marg = {"windows": nm, "ms": float(np.median(ms_m[1:])), "windows_per_s": nm / (float(np.median(ms_m[1:])) / 1e3), "ok_fraction": float((st_m == 0).mean()),
                "m_dims": int(m_m[0]), "kept_dims": int(sm.get_prior(1)["prior_rows"][0]), "timing": "CUDA events on the handle's stream (k_prepare + k_marg_build + k_marg_eig)"}
        if not args.no_cpu_baseline:
            import gf2_oracle as orc
            wc = synth.make_windows(4, n_landmarks=N_LM)
            orc.imu_preintegrate(wc)
            t0 = time.perf_counter()
            for i in range(4):
                orc.marginalize_window(wc, i, opts, mode=0)
            marg["cpu_port_ms_per_window_1thread"] = 1e3 * (time.perf_counter() - t0) / 4
        sm.close()

    # ------------------------------------------------------------ front-end leg (BASELINE.json config 3), reported beside the metric
    lk_line = None
    if rank == 0 and not args.no_lk:
        lk_line = run_lk(gf2, synth, streams=64, steps=10, cv2_seconds=0.0 if args.no_cpu_baseline else 2.0)

    lio_line = replay_line = None
    if rank == 0 and not args.no_lk:
        lio_line = run_lio(gf2, synth, steps=10, with_cpu=not args.no_cpu_baseline)
        replay_line = run_replay(gf2, synth, with_cpu=not args.no_cpu_baseline)
        replay_line["rgbd"] = run_replay_rgbd(gf2, synth)
        replay_line["full_fusion"] = run_replay_full(gf2, synth)
        replay_line["single_window"] = run_single_window(gf2, synth, local)

    if rank == 0:
        peaks, which = measured_peaks()
        n_lin = lin_launches * args.steps
        lin_avg_ms = lin_ms / n_lin
        achieved = BYTES_SWEEP * B / (lin_avg_ms / 1e3) / 1e9
        ncu = ncu_summary() or {}
        klin = (ncu.get("kernels") or {}).get("k_linearize") or {}
        traffic = klin.get("dram_bytes_per_window")
        line = {
            "metric": "sliding-window solves/sec (10-frame, 1k-feat)", "value": value, "unit": "solves/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": bench_config(B, min(args.distinct, B), world),
            "converged_fraction": ok_frac,
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "solves/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "chunks": nch, "observations": "positions only, 8 B (gf2_set_observations_xy; td == cur_td)" if xy_only else "full records, 16 B",
                    "timing": "host wall clock around the blocking ABI calls of all chunks + device synchronize"},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "kernel": "k_linearize", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                         "frac": achieved / peaks["hbm_gbs"], "traffic": (traffic * B if traffic else None),
                         "traffic_source": (f"profiles/ncu_summary_r2.json ({ncu.get('captured')}, {ncu.get('command')}): dram__bytes_read.sum + dram__bytes_write.sum of one k_linearize launch per window x windows per launch" if traffic else None),
                         "peak_source": which, "algorithmic_bytes_per_launch": BYTES_SWEEP * B, "avg_launch_ms": lin_avg_ms,
                         "fp64_pipe": {"note": "DFMA and the fp64 mma share one datapath on this part (profiles/ubench_fp64_pipes_r2.txt: 12.1 + 24.3 = 36.4 TFLOP/s mixed; "
                                               "profiles/ubench_fp64_latency_r2.txt: a DFMA holds a sub-partition's share for 2 cycles, an m8n8k4 DMMA for 16); the kernel's real bound is its total fp64 work",
                                       "dmma_pipe_pct_ncu": klin.get("dmma_pipe_pct"), "fp64_fma_pipe_pct_ncu": klin.get("fp64_pipe_pct"), "warps_active_pct_ncu": klin.get("warps_active_pct"),
                                       # per window (profiles/ncu_opcodes_k_linearize_r1.txt, same arithmetic): 39 k DP warp instructions x 2 + 9.5 k DMMA x 16 sub-partition cycles, 4 sub-partitions per SM
                                       "datapath_sm_cycles_per_window": FP64_SM_CYCLES_PER_WINDOW,
                                       "datapath_bound_ms_per_launch": fp64_bound_ms(B, clocks),
                                       "datapath_utilisation": (fp64_bound_ms(B, clocks) / lin_avg_ms) if fp64_bound_ms(B, clocks) else None,
                                       "hbm_frac_at_full_datapath": (BYTES_SWEEP * B / (fp64_bound_ms(B, clocks) / 1e3) / 1e9 / peaks["hbm_gbs"]) if fp64_bound_ms(B, clocks) else None}},
            "reduced_solve": {"kernels": "k_nonvis + k_solve2", "avg_ms_per_iteration": solve_ms / n_lin, "bound": "serial pivot chain (latency), not the tensor pipe",
                              "dmma_pipe_pct_ncu": ((ncu.get("kernels") or {}).get("k_solve2") or {}).get("dmma_pipe_pct"), "source": "profiles/ncu_summary_r2.json"},
            ("config4_sharded" if world > 1 else "config4"): cfg4,
            "marginalize": marg, "lk": lk_line, "lio": lio_line, "replay": replay_line,
            "phase_ms_per_step": {"linearize": lin_ms / args.steps, "reduced_solve": solve_ms / args.steps, "backsub_candidate": step_ms / args.steps,
                                  "solve_total": total_ms / args.steps},
        }
        if not args.no_cpu_baseline:
            import gf2_oracle as orc
            T = host_threads()
            n = max(T, 8) * 96  # bounded sample: ~0.1 s per window per thread => ~10 s of CPU work on every host thread
            wc = make_batch(gf2, synth, n, min(args.distinct, n), pinned=False)
            t0 = time.perf_counter()
            orc.imu_preintegrate(wc)
            orc.solve_batch(wc, opts, n_threads=T)
            dt = time.perf_counter() - t0
            line["cpu_baseline"] = {"value": n / dt, "unit": "solves/s", "cores": T, "kind": "port",
                                    "sample": f"{n} W10-F1000 windows, restated-reference CPU baseline (Ceres unavailable), {T} threads over independent windows, {dt:.1f} s"}
        print(json.dumps(line))
    pool.shutdown()
    for ch in chunks:
        ch[0].close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
