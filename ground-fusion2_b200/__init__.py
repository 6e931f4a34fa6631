"""gf2_b200 — host-side Python mirror of the C ABI in include/gf2_abi.h (ctypes over libgf2_b200.so).

The library holds only hand-written sm_100a kernels; there is NO CPU fallback: if the shared library is missing or
no CUDA device is present, every compute entry point raises."""
import ctypes as C
import os

import numpy as np

from . import abi  # noqa: F401

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libgf2_b200.so")
_lib = None


class Gf2Error(RuntimeError):
    pass


def lib():
    """Load libgf2_b200.so (built in-tree by build.py / __graft_entry__.build())."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise Gf2Error(f"{LIB_PATH} is missing: run `python __graft_entry__.py build` (the CUDA extension is the only "
                           "implementation; there is no CPU fallback)")
        _lib = C.CDLL(LIB_PATH)
        _lib.gf2_last_error.restype = C.c_char_p
        _lib.gf2_host_alloc.restype = C.c_void_p
        _lib.gf2_last_marginalize_ms.restype = C.c_double
        _lib.gf2_last_marginalize_ms.argtypes = [C.c_void_p]
        _lib.gf2_host_alloc.argtypes = [C.c_size_t]
        _lib.gf2_host_free.argtypes = [C.c_void_p]
    return _lib


ABI_SYMBOLS = [
    "gf2_last_error", "gf2_abi_version", "gf2_device_count", "gf2_solver_create", "gf2_solver_destroy", "gf2_set_states",
    "gf2_solver_set_stream", "gf2_snapshot_states", "gf2_restore_states", "gf2_host_alloc", "gf2_host_free",
    "gf2_imu_preintegrate_resident", "gf2_get_trace",
    "gf2_set_landmarks", "gf2_set_observations_xy", "gf2_set_imu", "gf2_imu_preintegrate", "gf2_get_imu", "gf2_set_wheel", "gf2_wheel_preintegrate", "gf2_get_wheel", "gf2_set_prior",
    "gf2_set_planes", "gf2_set_plane_alpha", "gf2_marginalize", "gf2_marginalize_async", "gf2_marginalize_wait", "gf2_get_prior", "gf2_last_marginalize_ms", "gf2_solve", "gf2_linearize", "gf2_reduced_dim", "gf2_get_reduced_system", "gf2_get_states",
    "gf2_get_landmarks", "gf2_comm_init", "gf2_comm_unique_id", "gf2_last_timing", "gf2_tracker_create",
    "gf2_tracker_destroy", "gf2_tracker_track", "gf2_tracker_track_fb", "gf2_tracker_last_timing",
    "gf2_tracker_track_image", "gf2_tracker_detect", "gf2_tracker_min_eigen_map", "gf2_detect_select",
    "gf2_tracker_equalize", "gf2_tracker_set_equalize", "gf2_tracker_get_image",
    "gf2_lio_create", "gf2_lio_destroy", "gf2_lio_set_map", "gf2_lio_build_factors", "gf2_lio_last_timing",
    "gf2_lio_add_points", "gf2_lio_map_size", "gf2_lio_get_map",
]


def _check(rc):
    if rc != 0:
        raise Gf2Error(f"gf2 error {rc}: {lib().gf2_last_error().decode()}")


def device_count():
    return lib().gf2_device_count()


def pinned_empty(shape, dtype):
    """numpy array backed by cudaHostAlloc'ed (pinned) memory, for full-speed H2D/D2H through the ABI."""
    dtype = np.dtype(dtype)
    n = int(np.prod(shape)) * dtype.itemsize
    p = lib().gf2_host_alloc(max(n, 1))
    if not p:
        raise Gf2Error("cudaHostAlloc failed")
    buf = (C.c_char * max(n, 1)).from_address(p)
    arr = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)
    return arr


_p = abi.ptr


class Solver:
    """Batched sliding-window solver: the ceres::Solve call of Estimator::optimization() for `max_windows` windows."""

    def __init__(self, max_windows, n_frames=11, max_landmarks=1000, max_obs=7500, max_planes=0, max_imu_samples=0,
                 use_wheel=False, device=0, max_prior_rows=0, max_wheel_samples=0, sweep=0):
        cfg = abi.SolverCfg()
        cfg.sweep = sweep   # abi.SWEEP_AUTO / SWEEP_BATCH / SWEEP_WINDOW
        cfg.device = device; cfg.max_windows = max_windows; cfg.n_frames = n_frames; cfg.max_landmarks = max_landmarks
        cfg.max_obs = max_obs; cfg.max_planes = max_planes; cfg.max_imu_samples = max_imu_samples; cfg.max_wheel_samples = max_wheel_samples
        cfg.use_wheel = 1 if use_wheel else 0
        cfg.max_prior_rows = max_prior_rows
        self.Pr = max_prior_rows or abi.MAX_PRIOR_DIM
        self.cfg = cfg
        self.h = C.c_void_p()
        _check(lib().gf2_solver_create(C.byref(cfg), C.byref(self.h)))
        self.B, self.F, self.Lm, self.Om = max_windows, n_frames, max_landmarks, max_obs

    def close(self):
        if self.h:
            lib().gf2_solver_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- uploads (arrays are window-major, see include/gf2_abi.h)
    def set_states(self, w, first=0, n=None):
        n = n if n is not None else w["para_pose"].shape[0]
        _check(lib().gf2_set_states(self.h, first, n, _p(w["para_pose"]), _p(w["para_speedbias"]), _p(w["ex_pose"]), _p(w["td"]),
                                    _p(w.get("ex_pose_wheel")), _p(w.get("sxsysw")), _p(w.get("td_wheel"))))

    def set_landmarks(self, w, first=0, n=None):
        """w["obs"] (full records) or, without it, w["obs_xy"] ([n][max_obs][2] float32: positions only, half the bytes; valid while
        td equals every frame's cur_td — see gf2_set_observations_xy)."""
        n = n if n is not None else w["para_pose"].shape[0]
        xy_only = "obs" not in w
        ob = w["obs_xy"] if xy_only else w["obs"]
        assert w["inv_depth"].shape[1] == self.Lm and ob.shape[1] == self.Om, "array strides must equal the solver capacities"
        _check(lib().gf2_set_landmarks(self.h, first, n, _p(w["n_landmarks"]), _p(w["inv_depth"]), _p(w["start_frame"]),
                                       _p(w["track_len"]), _p(w["fixed"]), None if xy_only else _p(ob), _p(w["frame_td"])))
        if xy_only:
            assert ob.dtype == np.float32 and ob.shape[2:] == (2,) and ob.flags.c_contiguous
            _check(lib().gf2_set_observations_xy(self.h, first, n, _p(ob)))

    def set_imu(self, rec, first=0):
        _check(lib().gf2_set_imu(self.h, first, rec.shape[0], _p(rec)))

    def imu_preintegrate(self, w, first=0):
        n = w["imu_n"].shape[0]
        noise = np.ascontiguousarray(w["imu_noise"], dtype=np.float64)
        _check(lib().gf2_imu_preintegrate(self.h, first, n, _p(w["imu_samples"]), _p(w["imu_n"]), _p(w["imu_first"]),
                                          _p(w["imu_lin_bias"]), _p(noise)))

    def imu_preintegrate_resident(self, noise, n=None, first=0):
        noise = np.ascontiguousarray(noise, dtype=np.float64)
        _check(lib().gf2_imu_preintegrate_resident(self.h, first, n if n is not None else self.B, _p(noise)))

    def set_stream(self, cuda_stream_ptr):
        _check(lib().gf2_solver_set_stream(self.h, C.c_void_p(cuda_stream_ptr)))

    def snapshot(self, n=None, first=0):
        _check(lib().gf2_snapshot_states(self.h, first, n if n is not None else self.B))

    def restore(self, n=None, first=0):
        _check(lib().gf2_restore_states(self.h, first, n if n is not None else self.B))

    def get_imu(self, n, first=0):
        rec = np.zeros((n, self.F - 1), abi.IMU_PREINT)
        _check(lib().gf2_get_imu(self.h, first, n, _p(rec)))
        return rec

    def set_wheel(self, rec, first=0):
        _check(lib().gf2_set_wheel(self.h, first, rec.shape[0], _p(rec)))

    def wheel_preintegrate(self, w, first=0):
        n = w["wheel_n"].shape[0]
        noise = np.ascontiguousarray(w["wheel_noise"], dtype=np.float64)
        _check(lib().gf2_wheel_preintegrate(self.h, first, n, _p(w["wheel_samples"]), _p(w["wheel_n"]), _p(w["wheel_first"]),
                                            _p(w["wheel_lin"]), _p(noise)))

    def get_wheel(self, n, first=0):
        rec = np.zeros((n, self.F - 1), abi.WHEEL_PREINT)
        _check(lib().gf2_get_wheel(self.h, first, n, _p(rec)))
        return rec

    def set_prior(self, w, first=0):
        n = w["prior_rows"].shape[0]
        assert w["prior_J0"].shape[1] == self.Pr, "prior arrays must use the solver's max_prior_rows as stride"
        _check(lib().gf2_set_prior(self.h, first, n, _p(w["prior_rows"]), _p(w["prior_J0"]), _p(w["prior_r0"]),
                                   _p(w["prior_nblocks"]), _p(w["prior_blocks"])))

    def marginalize(self, opts, mode=0, n=None, first=0):
        """gf2_marginalize: returns (status [n], m [n]); the new prior stays resident (get_prior downloads it)."""
        n = n if n is not None else self.B
        status = np.zeros(n, np.int32); m = np.zeros(n, np.int32)
        _check(lib().gf2_marginalize(self.h, first, n, int(mode), C.byref(opts), _p(status), _p(m)))
        return status, m

    def get_prior(self, n=None, first=0):
        n = n if n is not None else self.B
        out = {"prior_rows": np.zeros(n, np.int32), "prior_J0": np.zeros((n, self.Pr, self.Pr)), "prior_r0": np.zeros((n, self.Pr)),
               "prior_nblocks": np.zeros(n, np.int32), "prior_blocks": np.zeros((n, 2 * self.F + 8), abi.PRIOR_BLOCK)}
        _check(lib().gf2_get_prior(self.h, first, n, _p(out["prior_rows"]), _p(out["prior_J0"]), _p(out["prior_r0"]),
                                   _p(out["prior_nblocks"]), _p(out["prior_blocks"])))
        return out

    def last_marginalize_ms(self):
        return float(lib().gf2_last_marginalize_ms(self.h))

    def set_planes(self, w, first=0):
        _check(lib().gf2_set_planes(self.h, first, w["n_planes"].shape[0], _p(w["n_planes"]), _p(w["planes"])))
        if w.get("plane_alpha") is not None:   # alpha_time of the CTLidarPlaneNormFactor records (ct == 1)
            _check(lib().gf2_set_plane_alpha(self.h, first, w["n_planes"].shape[0], _p(np.ascontiguousarray(w["plane_alpha"], np.float64))))

    def upload(self, w, first=0, preintegrate="auto"):
        """Everything a synth window dict holds. preintegrate: 'device' (raw samples -> kernel), 'records' (w['imu'])."""
        self.set_states(w, first)
        self.set_landmarks(w, first)
        if preintegrate == "auto":
            preintegrate = "records" if "imu" in w else "device"
        if preintegrate == "device":
            self.imu_preintegrate(w, first)
        else:
            self.set_imu(w["imu"], first)
        if w.get("use_wheel") and "wheel" in w:
            self.set_wheel(w["wheel"], first)
        elif w.get("use_wheel") and self.cfg.max_wheel_samples > 0:
            self.wheel_preintegrate(w, first)
        self.set_prior(w, first)
        if w.get("max_planes", 0) > 0:
            self.set_planes(w, first)

    # ---- compute
    def solve(self, opts, n=None, first=0, summaries=None):
        n = n if n is not None else self.B
        if summaries is None:
            summaries = np.zeros(n, abi.SUMMARY)
        _check(lib().gf2_solve(self.h, first, n, C.byref(opts), _p(summaries)))
        return summaries

    def linearize(self, opts, n=None, first=0):
        n = n if n is not None else self.B
        _check(lib().gf2_linearize(self.h, first, n, C.byref(opts)))
        D = lib().gf2_reduced_dim(self.h, C.byref(opts))
        S = np.zeros((n, D, D)); g = np.zeros((n, D)); cost = np.zeros(n)
        _check(lib().gf2_get_reduced_system(self.h, first, n, _p(S), _p(g), _p(cost)))
        return S, g, cost

    def get_states(self, n=None, first=0, out=None):
        n = n if n is not None else self.B
        o = out if out is not None else {}
        o.setdefault("para_pose", np.zeros((n, self.F, 7))); o.setdefault("para_speedbias", np.zeros((n, self.F, 9)))
        o.setdefault("ex_pose", np.zeros((n, 7))); o.setdefault("td", np.zeros(n))
        if self.cfg.use_wheel:
            o.setdefault("ex_pose_wheel", np.zeros((n, 7))); o.setdefault("sxsysw", np.zeros((n, 3))); o.setdefault("td_wheel", np.zeros(n))
        _check(lib().gf2_get_states(self.h, first, n, _p(o["para_pose"]), _p(o["para_speedbias"]), _p(o["ex_pose"]), _p(o["td"]),
                                    _p(o.get("ex_pose_wheel")), _p(o.get("sxsysw")), _p(o.get("td_wheel"))))
        return o

    def get_landmarks(self, n=None, first=0, out=None):
        n = n if n is not None else self.B
        d = out if out is not None else np.zeros((n, self.Lm))
        _check(lib().gf2_get_landmarks(self.h, first, n, _p(d)))
        return d

    @staticmethod
    def comm_unique_id():
        """128-byte ncclUniqueId (rank 0 creates it, the caller broadcasts it)."""
        buf = np.zeros(128, np.uint8)
        _check(lib().gf2_comm_unique_id(_p(buf)))
        return buf

    def comm_init(self, rank, nranks, unique_id):
        uid = np.ascontiguousarray(unique_id, np.uint8)
        _check(lib().gf2_comm_init(self.h, int(rank), int(nranks), _p(uid)))

    def get_trace(self, n=None, first=0):
        n = n if n is not None else self.B
        tr = np.zeros((n, 64, 6))
        _check(lib().gf2_get_trace(self.h, first, n, _p(tr)))
        return tr

    def last_timing(self):
        t = np.zeros(8)
        _check(lib().gf2_last_timing(self.h, _p(t)))
        return {"total_ms": t[0], "linearize_ms": t[1], "solve_ms": t[2], "step_ms": t[3], "launches": int(t[4]), "linearize_launches": int(t[5]), "prepare_ms": t[6], "nccl_ms": t[7]}


def detect_select(idx, val, width, height, max_corners, min_distance):
    """Host half of the detector (no device needed): cv's corner selection over candidates (flat index y*W+x, float32 score)."""
    b = np.asarray(val, np.float32).view(np.uint32).astype(np.uint64)
    ordered = np.where(b & np.uint64(0x80000000), ~b & np.uint64(0xffffffff), b | np.uint64(0x80000000))
    keys = np.ascontiguousarray((ordered << np.uint64(32)) | np.asarray(idx, np.uint64))
    out = np.zeros((max(int(max_corners), 1), 2), np.float32); n = np.zeros(1, np.int32)
    _check(lib().gf2_detect_select(_p(keys), len(keys), int(width), int(height), int(max_corners), C.c_double(min_distance), _p(out), _p(n)))
    return out[:n[0]].copy()


class Tracker:
    """Batched pyramidal LK: cv::calcOpticalFlowPyrLK as FeatureTracker::trackImage calls it (feature_tracker.cpp:122-142)."""

    def __init__(self, width=640, height=480, max_pts=300, max_streams=1, max_level=3, win=21, max_iters=30, eps=0.01,
                 min_eig=1e-4, device=0):
        cfg = abi.TrackerCfg()
        cfg.device = device; cfg.width = width; cfg.height = height; cfg.max_pts = max_pts; cfg.win = win
        cfg.max_level = max_level; cfg.max_iters = max_iters; cfg.max_streams = max_streams; cfg.eps = eps; cfg.min_eig = min_eig
        self.cfg = cfg
        self.h = C.c_void_p()
        _check(lib().gf2_tracker_create(C.byref(cfg), C.byref(self.h)))

    def close(self):
        if self.h:
            lib().gf2_tracker_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _prep(self, prev, cur, prev_pts, n_pts):
        cur = np.ascontiguousarray(cur, np.uint8)
        S = cur.shape[0] if cur.ndim == 3 else 1
        if prev is not None:
            prev = np.ascontiguousarray(prev, np.uint8)
        pts = np.zeros((S, self.cfg.max_pts, 2), np.float32)
        prev_pts = np.asarray(prev_pts, np.float32)
        if prev_pts.ndim == 2:
            prev_pts = prev_pts[None]
        npts = np.zeros(S, np.int32)
        for s in range(S):
            k = prev_pts[s].shape[0] if n_pts is None else int(n_pts[s])
            pts[s, :k] = prev_pts[s][:k]; npts[s] = k
        return S, prev, cur, pts, npts

    def track(self, prev, cur, prev_pts, n_pts=None, init_pts=None, max_level=None, flags=0):
        """prev/cur: [H, W] or [S, H, W] uint8 (prev None = reuse the cached pyramid); prev_pts [n, 2] or [S, n, 2].
        Returns (cur_pts [S, max_pts, 2], status [S, max_pts], err [S, max_pts])."""
        S, prev, cur, pts, npts = self._prep(prev, cur, prev_pts, n_pts)
        out = np.zeros_like(pts)
        if init_pts is not None:
            ip = np.asarray(init_pts, np.float32)
            ip = ip[None] if ip.ndim == 2 else ip
            for s in range(S):
                out[s, :ip[s].shape[0]] = ip[s]
        status = np.zeros((S, self.cfg.max_pts), np.uint8); err = np.zeros((S, self.cfg.max_pts), np.float32)
        ml = self.cfg.max_level if max_level is None else max_level
        _check(lib().gf2_tracker_track(self.h, S, _p(prev), _p(cur), C.c_size_t(self.cfg.width), _p(npts), _p(pts), _p(out), _p(status), _p(err), int(flags), int(ml)))
        return out, status, err

    def track_fb(self, prev, cur, prev_pts, n_pts=None, max_level=None):
        """forward + reverse check fused (trackImage without prediction). Returns (cur_pts, status)."""
        S, prev, cur, pts, npts = self._prep(prev, cur, prev_pts, n_pts)
        out = np.zeros_like(pts); status = np.zeros((S, self.cfg.max_pts), np.uint8)
        ml = self.cfg.max_level if max_level is None else max_level
        _check(lib().gf2_tracker_track_fb(self.h, S, _p(prev), _p(cur), C.c_size_t(self.cfg.width), _p(npts), _p(pts), _p(out), _p(status), int(ml)))
        return out, status

    def track_image(self, prev, cur, prev_pts, predict_pts=None, n_pts=None, flow_back=True, max_level=None):
        """The whole LK stage of trackImage (feature_tracker.cpp:113-153): optional prediction (level-1 LK from predict_pts, per-stream
        fall-back to max_level when fewer than 10 points succeed) and the reverse check. Returns (cur_pts, status)."""
        S, prev, cur, pts, npts = self._prep(prev, cur, prev_pts, n_pts)
        pred = None
        if predict_pts is not None:
            pp = np.asarray(predict_pts, np.float32)
            pp = pp[None] if pp.ndim == 2 else pp
            pred = np.zeros_like(pts)
            for s in range(S):
                pred[s, :pp[s].shape[0]] = pp[s]
        out = np.zeros_like(pts); status = np.zeros((S, self.cfg.max_pts), np.uint8)
        ml = self.cfg.max_level if max_level is None else max_level
        _check(lib().gf2_tracker_track_image(self.h, S, _p(prev), _p(cur), C.c_size_t(self.cfg.width), _p(npts), _p(pts), _p(pred), int(bool(flow_back)),
                                             _p(out), _p(status), int(ml)))
        return out, status

    def _imgs(self, img):
        if img is None:
            return None, None
        img = np.ascontiguousarray(img, np.uint8)
        return (img[None] if img.ndim == 2 else img), None

    def detect(self, img, max_corners, mask=None, quality_level=0.01, min_distance=30.0, n_streams=None):
        """cv::goodFeaturesToTrack(img, max_corners, quality_level, min_distance, mask) as trackImage calls it (feature_tracker.cpp:198).
        img [H, W] or [S, H, W] uint8, or None = the `cur` image of the last track call. Returns a list of [n_s, 2] float32 arrays."""
        im, _ = self._imgs(img)
        S = im.shape[0] if im is not None else (n_streams or 1)
        mk = None
        if mask is not None:
            mk = np.ascontiguousarray(mask, np.uint8)
            mk = mk[None] if mk.ndim == 2 else mk
        mc = np.broadcast_to(np.asarray(max_corners, np.int32), (S,)).copy()
        out = np.zeros((S, self.cfg.max_pts, 2), np.float32); n = np.zeros(S, np.int32)
        _check(lib().gf2_tracker_detect(self.h, S, _p(im), C.c_size_t(self.cfg.width), _p(mk), _p(mc), C.c_double(quality_level), C.c_double(min_distance),
                                        _p(out), _p(n)))
        return [out[s, :n[s]].copy() for s in range(S)]

    def equalize(self, img, clip_limit=40.0, tiles=(8, 8)):
        """cv2.createCLAHE(clip_limit, tiles).apply(img) for [H, W] or [S, H, W] uint8 images -> uint8 [S, H, W]."""
        im, _ = self._imgs(img)
        out = np.zeros_like(im)
        _check(lib().gf2_tracker_equalize(self.h, im.shape[0], _p(im), C.c_size_t(self.cfg.width), C.c_double(clip_limit), int(tiles[0]), int(tiles[1]), _p(out)))
        return out

    def set_equalize(self, clip_limit=40.0, tiles=(8, 8)):
        """Equalise every uploaded image on the device before it is tracked / searched (EQUALIZE of the reference's node)."""
        _check(lib().gf2_tracker_set_equalize(self.h, C.c_double(clip_limit), int(tiles[0]), int(tiles[1])))

    def get_image(self, n_streams=1):
        out = np.zeros((n_streams, self.cfg.height, self.cfg.width), np.uint8)
        _check(lib().gf2_tracker_get_image(self.h, n_streams, _p(out)))
        return out

    def min_eigen_map(self, img):
        """cv::cornerMinEigenVal(img, 3, 3) of [H, W] or [S, H, W] uint8 images -> float32 [S, H, W]."""
        im, _ = self._imgs(img)
        eig = np.zeros((im.shape[0], self.cfg.height, self.cfg.width), np.float32)
        _check(lib().gf2_tracker_min_eigen_map(self.h, im.shape[0], _p(im), C.c_size_t(self.cfg.width), _p(eig)))
        return eig

    def last_timing(self):
        t = np.zeros(8)
        _check(lib().gf2_tracker_last_timing(self.h, _p(t)))
        return {"total_ms": t[0], "pyramid_ms": t[1], "lk_ms": t[2], "launches": int(t[3]), "detect_ms": t[4], "candidates": int(t[5])}


class Lio:
    """LIO factor construction: lidarodom::addSurfCostFactor (searchNeighbors + computeNeighborhoodDistribution + the residual gate,
    LIO/liw/lio/lidarodom.cpp:887-1165) over a snapshot of the voxel map."""

    def __init__(self, max_voxels, max_keypoints, max_points_per_voxel=20, device=0):
        cfg = abi.LioCfg()
        cfg.device = device; cfg.max_voxels = max_voxels; cfg.max_points_per_voxel = max_points_per_voxel; cfg.max_keypoints = max_keypoints
        self.cfg = cfg
        self.h = C.c_void_p()
        _check(lib().gf2_lio_create(C.byref(cfg), C.byref(self.h)))

    def close(self):
        if self.h:
            lib().gf2_lio_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_map(self, keys, n_points, points):
        """keys [n, 3] int16, n_points [n] int32, points [n, max_points_per_voxel, 3] float64 (insertion order)."""
        keys = np.ascontiguousarray(keys, np.int16); n_points = np.ascontiguousarray(n_points, np.int32); points = np.ascontiguousarray(points, np.float64)
        assert points.shape[1] == self.cfg.max_points_per_voxel
        _check(lib().gf2_lio_set_map(self.h, len(keys), _p(keys), _p(n_points), _p(points)))

    def add_points(self, points, size_voxel_map=0.2, min_distance_points=0.05, min_num_points=0):
        """lidarodom::map_incremental: addPointToMap for the points [n, 3] of a scan in order, on the device-resident map."""
        pts = np.ascontiguousarray(points, np.float64)
        _check(lib().gf2_lio_add_points(self.h, len(pts), _p(pts), C.c_double(size_voxel_map), C.c_double(min_distance_points), int(min_num_points)))

    def get_map(self):
        """Snapshot of the device map (keys [n, 3] int16 ascending, n_points [n], points [n, M, 3])."""
        n = C.c_int32(0)
        _check(lib().gf2_lio_map_size(self.h, C.byref(n)))
        keys = np.zeros((n.value, 3), np.int16); npts = np.zeros(n.value, np.int32); pts = np.zeros((n.value, self.cfg.max_points_per_voxel, 3))
        _check(lib().gf2_lio_get_map(self.h, _p(keys), _p(npts), _p(pts)))
        return keys, npts, pts

    def build_factors(self, keypoints, opts, want_neighbors=False):
        """Returns (factors [n] abi.PLANE, alpha [n], neighbors or None, n_neighbors or None)."""
        kp = np.ascontiguousarray(keypoints, abi.LIO_KEYPOINT)
        cap = max(int(opts.max_num_residuals), 1)
        fac = np.zeros(cap, abi.PLANE); alpha = np.zeros(cap); n = C.c_int32(0)
        nb = np.zeros((len(kp), opts.max_number_neighbors, 3)) if want_neighbors else None
        nn = np.zeros(len(kp), np.int32) if want_neighbors else None
        _check(lib().gf2_lio_build_factors(self.h, len(kp), _p(kp), C.byref(opts), _p(fac), _p(alpha), C.byref(n), _p(nb), _p(nn)))
        return fac[:n.value].copy(), alpha[:n.value].copy(), nb, nn

    def last_timing(self):
        t = np.zeros(8)
        _check(lib().gf2_lio_last_timing(self.h, _p(t)))
        return {"total_ms": t[0], "kernel_ms": t[1], "voxels": int(t[2]), "keypoints": int(t[3]), "new_voxels": int(t[4])}
