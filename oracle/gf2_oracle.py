"""TEST INFRASTRUCTURE — ctypes binding of oracle/_build/libgf2_oracle.so (the CPU restatement of the reference
path). Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this."""
import ctypes as C
import os
import subprocess
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_build", "libgf2_oracle.so")


def build():
    subprocess.check_call(["make", "-s", "-C", HERE])


def _load():
    if not os.path.exists(LIB_PATH):
        build()
    return C.CDLL(LIB_PATH)


lib = _load()


class Window(C.Structure):
    _fields_ = [("n_frames", C.c_int32), ("n_landmarks", C.c_int32), ("n_planes", C.c_int32), ("prior_rows", C.c_int32),
                ("prior_nblocks", C.c_int32), ("use_wheel", C.c_int32), ("prior_stride", C.c_int32), ("pad_", C.c_int32)] + [(n, C.c_void_p) for n in (
                    "para_pose", "para_speedbias", "ex_pose", "td", "ex_pose_wheel", "sxsysw", "td_wheel", "inv_depth",
                    "start_frame", "track_len", "fixed", "obs", "frame_td", "imu", "wheel", "prior_J0", "prior_r0",
                    "prior_blocks", "planes", "plane_alpha")]


class Batch(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("n_windows", "n_frames", "max_landmarks", "max_obs", "max_planes", "use_wheel", "prior_stride", "pad_")] + \
               [(n, C.c_void_p) for n in ("para_pose", "para_speedbias", "ex_pose", "td", "ex_pose_wheel", "sxsysw", "td_wheel",
                                          "inv_depth", "n_landmarks", "start_frame", "track_len", "fixed", "obs", "frame_td",
                                          "imu", "wheel", "prior_rows", "prior_J0", "prior_r0", "prior_nblocks", "prior_blocks",
                                          "n_planes", "planes", "plane_alpha")]


assert lib.gf2o_sizeof(4) == C.sizeof(Window) and lib.gf2o_sizeof(5) == C.sizeof(Batch)


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def make_batch(w, keep):
    """Batch struct over the arrays of a synth window dict `w` (arrays must stay alive: stored in `keep`)."""
    b = Batch()
    n = w["para_pose"].shape[0]
    b.n_windows = n; b.n_frames = w["n_frames"]; b.max_landmarks = w["max_landmarks"]; b.max_obs = w["max_obs"]
    b.max_planes = w.get("max_planes", 0); b.use_wheel = 1 if w.get("use_wheel") else 0
    b.prior_stride = int(w["prior_J0"].shape[1]) if w.get("prior_J0") is not None else 96
    for name in ("para_pose", "para_speedbias", "ex_pose", "td", "ex_pose_wheel", "sxsysw", "td_wheel", "inv_depth",
                 "n_landmarks", "start_frame", "track_len", "fixed", "obs", "frame_td", "imu", "wheel", "prior_rows",
                 "prior_J0", "prior_r0", "prior_nblocks", "prior_blocks", "n_planes", "planes", "plane_alpha"):
        a = w.get(name)
        if a is not None:
            a = np.ascontiguousarray(a); w[name] = a; keep.append(a)
        setattr(b, name, _p(a))
    return b


def solve_batch(w, opts, n_threads=1):
    """Solve every window of `w` in place (states, inverse depths); returns the summaries array."""
    from gf2_loader import load
    abi = load().abi
    keep = []
    b = make_batch(w, keep)
    n = b.n_windows
    summ = np.zeros(n, abi.SUMMARY)
    lib.gf2o_solve_batch(C.byref(b), C.byref(opts), _p(summ), int(n_threads))
    return summ


def solve_window_trace(w, i, opts):
    from gf2_loader import load
    abi = load().abi
    keep = []
    b = make_batch(w, keep)
    win = Window()
    lib.gf2o_batch_window(C.byref(b), int(i), C.byref(win))
    summ = np.zeros(1, abi.SUMMARY)
    trace = np.zeros((opts.max_iterations, 6))
    lib.gf2o_solve_window(C.byref(win), C.byref(opts), _p(summ), _p(trace))
    return summ[0], trace


def linearize_window(w, i, opts, max_dim=256):
    keep = []
    b = make_batch(w, keep)
    win = Window()
    lib.gf2o_batch_window(C.byref(b), int(i), C.byref(win))
    S = np.zeros((max_dim, max_dim)); g = np.zeros(max_dim); cost = C.c_double(0)
    L = int(w["n_landmarks"][i])
    ete = np.zeros(L + 1); etr = np.zeros(L + 1)
    D = lib.gf2o_linearize_window(C.byref(win), C.byref(opts), max_dim, _p(S), _p(g), C.byref(cost), _p(ete), _p(etr))
    assert D > 0
    S = S.reshape(-1)[:D * D].reshape(D, D).copy()
    return S, g[:D].copy(), cost.value, ete, etr


def factor_eval(kind, consts, params, extra=None, want_jac=True):
    sizes = {0: (2, [7, 7, 7, 1, 1]), 1: (15, [7, 9, 7, 9]), 2: (6, [7, 7, 7, 1, 1, 1, 1]), 3: (1, [3, 4]), 4: (1, [3, 4, 3, 4]), 5: (1, [7]), 6: (1, [7, 7])}[kind]
    nres, blocks = sizes
    params = np.ascontiguousarray(params, dtype=np.float64)
    assert params.size == sum(blocks)
    res = np.zeros(nres); jac = np.zeros(nres * sum(blocks))
    consts = np.ascontiguousarray(consts)
    ex = np.ascontiguousarray(extra if extra is not None else [0.0], dtype=np.float64)
    lib.gf2o_factor_eval(int(kind), _p(consts), _p(ex), _p(params), _p(res), _p(jac) if want_jac else None)
    out = []; o = 0
    for s in blocks:
        out.append(jac[o:o + nres * s].reshape(nres, s).copy()); o += nres * s
    return res, out


def imu_preintegrate(w):
    """Fill w['imu'] ([n][F-1] IMU_PREINT) from the raw samples with the restated IntegrationBase."""
    from gf2_loader import load
    abi = load().abi
    n, Fm1 = w["imu_n"].shape
    rec = np.zeros((n, Fm1), abi.IMU_PREINT)
    noise = np.ascontiguousarray(w["imu_noise"], dtype=np.float64)
    for i in range(n):
        for k in range(Fm1):
            smp = np.ascontiguousarray(w["imu_samples"][i, k]); first = np.ascontiguousarray(w["imu_first"][i, k]); lb = np.ascontiguousarray(w["imu_lin_bias"][i, k])
            one = np.zeros(1, abi.IMU_PREINT)
            lib.gf2o_imu_preintegrate(_p(smp), int(w["imu_n"][i, k]), _p(first), _p(lb), _p(noise), _p(one))
            rec[i, k] = one[0]
    w["imu"] = rec
    return rec


def wheel_preintegrate(w):
    from gf2_loader import load
    abi = load().abi
    n, Fm1 = w["wheel_n"].shape
    rec = np.zeros((n, Fm1), abi.WHEEL_PREINT)
    noise = np.ascontiguousarray(w["wheel_noise"], dtype=np.float64)
    for i in range(n):
        for k in range(Fm1):
            smp = np.ascontiguousarray(w["wheel_samples"][i, k]); first = np.ascontiguousarray(w["wheel_first"][i, k]); lin = np.ascontiguousarray(w["wheel_lin"][i, k])
            one = np.zeros(1, abi.WHEEL_PREINT)
            lib.gf2o_wheel_preintegrate(_p(smp), int(w["wheel_n"][i, k]), _p(first), _p(lin), _p(noise), _p(one))
            rec[i, k] = one[0]
    w["wheel"] = rec
    return rec


def marginalize_window(w, i, opts, mode=0, P=96):
    """Restated marginalization (estimator.cpp:3394-3690) of window i at its current states. Returns a dict
    {status, n, m, J0 [n][n], r0 [n], blocks [nb] PRIOR_BLOCK} in the indexing of the window after slideWindow."""
    from gf2_loader import load
    abi = load().abi
    keep = []
    b = make_batch(w, keep)
    win = Window()
    lib.gf2o_batch_window(C.byref(b), int(i), C.byref(win))
    J0 = np.zeros((P, P)); r0 = np.zeros(P); nb = C.c_int32(0); m = C.c_int32(0)
    blocks = np.zeros(2 * w["n_frames"] + 8, abi.PRIOR_BLOCK)
    n = lib.gf2o_marginalize_window(C.byref(win), C.byref(opts), int(mode), int(P), _p(J0), _p(r0), C.byref(nb), _p(blocks), C.byref(m))
    if n < 0:
        return {"status": n, "n": 0, "m": m.value}
    return {"status": 0, "n": n, "m": m.value, "J0": J0[:n, :n].copy(), "r0": r0[:n].copy(), "blocks": blocks[:nb.value].copy()}


BLOCK_LOCAL = {0: 6, 1: 9, 2: 6, 3: 1, 4: 6, 5: 1, 6: 1, 7: 1, 8: 1}


def prior_information(prior, n_frames=11):
    """Order-independent form of a prior: (H, g, x0) with H = J0^T J0 and g = J0^T r0 scattered into the canonical tangent
    layout [pose f (6) | speed-bias f (9)] per frame, then ex-pose 6, td 1, ex-wheel 6, sx sy sw, td-wheel; x0 maps
    (kind, index) -> linearisation point."""
    F = n_frames
    base = {2: 15 * F, 3: 15 * F + 6, 4: 15 * F + 7, 5: 15 * F + 13, 6: 15 * F + 14, 7: 15 * F + 15, 8: 15 * F + 16}
    T = 15 * F + 17
    cols = np.full(prior["n"], -1)
    x0 = {}
    for b in prior["blocks"]:
        kind, index, off = int(b["kind"]), int(b["index"]), int(b["offset"])
        start = 15 * index + (0 if kind == 0 else 6) if kind in (0, 1) else base[kind]
        cols[off:off + BLOCK_LOCAL[kind]] = start + np.arange(BLOCK_LOCAL[kind])
        x0[(kind, index)] = np.array(b["x0"])
    assert (cols >= 0).all()
    H = np.zeros((T, T)); g = np.zeros(T)
    Hs = prior["J0"].T @ prior["J0"]; gs = prior["J0"].T @ prior["r0"]
    H[np.ix_(cols, cols)] = Hs; g[cols] = gs
    return H, g, x0


def lio_build_factors(scene, opts, want_neighbors=False, timing=None):
    """lidarodom::addSurfCostFactor restated (gf2o_lio.cpp) on a synth.lio_scene dict. Returns (factors, alpha, neighbors, n_neighbors)."""
    from gf2_loader import load
    abi = load().abi
    kp = np.ascontiguousarray(scene["keypoints"]); keys = np.ascontiguousarray(scene["keys"], np.int16)
    npts = np.ascontiguousarray(scene["n_points"], np.int32); pts = np.ascontiguousarray(scene["points"], np.float64)
    cap = max(int(opts.max_num_residuals), 1)
    fac = np.zeros(cap, abi.PLANE); alpha = np.zeros(cap)
    nb = np.zeros((len(kp), opts.max_number_neighbors, 3)) if want_neighbors else None
    nn = np.full(len(kp), -1, np.int32) if want_neighbors else None
    secs = C.c_double(0.0)
    n = lib.gf2o_lio_build_factors(len(keys), _p(keys), _p(npts), _p(pts), int(pts.shape[1]), len(kp), _p(kp), C.byref(opts), _p(fac), _p(alpha), _p(nb), _p(nn), C.byref(secs))
    if timing is not None:
        timing["loop_seconds"] = secs.value   # the keypoint loop alone: the voxel map is rebuilt per call and excluded
    if n < 0:
        raise RuntimeError("a2D is NaN (the reference throws)")
    return fac[:n].copy(), alpha[:n].copy(), nb, nn
