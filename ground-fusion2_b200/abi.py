"""numpy mirrors of the C records of include/gf2_abi.h (aligned structured dtypes: arrays of these can be handed to
the C ABI as-is) and ctypes mirrors of the option/summary structs."""
import ctypes as C
import numpy as np

MAX_FRAMES = 11
MAX_LANDMARKS = 1000
MAX_PRIOR_DIM = 96

OBS = np.dtype([("x", "f4"), ("y", "f4"), ("vx", "f4"), ("vy", "f4")], align=True)
IMU_SAMPLE = np.dtype([("dt", "f8"), ("acc", "f8", 3), ("gyr", "f8", 3)], align=True)
WHEEL_SAMPLE = np.dtype([("dt", "f8"), ("vel", "f8", 3), ("gyr", "f8", 3)], align=True)
IMU_PREINT = np.dtype([("sum_dt", "f8"), ("delta_p", "f8", 3), ("delta_q", "f8", 4), ("delta_v", "f8", 3),
                       ("lin_ba", "f8", 3), ("lin_bg", "f8", 3), ("jacobian", "f8", 225), ("covariance", "f8", 225),
                       ("valid", "i4"), ("pad_", "i4")], align=True)
WHEEL_PREINT = np.dtype([("sum_dt", "f8"), ("delta_p", "f8", 3), ("delta_q", "f8", 4),
                         ("lin_sx", "f8"), ("lin_sy", "f8"), ("lin_sw", "f8"), ("lin_td", "f8"),
                         ("lin_vel", "f8", 3), ("lin_gyr", "f8", 3), ("vel_1", "f8", 3), ("gyr_1", "f8", 3),
                         ("jacobian", "f8", 18), ("covariance", "f8", 36), ("valid", "i4"), ("pad_", "i4")], align=True)
PLANE = np.dtype([("p_body", "f8", 3), ("normal", "f8", 3), ("offset", "f8"), ("weight", "f8"),
                  ("frame", "i4"), ("ct", "i4")], align=True)
PRIOR_BLOCK = np.dtype([("kind", "i4"), ("index", "i4"), ("offset", "i4"), ("pad_", "i4"), ("x0", "f8", 9)], align=True)

BLK_POSE, BLK_SPEEDBIAS, BLK_EX_POSE, BLK_TD, BLK_EX_WHEEL, BLK_SX, BLK_SY, BLK_SW, BLK_TD_WHEEL = range(9)
CONST_EX_POSE, CONST_TD, CONST_EX_WHEEL, CONST_WHEEL_INTRINSIC, CONST_TD_WHEEL = 1, 2, 4, 8, 16
CONST_ALL_CALIB = 31
TERM_NAMES = ["no_convergence", "function_tol", "gradient_tol", "parameter_tol", "min_radius", "failure"]
LK_USE_INITIAL_FLOW = 4


SWEEP_AUTO, SWEEP_BATCH, SWEEP_WINDOW = 0, 1, 2   # gf2_solver_cfg.sweep


class SolverCfg(C.Structure):
    _fields_ = [("device", C.c_int32), ("max_windows", C.c_int32), ("n_frames", C.c_int32),
                ("max_landmarks", C.c_int32), ("max_obs", C.c_int32), ("max_planes", C.c_int32),
                ("max_imu_samples", C.c_int32), ("max_wheel_samples", C.c_int32), ("use_wheel", C.c_int32),
                ("max_prior_rows", C.c_int32), ("sweep", C.c_int32), ("reserved_", C.c_int32 * 5)]


class SolveOpts(C.Structure):
    _fields_ = [("max_iterations", C.c_int32), ("const_mask", C.c_uint32), ("huber_delta", C.c_double),
                ("sqrt_info_px", C.c_double), ("g_norm", C.c_double), ("lidar_sqrt_info", C.c_double),
                ("max_time_s", C.c_double), ("initial_radius", C.c_double), ("function_tolerance", C.c_double),
                ("gradient_tolerance", C.c_double), ("parameter_tolerance", C.c_double), ("wheel_ext_const_components", C.c_uint32), ("marg_eig", C.c_uint32),
                ("reserved_", C.c_double * 5)]


def default_opts(max_iterations=8, const_mask=CONST_ALL_CALIB, g_norm=9.7944, lidar_sqrt_info=31.622776601683793):
    """NUM_ITERATIONS = 8, Huber(1.0), sqrt_info = FOCAL_LENGTH/1.5 = 400, g_norm of m3dgr.yaml:117,
    lidar sqrt_info = sqrt(1/0.001) (LIO/liw/lio/lidarodom.cpp:10,13)."""
    o = SolveOpts()
    o.max_iterations = max_iterations
    o.const_mask = const_mask
    o.huber_delta = 1.0
    o.sqrt_info_px = 400.0
    o.g_norm = g_norm
    o.lidar_sqrt_info = lidar_sqrt_info
    return o


SUMMARY = np.dtype([("initial_cost", "f8"), ("final_cost", "f8"), ("iterations", "i4"), ("successful_steps", "i4"),
                    ("termination", "i4"), ("pad_", "i4")], align=True)


LIO_KEYPOINT = np.dtype([("raw_point", "f8", 3), ("point", "f8", 3), ("alpha_time", "f8")], align=True)
ICP_CT_POINT_TO_PLANE, ICP_POINT_TO_PLANE = 0, 1


class LioCfg(C.Structure):
    _fields_ = [("device", C.c_int32), ("max_voxels", C.c_int32), ("max_points_per_voxel", C.c_int32), ("max_keypoints", C.c_int32)]


class LioOpts(C.Structure):
    _fields_ = [("size_voxel_map", C.c_double), ("max_dist_to_plane_icp", C.c_double), ("power_planarity", C.c_double),
                ("weight_alpha", C.c_double), ("weight_neighborhood", C.c_double), ("nb_voxels_visited", C.c_int32),
                ("threshold_voxel_capacity", C.c_int32), ("max_number_neighbors", C.c_int32), ("min_number_neighbors", C.c_int32),
                ("num_closest_neighbors", C.c_int32), ("max_num_residuals", C.c_int32), ("icp_model", C.c_int32), ("pad_", C.c_int32),
                ("translation_begin", C.c_double * 3), ("rotation", C.c_double * 4), ("translation", C.c_double * 3),
                ("R_IL", C.c_double * 9), ("t_IL", C.c_double * 3)]


def default_lio_opts(**kw):
    """odometry options of LIO/config/m3dgr.yaml (size_voxel_map 0.2, 20 neighbours, max_dist_to_plane_icp 0.3, power_planarity 2,
    weights 0.9 / 0.1, voxel_neighborhood 1, max_num_residuals 2000, CT_POINT_TO_PLANE), identity frame state."""
    o = LioOpts()
    o.size_voxel_map = 0.2; o.max_dist_to_plane_icp = 0.3; o.power_planarity = 2.0; o.weight_alpha = 0.9; o.weight_neighborhood = 0.1
    o.nb_voxels_visited = 1; o.threshold_voxel_capacity = 1; o.max_number_neighbors = 20; o.min_number_neighbors = 20
    o.num_closest_neighbors = 1; o.max_num_residuals = 2000; o.icp_model = ICP_CT_POINT_TO_PLANE
    o.rotation[3] = 1.0
    for i in (0, 4, 8):
        o.R_IL[i] = 1.0
    for k, v in kw.items():
        if isinstance(v, (list, tuple, np.ndarray)):
            for i, x in enumerate(np.asarray(v, np.float64).ravel()):
                getattr(o, k)[i] = float(x)
        else:
            setattr(o, k, v)
    return o


class TrackerCfg(C.Structure):
    _fields_ = [("device", C.c_int32), ("width", C.c_int32), ("height", C.c_int32), ("max_pts", C.c_int32),
                ("win", C.c_int32), ("max_level", C.c_int32), ("max_iters", C.c_int32), ("max_streams", C.c_int32),
                ("eps", C.c_double), ("min_eig", C.c_double)]


def ptr(a, ctype=C.c_void_p):
    """pointer to a C-contiguous numpy array (or None)"""
    if a is None:
        return None
    assert a.flags["C_CONTIGUOUS"], "array must be C-contiguous"
    return a.ctypes.data_as(ctype)

MARGIN_OLD, MARGIN_SECOND_NEW = 0, 1
MARG_INVALID, MARG_UNCHANGED, MARG_UNSUPPORTED, MARG_DEGENERATE, MARG_TOO_LARGE = -1, -2, -3, -4, -5
