"""ctypes access to the flat C test API of ground-fusion2_b200/libgf2_host.so (the C++ Estimator / FeatureTracker mirror)."""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(os.path.join(ROOT, "ground-fusion2_b200", "libgf2_host.so"))
        for f in ("gf2h_estimator_create", "gf2h_tracker_create"):
            getattr(_lib, f).restype = C.c_void_p
        _lib.gf2h_last_error.restype = C.c_char_p; _lib.gf2h_tracker_last_error.restype = C.c_char_p
    return _lib


def p(a):
    return None if a is None else np.ascontiguousarray(a).ctypes.data_as(C.c_void_p)


DETECTOR = C.CFUNCTYPE(C.c_int, C.POINTER(C.c_uint8), C.c_int, C.c_int, C.POINTER(C.c_uint8), C.c_int, C.c_int, C.POINTER(C.c_float), C.c_void_p)


def frame_states(P, R, V, Ba, Bg):
    s = np.zeros((11, 21))
    s[:, 0:3] = P; s[:, 3:12] = R.reshape(11, 9); s[:, 12:15] = V; s[:, 15:18] = Ba; s[:, 18:21] = Bg
    return s
