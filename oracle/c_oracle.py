"""
ORACLE (test infrastructure, NOT product code) -- ctypes front end of oracle/libgpp_oracle.so
(oracle/gpp_oracle.c), the multi-threaded C restatement of
/root/reference/keras_retinanet_3D/layers/fit_road_planes.py:49-139.  Same call surface as
``oracle.fit_road_planes_ref.fit_road_planes_ref``.  Build with ``make -C oracle``.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build():
    subprocess.check_call(['make', '-s', '-C', _HERE])


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, 'libgpp_oracle.so')
        if not os.path.exists(path):
            build()
        _LIB = ctypes.CDLL(path)
        fp = ctypes.POINTER(ctypes.c_float)
        dp = ctypes.POINTER(ctypes.c_double)
        ip = ctypes.POINTER(ctypes.c_int32)
        lp = ctypes.POINTER(ctypes.c_int64)
        common = [fp, fp, ip, fp, fp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int]
        _LIB.gpp_oracle_fit_f32.argtypes = common + [fp, fp, fp, lp, ctypes.c_int]
        _LIB.gpp_oracle_fit_f64.argtypes = common + [dp, dp, dp, lp, ctypes.c_int]
        _LIB.gpp_oracle_fit_f32.restype = ctypes.c_int
        _LIB.gpp_oracle_fit_f64.restype = ctypes.c_int
        _LIB.gpp_oracle_max_threads.restype = ctypes.c_int
    return _LIB


def max_threads():
    return int(lib().gpp_oracle_max_threads())


def fit_road_planes_c(boxes, dimensions, orientations, P_inv, planes, dtype=np.float32, return_index=False,
                      nthreads=0):
    """C oracle; inputs are cast to float32 first (the Keras feed), outputs are ``dtype`` (f32 or f64)."""
    L = lib()
    f32 = np.float32
    boxes = np.ascontiguousarray(boxes, dtype=f32)
    B, D = boxes.shape[:2]
    dims = np.ascontiguousarray(dimensions, dtype=f32)
    orient = np.ascontiguousarray(orientations, dtype=np.int32)
    P_inv = np.ascontiguousarray(P_inv, dtype=f32)
    planes = np.ascontiguousarray(planes, dtype=f32)
    if planes.ndim == 2:
        planes = planes[None]
    per_image = 0 if planes.shape[0] == 1 else 1
    if per_image:
        assert planes.shape[0] == B
    N = planes.shape[1]
    out_t = np.dtype(dtype)
    kp = np.empty((B, D, 4, 3), out_t)
    kpl = np.empty((B, D, 1, 4), out_t)
    res = np.empty((B, D), out_t)
    best = np.empty((B, D), np.int64)
    if out_t == np.float32:
        fn, cp = L.gpp_oracle_fit_f32, ctypes.POINTER(ctypes.c_float)
    else:
        fn, cp = L.gpp_oracle_fit_f64, ctypes.POINTER(ctypes.c_double)
    fp = ctypes.POINTER(ctypes.c_float)
    rc = fn(boxes.ctypes.data_as(fp), dims.ctypes.data_as(fp),
            orient.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), P_inv.ctypes.data_as(fp),
            planes.ctypes.data_as(fp), per_image, B, D, N,
            kp.ctypes.data_as(cp), kpl.ctypes.data_as(cp), res.ctypes.data_as(cp),
            best.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)), int(nthreads))
    if rc != 0:
        raise RuntimeError('gpp_oracle_fit failed with code %d' % rc)
    out = [kp, kpl, res]
    if return_index:
        out.append(best)
    return out
