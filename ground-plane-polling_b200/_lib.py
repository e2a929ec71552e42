"""ctypes binding of libgpp.so (C ABI: include/gpp.h).  There is no CPU fallback: if the CUDA library is
missing or no sm_100 device is usable, every entry point raises."""
import ctypes
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('GPP_LIB_PATH') or os.path.join(HERE, 'libgpp.so')   # override: kernel experiments only

GPP_MODE_EXACT, GPP_MODE_FAST, GPP_MODE_F64, GPP_MODE_VERIFIED = 0, 1, 2, 3
MODES = {'exact': GPP_MODE_EXACT, 'fast': GPP_MODE_FAST, 'f64': GPP_MODE_F64, 'verified': GPP_MODE_VERIFIED}

c_float_p = ctypes.POINTER(ctypes.c_float)
c_double_p = ctypes.POINTER(ctypes.c_double)
c_int32_p = ctypes.POINTER(ctypes.c_int32)
c_int64_p = ctypes.POINTER(ctypes.c_int64)
c_void_p = ctypes.c_void_p
c_int = ctypes.c_int

# name -> (restype, argtypes); mirrors include/gpp.h and include/gpp_debug.h one to one (tests/test_capi.py checks it)
SIGNATURES = {
    'gpp_version': (c_int, []),
    'gpp_device_count': (c_int, []),
    'gpp_last_error': (ctypes.c_char_p, []),
    'gpp_create': (c_int, [c_int, ctypes.POINTER(c_void_p)]),
    'gpp_destroy': (c_int, [c_void_p]),
    'gpp_set_planes': (c_int, [c_void_p, c_void_p, c_int]),
    'gpp_set_planes_raw': (c_int, [c_void_p, c_void_p, c_int, c_int, c_int]),
    'gpp_set_planes_device': (c_int, [c_void_p, c_void_p, c_int, c_void_p]),
    'gpp_num_planes': (c_int, [c_void_p]),
    'gpp_get_normalised_planes': (c_int, [c_void_p, c_void_p]),
    'gpp_fit_host': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int,
                             c_void_p, c_void_p, c_void_p, c_void_p, c_int]),
    'gpp_fit_host_multi': (c_int, [ctypes.POINTER(c_void_p), c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int,
                                   c_void_p, c_void_p, c_void_p, c_void_p, c_int]),
    'gpp_fit_planes_host': (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int,
                                    c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int]),
    'gpp_fit_host_multi_planes': (c_int, [ctypes.POINTER(c_void_p), c_int, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p,
                                          c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int]),
    'gpp_fit_device': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int,
                               c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p]),
    'gpp_fit_pose_host': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p,
                                  c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int]),
    'gpp_fit_pose_device': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p,
                                    c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p]),
    'gpp_fit_host_f64': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int,
                                 c_void_p, c_void_p, c_void_p, c_void_p]),
    'gpp_fit_device_f64': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int,
                                   c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    'gpp_pose_host': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, ctypes.c_long,
                              c_void_p, c_void_p, c_void_p]),
    'gpp_pose_device': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, ctypes.c_long,
                                c_void_p, c_void_p, c_void_p, c_void_p]),
    'gpp_kitti_host': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, ctypes.c_long, c_void_p]),
    'gpp_kitti_device': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, ctypes.c_long, c_void_p, c_void_p]),
    'gpp_decode_host': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p,
                                c_void_p, c_void_p]),
    'gpp_decode_device': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p,
                                  c_void_p, c_void_p, c_void_p]),
    'gpp_filter_host': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, ctypes.c_float, ctypes.c_float,
                                c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    'gpp_filter_device': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, ctypes.c_float, ctypes.c_float,
                                  c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    'gpp_last_kernel_ms': (c_int, [c_void_p, c_float_p]),
    'gpp_launch_count': (ctypes.c_int64, [c_void_p]),
    'gpp_microbench': (c_int, [c_void_p, c_int, c_double_p, c_float_p, c_double_p]),
    'gpp_debug_set_schedule': (c_int, [c_void_p, c_int, c_int]),
    'gpp_debug_scan_order': (c_int, [c_void_p, c_int, c_void_p]),
    'gpp_audit_set': (c_int, [c_void_p, c_int]),
    'gpp_audit_counts': (c_int, [c_void_p, c_int64_p, c_int64_p]),
    'gpp_debug_scores': (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p, c_void_p,
                                 c_void_p]),
}

_LIB = None


class GppError(RuntimeError):
    pass


def load():
    """Load libgpp.so (never builds, never falls back)."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise GppError('libgpp.so not found at %s -- build it with `python -c "import __graft_entry__ as g; '
                           'g.build()"` (nvcc, sm_100a); there is no CPU fallback' % LIB_PATH)
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _LIB = lib
    return _LIB


def check(rc, what):
    if rc != 0:
        msg = load().gpp_last_error()
        msg = msg.decode('utf-8', 'replace') if msg else ''
        if rc == 1:
            raise ValueError('%s: %s' % (what, msg))
        raise GppError('%s failed (code %d): %s' % (what, rc, msg))


def ptr(a):
    """Raw data address of a C-contiguous numpy array (or None) for a ``c_void_p`` argument.  (The plain integer:
    ``ndarray.ctypes.data_as`` costs 2.6 us per array, which adds up to a fifth of a single-image call.)"""
    if a is None:
        return None
    assert isinstance(a, np.ndarray) and a.flags['C_CONTIGUOUS']
    return a.ctypes.data
