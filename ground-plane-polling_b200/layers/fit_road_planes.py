"""
Ground-plane polling -- host-side mirror of the reference's operator interface.

Same names, argument order, shapes and return list as
/root/reference/keras_retinanet_3D/layers/fit_road_planes.py: the free function ``fit_road_planes(boxes,
dimensions, orientations, P_inv, planes)`` (:49-139) and the layer ``FitRoadPlanes`` (:142-186, used at
models/retinanet.py:416 and registered as a custom object at models/__init__.py:15).  The TensorFlow graph
is replaced by one hand-written CUDA kernel in libgpp.so (sm_100a) reached through the C ABI of
include/gpp.h; numpy in / numpy out here, device tensors through ``fit_road_planes_torch`` /
``fit_road_planes_dlpack``.  There is no TensorFlow, no Triton and no CPU fallback: without the CUDA
library or a B200 these functions raise.
"""
import ctypes
import os
import threading

import numpy as np

from .. import _lib

__all__ = ['PlanePoller', 'get_poller', 'fit_road_planes', 'fit_road_planes_torch', 'fit_road_planes_dlpack',
           'FitRoadPlanes', 'DEFAULT_MODE']

# 'verified' (default): FAST arithmetic as a filter + EXACT re-evaluation of everything that could win -> the same
# results as 'exact' (bit-identical to the oracle's canonical fp32 arithmetic) at about twice its speed;
# 'exact': every hypothesis in the canonical arithmetic; 'fast': FMA/MUFU search only (near-ties may differ);
# 'f64': FP64 verify mode.
DEFAULT_MODE = os.environ.get('GPP_MODE', 'verified')


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)      # the cast Keras applies at feed (run_network.py:105)


class PlanePoller(object):
    """One libgpp context: a CUDA device, its resident (normalised) plane database and staging buffers."""

    def __init__(self, device=0):
        self._lib = _lib.load()
        h = ctypes.c_void_p()
        _lib.check(self._lib.gpp_create(int(device), ctypes.byref(h)), 'gpp_create')
        self._h = h
        self.device = int(device)
        self._dev_planes = None      # (tensor, version) of the last device-side database, held so that it stays alive

    def close(self):
        if getattr(self, '_h', None):
            self._lib.gpp_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ plane database
    def set_planes(self, planes):
        """Upload a raw (N, 4) road-plane database (any float dtype / memory order, e.g. the float64
        Fortran-ordered array scipy.io.loadmat returns, run_network.py:75).  Idempotent on equal content."""
        p = np.asarray(planes)
        if p.ndim != 2 or p.shape[1] != 4 or p.shape[0] < 1:
            raise ValueError('planes must have shape (N, 4) with N >= 1, got %r' % (p.shape,))
        # the callers' array goes to the library as it is (float64 / Fortran order from loadmat included): the cast
        # Keras applies at feed and the "same database as last time?" test (an exact byte comparison) happen there
        if p.dtype not in (np.float32, np.float64) or not (p.flags['C_CONTIGUOUS'] or p.flags['F_CONTIGUOUS']):
            p = _f32(p)
        order = 0 if p.flags['C_CONTIGUOUS'] else 1
        rc = self._lib.gpp_set_planes_raw(self._h, p.ctypes.data, p.shape[0],
                                          1 if p.dtype == np.float64 else 0, order)
        _lib.check(rc, 'gpp_set_planes_raw')
        self._dev_planes = None

    @property
    def num_planes(self):
        return int(self._lib.gpp_num_planes(self._h))

    def normalised_planes(self):
        out = np.empty((self.num_planes, 4), np.float32)
        _lib.check(self._lib.gpp_get_normalised_planes(self._h, _lib.ptr(out)), 'gpp_get_normalised_planes')
        return out

    # ------------------------------------------------------------------ numpy entry
    def fit(self, boxes, dimensions, orientations, P_inv, mode=None, return_index=False, out=None, return_pose=False,
            return_kitti=False, planes=None):
        """fit_road_planes against the resident database, or against ``planes`` ((N, 4), as ``set_planes`` takes it: the
        comparison with the resident database then overlaps the kernel of a small call, gpp_fit_planes_host).  Returns [keypoints (B, D, 4, 3),
        keyplanes (B, D, 1, 4), residuals (B, D)] float32 (float64 in mode 'f64') [+ best index int64]
        [+ locations, angles, dimensions (B, D, 3) with ``return_pose``] [+ KITTI records (B, D, 4) with ``return_kitti``:
        both computed in the polling kernel's epilogue, one launch for everything].
        ``out`` may hold preallocated C-contiguous result arrays (e.g. pinned host memory) to write into."""
        mode = DEFAULT_MODE if mode is None else mode
        if mode not in _lib.MODES:
            raise ValueError('unknown mode %r (expected one of %s)' % (mode, sorted(_lib.MODES)))
        if planes is not None:
            planes = np.asarray(planes)
            if planes.ndim != 2 or planes.shape[1] != 4 or planes.shape[0] < 1:
                raise ValueError('planes must have shape (N, 4) with N >= 1, got %r' % (planes.shape,))
            if planes.dtype not in (np.float32, np.float64) or not (planes.flags['C_CONTIGUOUS'] or planes.flags['F_CONTIGUOUS']):
                planes = _f32(planes)
            if mode == 'f64':
                self.set_planes(planes)
                planes = None
        boxes = _f32(boxes)
        if boxes.ndim != 3 or boxes.shape[2] != 12:
            raise ValueError('boxes must have shape (B, D, 12), got %r' % (boxes.shape,))
        B, D = boxes.shape[:2]
        dimensions = _f32(dimensions)
        if dimensions.shape != (B, D, 3):
            raise ValueError('dimensions must have shape (%d, %d, 3), got %r' % (B, D, dimensions.shape))
        orientations = np.ascontiguousarray(orientations, dtype=np.int32)
        if orientations.shape != (B, D):
            raise ValueError('orientations must have shape (%d, %d), got %r' % (B, D, orientations.shape))
        P_inv = _f32(P_inv)
        if P_inv.shape != (B, 4, 3):
            raise ValueError('P_inv must have shape (%d, 4, 3), got %r' % (B, P_inv.shape))
        out_t = np.float64 if mode == 'f64' else np.float32
        if out is not None:
            want = [((B, D, 4, 3), out_t), ((B, D, 1, 4), out_t), ((B, D), out_t)] + \
                   ([((B, D), np.int64)] if return_index else [])
            if len(out) != len(want):
                raise ValueError('out must hold %d arrays' % len(want))
            for a, (shape, dt) in zip(out, want):
                if not isinstance(a, np.ndarray) or a.shape != shape or a.dtype != dt or not a.flags['C_CONTIGUOUS']:
                    raise ValueError('out array must be C-contiguous %s %r' % (np.dtype(dt).name, shape))
            keypoints, keyplanes, residuals = out[0], out[1], out[2]
            best = out[3] if return_index else None
        else:
            keypoints = np.empty((B, D, 4, 3), out_t)
            keyplanes = np.empty((B, D, 1, 4), out_t)
            residuals = np.empty((B, D), out_t)
            best = np.empty((B, D), np.int64) if return_index else None
        pose = None
        if return_pose or return_kitti:
            if mode == 'f64':
                raise ValueError("return_pose / return_kitti are computed in float32 like the reference's driver: "
                                 "not available in mode 'f64'")
            pose = [np.empty((B, D, 3), np.float32) for _ in range(3)] + \
                   ([np.empty((B, D, 4), np.float32)] if return_kitti else [None])
        if planes is not None:
            p4 = pose if pose is not None else [None] * 4
            rc = self._lib.gpp_fit_planes_host(self._h, planes.ctypes.data, planes.shape[0], 1 if planes.dtype == np.float64 else 0,
                                               0 if planes.flags['C_CONTIGUOUS'] else 1, _lib.ptr(boxes), _lib.ptr(dimensions),
                                               _lib.ptr(orientations), _lib.ptr(P_inv), B, D, _lib.ptr(keypoints),
                                               _lib.ptr(keyplanes), _lib.ptr(residuals), _lib.ptr(best), _lib.ptr(p4[0]),
                                               _lib.ptr(p4[1]), _lib.ptr(p4[2]), _lib.ptr(p4[3]), _lib.MODES[mode])
            _lib.check(rc, 'gpp_fit_planes_host')
            self._dev_planes = None
        elif pose is not None:
            rc = self._lib.gpp_fit_pose_host(self._h, _lib.ptr(boxes), _lib.ptr(dimensions), _lib.ptr(orientations),
                                             _lib.ptr(P_inv), B, D, _lib.ptr(keypoints), _lib.ptr(keyplanes),
                                             _lib.ptr(residuals), _lib.ptr(best), _lib.ptr(pose[0]), _lib.ptr(pose[1]),
                                             _lib.ptr(pose[2]), _lib.ptr(pose[3]), _lib.MODES[mode])
            _lib.check(rc, 'gpp_fit_pose_host')
        elif mode == 'f64':
            rc = self._lib.gpp_fit_host_f64(self._h, _lib.ptr(boxes), _lib.ptr(dimensions), _lib.ptr(orientations),
                                            _lib.ptr(P_inv), B, D, _lib.ptr(keypoints), _lib.ptr(keyplanes),
                                            _lib.ptr(residuals), _lib.ptr(best))
            _lib.check(rc, 'gpp_fit_host_f64')
        else:
            rc = self._lib.gpp_fit_host(self._h, _lib.ptr(boxes), _lib.ptr(dimensions), _lib.ptr(orientations),
                                        _lib.ptr(P_inv), B, D, _lib.ptr(keypoints), _lib.ptr(keyplanes),
                                        _lib.ptr(residuals), _lib.ptr(best), _lib.MODES[mode])
            _lib.check(rc, 'gpp_fit_host')
        out = [keypoints, keyplanes, residuals]
        if return_index:
            out.append(best)
        if return_pose:
            out += pose[:3]
        if return_kitti:
            out.append(pose[3])
        return out

    # ------------------------------------------------------------------ torch (device tensor) entry
    def set_planes_torch(self, planes):
        import torch
        if not planes.is_cuda or planes.device.index != self.device:
            return self.set_planes(planes.detach().cpu().numpy())
        p = planes.detach().to(torch.float32).contiguous()
        if p.dim() != 2 or p.shape[1] != 4 or p.shape[0] < 1:
            raise ValueError('planes must have shape (N, 4), got %r' % (tuple(p.shape),))
        # Skip the upload only for the very same tensor OBJECT with an unchanged version counter.  The object is kept
        # alive here, so neither its identity nor its storage can be recycled for other content (a data_ptr / shape
        # key would be: the caching allocator hands a freed block to the next tensor of the same size).
        if self._dev_planes is not None and self._dev_planes[0] is planes and self._dev_planes[1] == planes._version:
            return
        stream = torch.cuda.current_stream(self.device).cuda_stream
        _lib.check(self._lib.gpp_set_planes_device(self._h, ctypes.c_void_p(p.data_ptr()), p.shape[0],
                                                   ctypes.c_void_p(stream)), 'gpp_set_planes_device')
        self._dev_planes = (planes, planes._version)

    def fit_torch(self, boxes, dimensions, orientations, P_inv, mode=None, return_index=False, return_pose=False,
                  return_kitti=False):
        """Device-resident call: CUDA tensors in, CUDA tensors out, enqueued on torch's current stream,
        no host synchronisation.  ``return_pose`` / ``return_kitti`` as in ``fit``."""
        import torch
        mode = DEFAULT_MODE if mode is None else mode
        if mode not in _lib.MODES:
            raise ValueError('unknown mode %r' % (mode,))
        dev = torch.device('cuda', self.device)
        for name, t in (('boxes', boxes), ('dimensions', dimensions), ('orientations', orientations),
                        ('P_inv', P_inv)):
            if not isinstance(t, torch.Tensor) or t.device != dev:
                raise ValueError('%s must be a torch tensor on %s' % (name, dev))
        boxes = boxes.to(torch.float32).contiguous()
        if boxes.dim() != 3 or boxes.shape[2] != 12:
            raise ValueError('boxes must have shape (B, D, 12), got %r' % (tuple(boxes.shape),))
        B, D = int(boxes.shape[0]), int(boxes.shape[1])
        dimensions = dimensions.to(torch.float32).contiguous()
        orientations = orientations.to(torch.int32).contiguous()
        P_inv = P_inv.to(torch.float32).contiguous()
        if tuple(dimensions.shape) != (B, D, 3) or tuple(orientations.shape) != (B, D) or \
                tuple(P_inv.shape) != (B, 4, 3):
            raise ValueError('inconsistent shapes: dimensions %r orientations %r P_inv %r for B=%d D=%d' % (
                tuple(dimensions.shape), tuple(orientations.shape), tuple(P_inv.shape), B, D))
        out_t = torch.float64 if mode == 'f64' else torch.float32
        keypoints = torch.empty((B, D, 4, 3), dtype=out_t, device=dev)
        keyplanes = torch.empty((B, D, 1, 4), dtype=out_t, device=dev)
        residuals = torch.empty((B, D), dtype=out_t, device=dev)
        best = torch.empty((B, D), dtype=torch.int64, device=dev) if return_index else None
        stream = ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        vp = lambda t: ctypes.c_void_p(t.data_ptr()) if t is not None and t.numel() else None  # noqa: E731
        pose = None
        if return_pose or return_kitti:
            if mode == 'f64':
                raise ValueError("return_pose / return_kitti are not available in mode 'f64'")
            pose = [torch.empty((B, D, 3), dtype=torch.float32, device=dev) for _ in range(3)] + \
                   ([torch.empty((B, D, 4), dtype=torch.float32, device=dev)] if return_kitti else [None])
        if B * D > 0:
            if pose is not None:
                rc = self._lib.gpp_fit_pose_device(self._h, vp(boxes), vp(dimensions), vp(orientations), vp(P_inv), B, D,
                                                   vp(keypoints), vp(keyplanes), vp(residuals), vp(best), vp(pose[0]),
                                                   vp(pose[1]), vp(pose[2]), vp(pose[3]), _lib.MODES[mode], stream)
                _lib.check(rc, 'gpp_fit_pose_device')
            elif mode == 'f64':
                rc = self._lib.gpp_fit_device_f64(self._h, vp(boxes), vp(dimensions), vp(orientations), vp(P_inv),
                                                  B, D, vp(keypoints), vp(keyplanes), vp(residuals), vp(best),
                                                  stream)
                _lib.check(rc, 'gpp_fit_device_f64')
            else:
                rc = self._lib.gpp_fit_device(self._h, vp(boxes), vp(dimensions), vp(orientations), vp(P_inv),
                                              B, D, vp(keypoints), vp(keyplanes), vp(residuals), vp(best),
                                              _lib.MODES[mode], stream)
                _lib.check(rc, 'gpp_fit_device')
        out = [keypoints, keyplanes, residuals]
        if return_index:
            out.append(best)
        if return_pose:
            out += pose[:3]
        if return_kitti:
            out.append(pose[3])
        return out

    # ------------------------------------------------------------------ measurement helpers
    def last_kernel_ms(self):
        ms = ctypes.c_float()
        _lib.check(self._lib.gpp_last_kernel_ms(self._h, ctypes.byref(ms)), 'gpp_last_kernel_ms')
        return float(ms.value)

    def launch_count(self):
        return int(self._lib.gpp_launch_count(self._h))

    def microbench(self, kind):
        ops, ms, opc = ctypes.c_double(), ctypes.c_float(), ctypes.c_double()
        _lib.check(self._lib.gpp_microbench(self._h, int(kind), ctypes.byref(ops), ctypes.byref(ms),
                                            ctypes.byref(opc)), 'gpp_microbench')
        return dict(ops_per_s=ops.value, ms=ms.value, ops_per_clk_sm=opc.value)

    def debug_scores(self, box12, dims3, orientation, pinv12, which=0, with_margin=False):
        """Test hook: (votes, residual sum, z_dir_check < 0 [, margin]) of one detection against every
        resident plane, from the device functions of the search loop (which: 0 exact, 1 fast general, 2 fast
        all-six path, 3 stage 1 of the verified all-six path: bottom-face residual sum and its margin); margin = the VERIFIED mode's bound on |fast - exact| of the residual sum."""
        n = self.num_planes
        votes, zneg, resid = np.empty(n, np.int32), np.empty(n, np.int32), np.empty(n, np.float32)
        margin = np.zeros(n, np.float32) if with_margin else None
        rc = self._lib.gpp_debug_scores(self._h, _lib.ptr(_f32(box12).reshape(12)), _lib.ptr(_f32(dims3).reshape(3)),
                                        int(orientation), _lib.ptr(_f32(pinv12).reshape(12)), int(which),
                                        _lib.ptr(votes), _lib.ptr(resid), _lib.ptr(zneg), _lib.ptr(margin))
        _lib.check(rc, 'gpp_debug_scores')
        if with_margin:
            # + what the VERIFIED filters test: votes possible within the margin, z-check passable within it
            return votes & 15, resid, (zneg & 1).astype(bool), margin, votes >> 4, (zneg & 2).astype(bool)
        return votes & 15, resid, (zneg & 1).astype(bool)

    def debug_set_schedule(self, n_seg=0, resident_rows=-1):
        """Test / tuning hook of the resident-database kernel: plane segments per detection (0 = automatic) and rows
        of 64 planes kept in shared memory (-1 = automatic, 0 = stream everything from L2)."""
        _lib.check(self._lib.gpp_debug_set_schedule(self._h, int(n_seg), int(resident_rows)),
                   'gpp_debug_set_schedule')

    # ------------------------------------------------------------------ runtime audit of the VERIFIED mode
    def audit_set(self, every):
        """Re-poll every ``every``-th detection of each 'verified' call in the EXACT arithmetic on the device and
        count disagreements (0 = off; the environment variable GPP_AUDIT=n does the same for new handles)."""
        _lib.check(self._lib.gpp_audit_set(self._h, int(every)), 'gpp_audit_set')

    def audit_counts(self):
        """(rows checked, rows whose plane index or residual differed) since the handle was created."""
        checked, bad = ctypes.c_int64(), ctypes.c_int64()
        _lib.check(self._lib.gpp_audit_counts(self._h, ctypes.byref(checked), ctypes.byref(bad)), 'gpp_audit_counts')
        return int(checked.value), int(bad.value)


_POLLERS = {}
_POLLERS_LOCK = threading.Lock()


def scan_order(planes):
    """The order in which the FAST / VERIFIED scans visit a database (csrc/gpp_order.cu): ``order[position] = plane
    index``.  Host code of libgpp, needs no device; ``planes``: (N, 4) as fed (float32 after the Keras cast)."""
    rows = np.ascontiguousarray(np.asarray(planes).reshape(-1, 4), dtype=np.float32)
    order = np.empty(rows.shape[0], np.int32)
    _lib.check(_lib.load().gpp_debug_scan_order(_lib.ptr(rows), rows.shape[0], _lib.ptr(order)), 'gpp_debug_scan_order')
    return order


def default_device():
    """LOCAL_RANK (one process per GPU under torchrun) or GPP_DEVICE, else 0."""
    for var in ('GPP_DEVICE', 'LOCAL_RANK'):
        if os.environ.get(var, '') != '':
            return int(os.environ[var])
    return 0


def get_poller(device=None):
    """Process-wide PlanePoller of a device (created on first use)."""
    device = default_device() if device is None else int(device)
    with _POLLERS_LOCK:
        p = _POLLERS.get(device)
        if p is None:
            p = _POLLERS[device] = PlanePoller(device)
        return p


def _plane_groups(planes, B):
    """Split the image axis into runs that share one database.  Reference callers always feed the same
    database for every image ((1, N, 4) from run_network.py:105, np.tile'd (B, N, 4) from
    preprocessing/kitti.py:220); genuinely different per-image databases are supported too."""
    planes = np.asarray(planes)
    if planes.ndim == 2:
        return [(0, B, planes)]
    if planes.ndim != 3 or planes.shape[2] != 4:
        raise ValueError('planes must have shape (B, N, 4), (1, N, 4) or (N, 4), got %r' % (planes.shape,))
    if planes.shape[0] == 1 or planes.strides[0] == 0:
        return [(0, B, planes[0])]
    if planes.shape[0] != B:
        raise ValueError('planes batch %d does not match boxes batch %d' % (planes.shape[0], B))
    if planes.flags['C_CONTIGUOUS'] and planes[0].nbytes > 0:
        # every image as ONE opaque item: neighbouring images are compared with memcmp instead of element by element
        # (a np.tile'd database for a large batch is gigabytes of compares otherwise)
        rows = planes.reshape(B, -1).view(np.dtype((np.void, planes[0].nbytes))).ravel()
        cuts = np.flatnonzero(rows[1:] != rows[:-1]) + 1
    else:
        cuts = np.array([b for b in range(1, B) if not np.array_equal(planes[b], planes[b - 1])], dtype=np.int64)
    bounds = [0] + [int(c) for c in cuts] + [B]
    return [(bounds[i], bounds[i + 1], planes[bounds[i]]) for i in range(len(bounds) - 1)]


def fit_road_planes(boxes, dimensions, orientations, P_inv, planes, mode=None, return_index=False, device=None,
                    out=None, return_pose=False, return_kitti=False):
    """ Identify 3D keypoints and keyplane for each detection (drop-in for fit_road_planes.py:49).
    Args
        boxes                 : (num_batch, num_dets, 12) boxes in (x1, y1, x2, y2, xl, yl, xm, ym, xr, yr, xt, yt) format.
        dimensions            : (num_batch, num_dets, 3) predicted (height, width, length) of object.
        orientations          : (num_batch, num_dets) predicted orientation class.
        P_inv                 : (num_batch, 4, 3) pseudo-inverse of camera projection matrices.
        planes                : (num_batch, num_planes, 4) road planes (also (1, N, 4) or (N, 4)).
    Returns
        A list of [keypoints, keyplanes, residuals].
        keypoints is shaped (num_batch, num_dets, 4, 3) and consists of the 3D location of each of 4 keypoints.
        keyplanes is shaped (num_batch, num_dets, 1, 4) and contains the fitted road plane corresponding to each detection.
        residuals is shaped (num_batch, num_dets) and contains the best 'error of fit' corresponding to the keyplane.
    Extensions (do not change the default return list): ``mode`` 'verified' | 'exact' | 'fast' | 'f64',
    ``return_index`` appends the winning plane index (int64), ``device`` picks the GPU, ``out`` is a list of
    preallocated result arrays (single shared database only), ``return_pose`` appends ``locations (B, D, 3)``,
    ``angles (B, D, 3)`` (Rodrigues vector) and the corrected ``dimensions (B, D, 3)`` of EVERY row as the driver
    computes them for the rows it keeps (bin/run_network.py:137-247), ``return_kitti`` appends the KITTI writer's
    ``(alpha, h, Y, r_y)`` rows ``(B, D, 4)`` (:297-327) -- both run in the polling kernel's epilogue (one launch).
    """
    poller = get_poller(device)
    boxes = np.asarray(boxes)
    if boxes.ndim != 3:
        raise ValueError('boxes must have shape (B, D, 12), got %r' % (boxes.shape,))
    B = boxes.shape[0]
    groups = _plane_groups(planes, B)
    if len(groups) == 1:
        return poller.fit(boxes, dimensions, orientations, P_inv, mode=mode, return_index=return_index, out=out,
                          return_pose=return_pose, return_kitti=return_kitti, planes=groups[0][2])
    if out is not None:
        raise ValueError('out= is only supported with one plane database shared by the batch')
    dimensions, orientations, P_inv = np.asarray(dimensions), np.asarray(orientations), np.asarray(P_inv)
    parts = []
    for b0, b1, db in groups:
        poller.set_planes(db)
        parts.append(poller.fit(boxes[b0:b1], dimensions[b0:b1], orientations[b0:b1], P_inv[b0:b1], mode=mode,
                                return_index=return_index, return_pose=return_pose, return_kitti=return_kitti))
    return [np.concatenate([p[i] for p in parts], axis=0) for i in range(len(parts[0]))]


def fit_road_planes_torch(boxes, dimensions, orientations, P_inv, planes, mode=None, return_index=False,
                          return_pose=False, return_kitti=False):
    """Same operator on CUDA tensors (zero-copy, torch's current stream, no host sync once the database is resident:
    a NEW device-resident database of 2048 planes or more is read back once to derive its scan order).  ``planes`` is one
    (N, 4) / (1, N, 4) database shared by the batch (torch tensor on any device, or numpy).  ``return_pose`` appends
    locations, angles and corrected dimensions, (B, D, 3) each, ``return_kitti`` the (B, D, 4) KITTI records -- computed in
    the polling kernel's epilogue."""
    import torch
    poller = get_poller(boxes.device.index if boxes.device.index is not None else torch.cuda.current_device())
    if isinstance(planes, torch.Tensor):
        if planes.dim() == 3:
            # one database per call: (1, N, 4), or a broadcast view of one (expand(): stride 0 along the batch).  A
            # materialised (B, N, 4) tensor would need a device-wide compare and a host sync to accept, so it is not.
            if planes.shape[0] != 1 and planes.stride(0) != 0:
                raise ValueError('fit_road_planes_torch takes one database per call: pass (N, 4), (1, N, 4) or an '
                                 'expand()ed view; use fit_road_planes (numpy) for per-image databases')
            planes = planes[0]
        poller.set_planes_torch(planes)
    else:
        groups = _plane_groups(planes, int(boxes.shape[0]))
        if len(groups) != 1:
            raise ValueError('fit_road_planes_torch takes one database per call')
        poller.set_planes(groups[0][2])
    return poller.fit_torch(boxes, dimensions, orientations, P_inv, mode=mode, return_index=return_index,
                            return_pose=return_pose, return_kitti=return_kitti)


def fit_road_planes_dlpack(boxes, dimensions, orientations, P_inv, planes, mode=None, return_index=False):
    """DLPack entry: every argument is any object exporting ``__dlpack__`` (CUDA memory); the results are
    torch CUDA tensors, which export ``__dlpack__`` themselves."""
    import torch
    conv = lambda x: x if isinstance(x, torch.Tensor) else torch.from_dlpack(x)  # noqa: E731
    return fit_road_planes_torch(conv(boxes), conv(dimensions), conv(orientations), conv(P_inv),
                                 planes if isinstance(planes, np.ndarray) else conv(planes),
                                 mode=mode, return_index=return_index)


class FitRoadPlanes(object):
    """ Layer for identifying 3D keypoints and keyplanes (mirror of the Keras layer, fit_road_planes.py:142-186).
    Not a Keras object (Keras/TensorFlow are not part of this build): it keeps the layer's methods so code
    that drives the layer directly keeps working.
    """

    def __init__(self, mode=None, device=None, **kwargs):
        self.name = kwargs.pop('name', 'fit_road_planes')
        self.mode = mode
        self.device = device
        self._kwargs = kwargs

    def call(self, inputs, **kwargs):
        """
        Args
            inputs : List of [boxes, dimensions, orientations, P_inv, planes] arrays.
        """
        boxes, dimensions, orientations, P_inv, planes = inputs[0], inputs[1], inputs[2], inputs[3], inputs[4]
        return fit_road_planes(boxes, dimensions, orientations, P_inv, planes, mode=self.mode, device=self.device)

    __call__ = call

    def compute_output_shape(self, input_shape):
        """ [(num_batch, num_dets, 4, 3), (num_batch, num_dets, 1, 4), (num_batch, num_dets)] (:165-173) """
        return [(input_shape[0][0], input_shape[0][1], 4, 3), (input_shape[0][0], input_shape[0][1], 1, 4),
                (input_shape[0][0], input_shape[0][1])]

    def compute_mask(self, inputs, mask=None):
        return len(inputs) * [None]

    def get_config(self):
        config = {'name': self.name}
        config.update(self._kwargs)
        return config
