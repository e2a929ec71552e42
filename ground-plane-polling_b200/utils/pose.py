"""
6-DoF pose recovery from the polled key-points on the GPU -- the step right after FitRoadPlanes in the
reference's driver (/root/reference/keras_retinanet_3D/bin/run_network.py:137-247; reachable branches only,
see csrc/gpp_pose.cu).  numpy in / numpy out through libgpp's ``gpp_pose_host``; device tensors through
``recover_pose_torch``.  No CPU fallback.
"""
import ctypes

import numpy as np

from .. import _lib
from ..layers.fit_road_planes import get_poller

__all__ = ['recover_pose', 'recover_pose_torch', 'kitti_yaw']


def recover_pose(keypoints, dimensions, orientations, device=None):
    """keypoints (n, 12) or (n, 4, 3); dimensions (n, 3) network (h, w, l); orientations (n,).
    Returns (locations (n, 3), angles (n, 3), dimensions (n, 3)) float32 like run_network.py:137-247:
    h and l are replaced by the measured key-point distances, w is kept; ``angles`` is the Rodrigues vector
    of [x_dir y_dir z_dir].  Rows whose orientation is outside 0..3 come back as zeros (the reference leaves
    uninitialised memory there)."""
    kp = np.ascontiguousarray(keypoints, dtype=np.float32).reshape(-1, 12)
    n = kp.shape[0]
    dims = np.ascontiguousarray(dimensions, dtype=np.float32)
    orient = np.ascontiguousarray(orientations, dtype=np.int32).reshape(-1)
    if dims.shape != (n, 3) or orient.shape != (n,):
        raise ValueError('inconsistent shapes: keypoints %r dimensions %r orientations %r' % (
            kp.shape, dims.shape, orient.shape))
    poller = get_poller(device)
    locations = np.zeros((n, 3), np.float32)
    angles = np.zeros((n, 3), np.float32)
    dims_out = dims.copy()
    rc = poller._lib.gpp_pose_host(poller._h, _lib.ptr(kp), _lib.ptr(dims), _lib.ptr(orient), n,
                                   _lib.ptr(locations), _lib.ptr(angles), _lib.ptr(dims_out))
    _lib.check(rc, 'gpp_pose_host')
    return locations, angles, dims_out


def recover_pose_torch(keypoints, dimensions, orientations):
    """Device-resident variant (CUDA tensors in/out, current stream, no host sync)."""
    import torch
    dev = keypoints.device
    poller = get_poller(dev.index)
    kp = keypoints.to(torch.float32).contiguous().view(-1, 12)
    n = int(kp.shape[0])
    dims = dimensions.to(torch.float32).contiguous().view(n, 3)
    orient = orientations.to(torch.int32).contiguous().view(n)
    locations = torch.zeros((n, 3), dtype=torch.float32, device=dev)
    angles = torch.zeros((n, 3), dtype=torch.float32, device=dev)
    dims_out = dims.clone()
    if n:
        stream = ctypes.c_void_p(torch.cuda.current_stream(dev.index).cuda_stream)
        vp = lambda t: ctypes.c_void_p(t.data_ptr())  # noqa: E731
        rc = poller._lib.gpp_pose_device(poller._h, vp(kp), vp(dims), vp(orient), n, vp(locations), vp(angles),
                                         vp(dims_out), stream)
        _lib.check(rc, 'gpp_pose_device')
    return locations, angles, dims_out


def kitti_yaw(angles):
    """r_y as the KITTI writer derives it (run_network.py:312-316): angles[:, 1] wrapped to [-pi, pi)."""
    r_y = np.asarray(angles)[:, 1] % (2 * np.pi)
    return np.where(r_y >= np.pi, r_y - 2 * np.pi, r_y)
