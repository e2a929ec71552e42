"""
Box / dimension decoding on the GPU -- host-side mirror of the reference's ``RegressBoxes`` and ``RegressDims``
layers (/root/reference/keras_retinanet_3D/layers/_misc.py:103-199; arithmetic in
backend/common.py:23-84 ``dim_transform_inv`` / ``bbox_transform_inv``).  numpy in / numpy out through libgpp's
``gpp_decode_host``; CUDA tensors through ``decode_torch``.  No CPU fallback.
"""
import ctypes

import numpy as np

from .. import _lib
from .fit_road_planes import get_poller

__all__ = ['RegressBoxes', 'RegressDims', 'decode', 'decode_torch', 'BOX_MEAN', 'BOX_STD', 'DIM_MEAN', 'DIM_STD']

BOX_MEAN = np.array([-0.0373, -0.0165, 0.0373, 0.0171, -0.0286, -0.0478, 0.2929, 0.0114, 0.0288, -0.0589, 0.2932, -0.0007])
BOX_STD = np.array([0.1957, 0.1896, 0.1957, 0.1897, 0.1967, 0.2034, 0.2046, 0.1898, 0.1964, 0.2052, 0.2048, 0.1903])
DIM_MEAN = np.array([1.6570, 1.7999, 4.2907])
DIM_STD = np.array([0.2681, 0.2243, 0.6281])


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _mean_std(mean, std, n, dmean, dstd):
    mean = dmean if mean is None else np.asarray(mean)
    std = dstd if std is None else np.asarray(std)
    if mean.shape != (n,) or std.shape != (n,):
        raise ValueError('mean and std must have shape (%d,)' % n)
    return _f32(np.concatenate([mean, std]))


def decode(anchors, regression, classification, regression_dim, box_mean=None, box_std=None, dim_mean=None,
           dim_std=None, device=None):
    """RegressBoxes + RegressDims in one pass.  anchors (A, 4) or (B, A, 4) [one anchor set for the batch],
    regression (B, A, 12), classification (B, A, 8), regression_dim (B, A, 3) -> boxes (B, A, 12), dims (B, A, 3)."""
    regression = _f32(regression)
    if regression.ndim != 3 or regression.shape[2] != 12:
        raise ValueError('regression must have shape (B, A, 12), got %r' % (regression.shape,))
    B, A = regression.shape[:2]
    anchors = _f32(anchors)
    if anchors.ndim == 3:
        if anchors.shape[0] != 1 and not np.array_equal(anchors, np.broadcast_to(anchors[:1], anchors.shape)):
            raise ValueError('decode takes one anchor set for the whole batch')
        anchors = _f32(anchors[0])
    classification = _f32(classification)
    regression_dim = _f32(regression_dim)
    if anchors.shape != (A, 4) or classification.shape != (B, A, 8) or regression_dim.shape != (B, A, 3):
        raise ValueError('inconsistent shapes: anchors %r classification %r regression_dim %r for B=%d A=%d' % (
            anchors.shape, classification.shape, regression_dim.shape, B, A))
    bms = _mean_std(box_mean, box_std, 12, BOX_MEAN, BOX_STD)
    dms = _mean_std(dim_mean, dim_std, 3, DIM_MEAN, DIM_STD)
    boxes = np.empty((B, A, 12), np.float32)
    dims = np.empty((B, A, 3), np.float32)
    poller = get_poller(device)
    rc = poller._lib.gpp_decode_host(poller._h, _lib.ptr(anchors), _lib.ptr(regression), _lib.ptr(classification),
                                     _lib.ptr(regression_dim), B, A, _lib.ptr(bms), _lib.ptr(dms), _lib.ptr(boxes),
                                     _lib.ptr(dims))
    _lib.check(rc, 'gpp_decode_host')
    return boxes, dims


def decode_torch(anchors, regression, classification, regression_dim, box_mean=None, box_std=None, dim_mean=None,
                 dim_std=None):
    """Device-resident variant: CUDA tensors in / out on torch's current stream."""
    import torch
    dev = regression.device
    poller = get_poller(dev.index)
    regression = regression.to(torch.float32).contiguous()
    B, A = int(regression.shape[0]), int(regression.shape[1])
    anchors = anchors.to(torch.float32)
    anchors = (anchors[0] if anchors.dim() == 3 else anchors).contiguous()
    classification = classification.to(torch.float32).contiguous()
    regression_dim = regression_dim.to(torch.float32).contiguous()
    bms = _mean_std(box_mean, box_std, 12, BOX_MEAN, BOX_STD)
    dms = _mean_std(dim_mean, dim_std, 3, DIM_MEAN, DIM_STD)
    boxes = torch.empty((B, A, 12), dtype=torch.float32, device=dev)
    dims = torch.empty((B, A, 3), dtype=torch.float32, device=dev)
    if B * A:
        vp = lambda t: ctypes.c_void_p(t.data_ptr())  # noqa: E731
        stream = ctypes.c_void_p(torch.cuda.current_stream(dev.index).cuda_stream)
        rc = poller._lib.gpp_decode_device(poller._h, vp(anchors), vp(regression), vp(classification), vp(regression_dim),
                                           B, A, _lib.ptr(bms), _lib.ptr(dms), vp(boxes), vp(dims), stream)
        _lib.check(rc, 'gpp_decode_device')
    return boxes, dims


class RegressBoxes(object):
    """ Layer for applying regression values to boxes (mirror of layers/_misc.py:103-153). """

    def __init__(self, mean=None, std=None, *args, **kwargs):
        self.mean = BOX_MEAN if mean is None else np.asarray(mean)
        self.std = BOX_STD if std is None else np.asarray(std)
        self.name = kwargs.get('name', 'boxes')

    def call(self, inputs, **kwargs):
        anchors, regression, classification = inputs
        B, A = np.asarray(regression).shape[:2]
        return decode(anchors, regression, classification, np.zeros((B, A, 3), np.float32), box_mean=self.mean,
                      box_std=self.std)[0]

    __call__ = call

    def compute_output_shape(self, input_shape):
        return (input_shape[1][0], input_shape[1][1], 12)

    def get_config(self):
        return {'name': self.name, 'mean': self.mean.tolist(), 'std': self.std.tolist()}


class RegressDims(object):
    """ Layer for applying regression values to dimensions (mirror of layers/_misc.py:156-199). """

    def __init__(self, mean=None, std=None, *args, **kwargs):
        self.mean = DIM_MEAN if mean is None else np.asarray(mean)
        self.std = DIM_STD if std is None else np.asarray(std)
        self.name = kwargs.get('name', 'dims')

    def call(self, inputs, **kwargs):
        rd = np.asarray(inputs)
        B, A = rd.shape[:2]
        z = np.zeros((B, A, 12), np.float32)
        return decode(np.zeros((A, 4), np.float32), z, np.zeros((B, A, 8), np.float32), rd, dim_mean=self.mean,
                      dim_std=self.std)[1]

    __call__ = call

    def compute_output_shape(self, input_shape):
        return input_shape

    def get_config(self):
        return {'name': self.name, 'mean': self.mean.tolist(), 'std': self.std.tolist()}
