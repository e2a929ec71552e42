"""
FilterDetections on the GPU -- host-side mirror of /root/reference/keras_retinanet_3D/layers/filter_detections.py
(``filter_detections`` :18-189, layer ``FilterDetections`` :192-305) for the configuration the reference model is
built with (models/retinanet.py:415): one object class, ``class_specific_filter=True``,
``orientation_specific_filter=False``, ``nms=True``, no ``other`` tensors.  Other configurations raise.
numpy in / numpy out through libgpp's ``gpp_filter_host``; CUDA tensors through ``filter_detections_torch``.
"""
import ctypes

import numpy as np

from .. import _lib
from .fit_road_planes import get_poller

__all__ = ['filter_detections', 'filter_detections_batch', 'filter_detections_torch', 'FilterDetections']


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _check_config(other, class_specific_filter, orientation_specific_filter, nms):
    if other:
        raise NotImplementedError('filter_detections: `other` tensors are not supported (the reference model has none)')
    if not class_specific_filter or orientation_specific_filter or not nms:
        raise NotImplementedError('filter_detections: only the configuration of the reference model is built '
                                  '(class_specific_filter=True, orientation_specific_filter=False, nms=True)')


def filter_detections_batch(boxes, dimensions, classification, score_threshold=0.05, max_detections=100,
                            nms_threshold=0.5, device=None):
    """Batched form: boxes (B, A, 12), dimensions (B, A, 3), classification (B, A, 8) ->
    [boxes (B, max, 12), dimensions (B, max, 3), scores (B, max), labels (B, max) int32, orientations (B, max) int32],
    rows sorted by score and padded with -1 (what FilterDetections.call returns)."""
    boxes = _f32(boxes)
    if boxes.ndim != 3 or boxes.shape[2] != 12:
        raise ValueError('boxes must have shape (B, A, 12), got %r' % (boxes.shape,))
    B, A = boxes.shape[:2]
    dimensions, classification = _f32(dimensions), _f32(classification)
    if dimensions.shape != (B, A, 3) or classification.shape != (B, A, 8):
        raise ValueError('inconsistent shapes: dimensions %r classification %r for B=%d A=%d (one class: 3 and 8 columns)'
                         % (dimensions.shape, classification.shape, B, A))
    D = int(max_detections)
    out = [np.empty((B, D, 12), np.float32), np.empty((B, D, 3), np.float32), np.empty((B, D), np.float32),
           np.empty((B, D), np.int32), np.empty((B, D), np.int32)]
    poller = get_poller(device)
    rc = poller._lib.gpp_filter_host(poller._h, _lib.ptr(boxes), _lib.ptr(dimensions), _lib.ptr(classification), B, A,
                                     float(score_threshold), float(nms_threshold), D, *[_lib.ptr(o) for o in out])
    _lib.check(rc, 'gpp_filter_host')
    return out


def filter_detections(boxes, dimensions, classification, other=[], class_specific_filter=True,
                      orientation_specific_filter=False, nms=True, score_threshold=0.05, max_detections=100,
                      nms_threshold=0.5):
    """ Filter detections of ONE image using the boxes and classification values (drop-in for
    filter_detections.py:18; same argument order).
    Args
        boxes          : (num_boxes, 12) boxes in (x1, y1, x2, y2, xl, yl, xm, ym, xr, yr, xt, yt) format.
        dimensions     : (num_boxes, 3) (height, width, length).
        classification : (num_boxes, 8) classification scores.
    Returns
        [boxes (max_detections, 12), dimensions (max_detections, 3), scores, labels, orientations (max_detections,)],
        padded with -1.
    """
    _check_config(other, class_specific_filter, orientation_specific_filter, nms)
    out = filter_detections_batch(np.asarray(boxes)[None], np.asarray(dimensions)[None], np.asarray(classification)[None],
                                  score_threshold, max_detections, nms_threshold)
    return [o[0] for o in out]


def filter_detections_torch(boxes, dimensions, classification, score_threshold=0.05, max_detections=100,
                            nms_threshold=0.5):
    """Device-resident batched variant (CUDA tensors in / out on torch's current stream)."""
    import torch
    dev = boxes.device
    poller = get_poller(dev.index)
    boxes = boxes.to(torch.float32).contiguous()
    dimensions = dimensions.to(torch.float32).contiguous()
    classification = classification.to(torch.float32).contiguous()
    B, A = int(boxes.shape[0]), int(boxes.shape[1])
    D = int(max_detections)
    out = [torch.empty((B, D, 12), dtype=torch.float32, device=dev), torch.empty((B, D, 3), dtype=torch.float32, device=dev),
           torch.empty((B, D), dtype=torch.float32, device=dev), torch.empty((B, D), dtype=torch.int32, device=dev),
           torch.empty((B, D), dtype=torch.int32, device=dev)]
    if B:
        vp = lambda t: ctypes.c_void_p(t.data_ptr()) if t.numel() else None  # noqa: E731
        stream = ctypes.c_void_p(torch.cuda.current_stream(dev.index).cuda_stream)
        rc = poller._lib.gpp_filter_device(poller._h, vp(boxes), vp(dimensions), vp(classification), B, A,
                                           float(score_threshold), float(nms_threshold), D, *[vp(o) for o in out], stream)
        _lib.check(rc, 'gpp_filter_device')
    return out


class FilterDetections(object):
    """ Layer for filtering detections using score threshold and NMS (mirror of filter_detections.py:192-305). """

    def __init__(self, nms=True, class_specific_filter=True, orientation_specific_filter=False, nms_threshold=0.5,
                 score_threshold=0.05, max_detections=100, parallel_iterations=32, **kwargs):
        _check_config([], class_specific_filter, orientation_specific_filter, nms)
        self.nms = nms
        self.class_specific_filter = class_specific_filter
        self.orientation_specific_filter = orientation_specific_filter
        self.nms_threshold = nms_threshold
        self.score_threshold = score_threshold
        self.max_detections = max_detections
        self.parallel_iterations = parallel_iterations
        self.name = kwargs.get('name', 'filtered_detections')

    def call(self, inputs, **kwargs):
        """ inputs : List of [boxes, dimensions, classification] arrays (batched). """
        if len(inputs) != 3:
            raise NotImplementedError('FilterDetections: `other` tensors are not supported')
        return filter_detections_batch(inputs[0], inputs[1], inputs[2], self.score_threshold, self.max_detections,
                                       self.nms_threshold)

    __call__ = call

    def compute_output_shape(self, input_shape):
        return [(input_shape[0][0], self.max_detections, 12), (input_shape[1][0], self.max_detections, 3),
                (input_shape[1][0], self.max_detections), (input_shape[1][0], self.max_detections),
                (input_shape[1][0], self.max_detections)]

    def compute_mask(self, inputs, mask=None):
        return (len(inputs) + 2) * [None]

    def get_config(self):
        return {'name': self.name, 'nms': self.nms, 'class_specific_filter': self.class_specific_filter,
                'orientation_specific_filter': self.orientation_specific_filter, 'nms_threshold': self.nms_threshold,
                'score_threshold': self.score_threshold, 'max_detections': self.max_detections,
                'parallel_iterations': self.parallel_iterations}
