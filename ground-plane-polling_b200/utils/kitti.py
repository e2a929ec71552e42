"""
Post-processing of the polled detections -- the part of the reference's driver that follows
``model.predict_on_batch`` (/root/reference/keras_retinanet_3D/bin/run_network.py:113-135 selection,
:137-247 pose, :295-330 KITTI records).  The per-detection arithmetic (pose, Rodrigues, box extents, angle
wrapping) runs on the GPU through libgpp (``gpp_pose_host`` / ``gpp_kitti_host``); the selection / sorting of
at most 100 rows per image and the text formatting are host-side numpy like in the reference.
"""
import numpy as np

from .. import _lib
from ..layers.fit_road_planes import get_poller
from .pose import recover_pose

__all__ = ['select_detections', 'kitti_records', 'format_kitti_lines', 'postprocess_image', 'image_detections']


def select_detections(scores, score_threshold=0.05, max_detections=100):
    """Indices of one image's rows kept by the driver: score above the threshold, sorted by decreasing score,
    at most ``max_detections`` (run_network.py:117-125)."""
    scores = np.asarray(scores)
    indices = np.where(scores > score_threshold)[0]
    order = np.argsort(-scores[indices])[:max_detections]
    return indices[order]


def kitti_records(locations, angles, dimensions, device=None):
    """(alpha, h, Y, r_y) per detection, float32 (n, 4) -- run_network.py:297-323 on the GPU."""
    loc = np.ascontiguousarray(locations, dtype=np.float32).reshape(-1, 3)
    ang = np.ascontiguousarray(angles, dtype=np.float32).reshape(-1, 3)
    dims = np.ascontiguousarray(dimensions, dtype=np.float32).reshape(-1, 3)
    n = loc.shape[0]
    if ang.shape[0] != n or dims.shape[0] != n:
        raise ValueError('locations, angles and dimensions must have the same number of rows')
    out = np.zeros((n, 4), np.float32)
    poller = get_poller(device)
    _lib.check(poller._lib.gpp_kitti_host(poller._h, _lib.ptr(loc), _lib.ptr(ang), _lib.ptr(dims), n, _lib.ptr(out)),
               'gpp_kitti_host')
    return out


def format_kitti_lines(boxes, dimensions, locations, scores, records, image_wh, name='Car'):
    """The text lines of run_network.py:325-326 (box clipped to the image)."""
    W, H = image_wh
    lines = []
    for i in range(len(scores)):
        alpha, h, Y, r_y = [float(v) for v in records[i]]
        lines.append('%s -1 -1 %.2f %.2f %.2f %.2f %.2f %.2f %.2f %.2f %.2f %.2f %.2f %.2f %.2f\n' % (
            name, alpha, max(float(boxes[i, 0]), 0.0), max(float(boxes[i, 1]), 0.0), min(float(boxes[i, 2]), W),
            min(float(boxes[i, 3]), H), h, float(dimensions[i, 1]), float(dimensions[i, 2]), float(locations[i, 0]),
            Y, float(locations[i, 2]), r_y, float(scores[i])))
    return lines


def postprocess_image(boxes, dimensions, scores, labels, orientations, keypoints, keyplanes, residuals, scale,
                      score_threshold=0.05, max_detections=100, device=None):
    """Everything the driver does with one image's network outputs (run_network.py:113-287 + the KITTI record
    arithmetic).  Inputs are the (100, ...) rows of ONE image; returns a dict like the driver's ``outputs``
    (:291) plus 'kitti' = (alpha, h, Y, r_y) rows and the selected 'orientations' / 'keyplanes'."""
    boxes = np.asarray(boxes, dtype=np.float32) / np.float32(scale)          # :114
    keep = select_detections(scores, score_threshold, max_detections)
    boxes = boxes[keep]
    dims = np.asarray(dimensions, dtype=np.float32)[keep]
    kp = np.asarray(keypoints, dtype=np.float32).reshape(-1, 12)[keep]
    orient = np.asarray(orientations)[keep]
    locations, angles, dims_out = recover_pose(kp, dims, orient, device=device)
    records = kitti_records(locations, angles, dims_out, device=device)
    return {'boxes': boxes[:, :4], 'keypoints': boxes[:, 4:], 'labels': np.asarray(labels)[keep],
            'scores': np.asarray(scores)[keep], 'locations': locations, 'angles': angles, 'dimensions': dims_out,
            'residuals': np.asarray(residuals)[keep], 'orientations': orient,
            'keyplanes': np.asarray(keyplanes).reshape(-1, 4)[keep], 'kitti': records}


def image_detections(boxes, dimensions, scores, labels, orientations, keypoints, keyplanes, score_threshold=0.05,
                     max_detections=100):
    """The per-image detection table of the reference's second caller (utils/eval.py:96-118, `_get_detections`):
    rows kept by the score filter in decreasing-score order, columns = 12 box values, 3 dimensions, score, the four
    polled 3-D key-points flattened to 12, the 4 key-plane parameters, orientation, label (34 columns).  Inputs are
    the (1, 100, ...) arrays the inference model returns -- ``keypoints`` (1, D, 4, 3) and ``keyplanes`` (1, D, 1, 4)
    exactly as ``fit_road_planes`` returns them."""
    scores = np.asarray(scores)
    keep = select_detections(scores[0], score_threshold, max_detections)
    plane_pts = np.asarray(keypoints)[0, keep, :, :]
    plane_pts = np.reshape(plane_pts, (plane_pts.shape[0], 12))
    planes = np.asarray(keyplanes)[0, keep, :, :]
    planes = np.reshape(planes, (planes.shape[0], 4))
    return np.concatenate([np.asarray(boxes)[0, keep, :], np.asarray(dimensions)[0, keep, :],
                           np.expand_dims(scores[0][keep], axis=1), plane_pts, planes,
                           np.expand_dims(np.asarray(orientations)[0, keep], axis=1),
                           np.expand_dims(np.asarray(labels)[0, keep], axis=1)], axis=1)
