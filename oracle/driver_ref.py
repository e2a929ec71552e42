"""
ORACLE (test infrastructure, NOT product code) -- the reference driver's per-image work after the CNN, restated
from /root/reference/keras_retinanet_3D/bin/run_network.py:110-326 as one chain over the other oracles:

  heads -> decode_ref -> filter_detections_ref -> C/numpy polling oracle      (= model.predict_on_batch, :110)
        -> boxes /= scale, score filter, argsort, at most 100 rows            (:113-135)
        -> pose_ref                                                           (:137-287)
        -> outputs dict of the .mat file                                      (:291)
        -> KITTI text lines                                                   (:295-326)

Pinning: every link is pinned separately (tests/golden/*.npz), and the selection / pose / outputs / KITTI-line part
(driver_image_ref) equals, bit for bit and character for character, what the reference's own lines :113-330 produce
when they are cut out of run_network.py and exec'd (tests/golden/pose_*.npz, tests/test_pose_oracle_golden.py).  The image utilities and the anchors are pinned by tests/golden/driver_utils.npz
(make_golden_driver.py runs the reference's own utils/anchors.py and utils/image.py).
"""
import numpy as np

from .c_oracle import fit_road_planes_c
from .detect_ref import decode_ref, filter_detections_ref
from .kitti_ref import kitti_records_ref
from .pose_ref import pose_ref


def predict_on_batch_ref(anchors, regression, regression_dim, classification, P_inv, planes):
    """The 8 outputs of the inference model for a batch of head tensors (models/retinanet.py:411-419)."""
    boxes, dims = decode_ref(anchors, regression, classification, regression_dim)
    det = filter_detections_ref(boxes, dims, classification)
    kp, kpl, res = fit_road_planes_c(det[0], det[1], det[4], P_inv, planes)
    return det + [kp, kpl, res]


def driver_image_ref(outs, scale, image_wh):
    """run_network.py:113-326 for ONE image (batch index 0 of ``outs``).  Returns (outputs dict, KITTI lines)."""
    boxes, dimensions, scores, labels, orientations, keypoints, keyplanes, residuals = [np.array(o, copy=True) for o in outs]
    boxes /= scale
    indices = np.where(scores[0, :] > 0.05)[0]
    scores = scores[0][indices]
    scores_sort = np.argsort(-scores)[:100]
    boxes = boxes[0, indices[scores_sort], :]
    dimensions = dimensions[0, indices[scores_sort], :]
    scores = scores[scores_sort]
    labels = labels[0, indices[scores_sort]]
    orientations = orientations[0, indices[scores_sort]]
    keypoints = np.reshape(keypoints[0, indices[scores_sort], :, :], (-1, 12))
    residuals = residuals[0, indices[scores_sort]]
    locations, angles, dimensions = pose_ref(keypoints, dimensions, orientations)
    outputs = {'boxes': boxes[:, :4], 'keypoints': boxes[:, 4:], 'labels': labels, 'scores': scores,
               'locations': locations, 'angles': angles, 'dimensions': dimensions, 'residuals': residuals}
    rec = kitti_records_ref(locations, angles, dimensions)          # (alpha, h, Y, r_y) per row
    W, H = image_wh
    lines = []
    for i in range(len(scores)):
        alpha, h, Y, r_y = rec[i]
        lines.append("Car -1 -1 %.2f %.2f %.2f %.2f %.2f %.2f %.2f %.2f %.2f %.2f %.2f %.2f %.2f\n" % (
            alpha, np.maximum(boxes[i, 0], 0.0), np.maximum(boxes[i, 1], 0.0), np.minimum(boxes[i, 2], W),
            np.minimum(boxes[i, 3], H), h, dimensions[i, 1], dimensions[i, 2], locations[i, 0], Y, locations[i, 2],
            r_y, scores[i]))
    return outputs, lines
