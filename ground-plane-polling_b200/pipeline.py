"""
The inference tail of the reference model as one device-resident pipeline
(/root/reference/keras_retinanet_3D/models/retinanet.py:411-419): detector head outputs ->
RegressBoxes / RegressDims -> FilterDetections -> FitRoadPlanes, i.e. everything ``model.predict_on_batch`` computes
after the CNN (bin/run_network.py:110).  The CNN itself is out of scope: ``heads`` are whatever detector produced
them (tests use synthetic heads).  CUDA tensors in, CUDA tensors out, torch's current stream, no host sync.
"""
from .layers._misc import decode_torch
from .layers.filter_detections import filter_detections_torch
from .layers.fit_road_planes import fit_road_planes_torch

__all__ = ['detections_from_heads']


def detections_from_heads(anchors, regression, regression_dim, classification, P_inv, planes, mode=None,
                          score_threshold=0.05, max_detections=100, nms_threshold=0.5, return_pose=False,
                          return_kitti=False):
    """anchors (A, 4), regression (B, A, 12), regression_dim (B, A, 3), classification (B, A, 8), P_inv (B, 4, 3),
    planes (N, 4).  Returns the reference model's output list
    [boxes, dimensions, scores, labels, orientations, keypoints, keyplanes, residuals] (models/retinanet.py:418);
    ``return_pose`` / ``return_kitti`` append what the driver computes from it afterwards (locations, angles,
    dimensions; KITTI records -- bin/run_network.py:137-247, :297-327), from the polling kernel's epilogue: the whole
    post-CNN tail of an image in four launches."""
    boxes, dims = decode_torch(anchors, regression, classification, regression_dim)
    det = filter_detections_torch(boxes, dims, classification, score_threshold, max_detections, nms_threshold)
    poll = fit_road_planes_torch(det[0], det[1], det[4], P_inv, planes, mode=mode, return_pose=return_pose,
                                 return_kitti=return_kitti)
    return det + poll
