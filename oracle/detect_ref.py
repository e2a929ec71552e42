"""
ORACLE (test infrastructure, NOT product code) -- numpy restatement of the two steps that precede ground-plane
polling in the reference's inference graph:

  decode_ref            RegressBoxes.call   /root/reference/keras_retinanet_3D/layers/_misc.py:132-140
                        bbox_transform_inv  /root/reference/keras_retinanet_3D/backend/common.py:43-84
                        RegressDims.call    layers/_misc.py:185-186, dim_transform_inv backend/common.py:23-40
  filter_detections_ref filter_detections   /root/reference/keras_retinanet_3D/layers/filter_detections.py:18-189
                        for the configuration the model is built with (models/retinanet.py:415: one class,
                        class_specific_filter=True, orientation_specific_filter=False, nms=True, no `other`).

Pinning: the reference files are executed unmodified over numpy stand-ins (tests/golden/tf_numpy_shim.py,
make_golden_detect.py) and the vectors are committed under tests/golden/.  Not pinned (TensorFlow is absent):
the behaviour of tf.image.non_max_suppression / tf.nn.top_k among EXACTLY equal scores -- assumed lower index
first -- and the fp32 rounding inside TF's IoU (restated here from the kernel's published formula).
float32 throughout, no FMA contraction, operations in the order the reference writes them.
"""
import numpy as np

BOX_MEAN = np.array([-0.0373, -0.0165, 0.0373, 0.0171, -0.0286, -0.0478, 0.2929, 0.0114, 0.0288, -0.0589, 0.2932,
                     -0.0007])                                                    # layers/_misc.py:115
BOX_STD = np.array([0.1957, 0.1896, 0.1957, 0.1897, 0.1967, 0.2034, 0.2046, 0.1898, 0.1964, 0.2052, 0.2048,
                    0.1903])                                                      # layers/_misc.py:117
DIM_MEAN = np.array([1.6570, 1.7999, 4.2907])                                     # layers/_misc.py:168
DIM_STD = np.array([0.2681, 0.2243, 0.6281])                                      # layers/_misc.py:170


def decode_ref(anchors, regression, classification, regression_dim, box_mean=BOX_MEAN, box_std=BOX_STD,
               dim_mean=DIM_MEAN, dim_std=DIM_STD):
    """anchors (B, A, 4) or (A, 4); regression (B, A, 12); classification (B, A, 8); regression_dim (B, A, 3).
    Returns boxes (B, A, 12), dimensions (B, A, 3) float32."""
    f = np.float32
    a = np.asarray(anchors, dtype=f)
    if a.ndim == 2:
        a = a[None]
    d = np.asarray(regression, dtype=f)
    c = np.asarray(classification, dtype=f)
    mean, std = np.asarray(box_mean, dtype=f), np.asarray(box_std, dtype=f)
    # sign of the x offsets of the middle / top key-points: -1 if the arg-max score is in the first half (:133-136)
    half = c.shape[2] // 2
    sign = np.where(np.argmax(c, axis=2) < half, f(-1), f(1)).astype(f)
    width = a[:, :, 2] - a[:, :, 0]
    height = a[:, :, 3] - a[:, :, 1]
    t = lambda k: d[:, :, k] * std[k] + mean[k]  # noqa: E731
    x1 = a[:, :, 0] + t(0) * width
    y1 = a[:, :, 1] + t(1) * height
    x2 = a[:, :, 2] + t(2) * width
    y2 = a[:, :, 3] + t(3) * height
    xl = a[:, :, 0] + t(4) * width
    yl = a[:, :, 3] + t(5) * height
    xm = (a[:, :, 0] + a[:, :, 2]) / f(2) + (t(6) * width) * sign
    ym = a[:, :, 3] + t(7) * height
    xr = a[:, :, 2] + t(8) * width
    yr = a[:, :, 3] + t(9) * height
    xt = (a[:, :, 0] + a[:, :, 2]) / f(2) + (t(10) * width) * sign
    yt = a[:, :, 1] + t(11) * height
    boxes = np.stack([x1, y1, x2, y2, xl, yl, xm, ym, xr, yr, xt, yt], axis=2).astype(f)
    dims = (np.asarray(regression_dim, dtype=f) * np.asarray(dim_std, dtype=f) + np.asarray(dim_mean, dtype=f)).astype(f)
    return boxes, dims


def tf_iou_gt(bi, bj, thr):
    """TF's IOUGreaterThanThreshold on two boxes given in the layout the reference passes ((x1, y1, x2, y2);
    the formula is symmetric in the axis naming), float32."""
    f = np.float32
    y1i, x1i, y2i, x2i = min(bi[0], bi[2]), min(bi[1], bi[3]), max(bi[0], bi[2]), max(bi[1], bi[3])
    y1j, x1j, y2j, x2j = min(bj[0], bj[2]), min(bj[1], bj[3]), max(bj[0], bj[2]), max(bj[1], bj[3])
    ai = f(f(y2i - y1i) * f(x2i - x1i))
    aj = f(f(y2j - y1j) * f(x2j - x1j))
    if ai <= 0 or aj <= 0:
        return False
    ih = max(f(min(y2i, y2j) - max(y1i, y1j)), f(0))
    iw = max(f(min(x2i, x2j) - max(x1i, x1j)), f(0))
    inter = f(ih * iw)
    return bool(f(inter / f(f(ai + aj) - inter)) > f(thr))


def nms_ref(boxes4, scores, max_output_size, iou_threshold):
    """Greedy non-maximum suppression in descending score order (ties: lower index first)."""
    order = np.lexsort((np.arange(scores.shape[0]), -scores))
    sel = []
    for i in order:
        if len(sel) >= max_output_size:
            break
        if not any(tf_iou_gt(boxes4[i], boxes4[j], iou_threshold) for j in sel):
            sel.append(int(i))
    return np.asarray(sel, dtype=np.int64)


def filter_detections_ref(boxes, dimensions, classification, score_threshold=0.05, max_detections=100,
                          nms_threshold=0.5):
    """boxes (B, A, 12), dimensions (B, A, 3), classification (B, A, 8) ->
    [boxes (B, 100, 12) f32, dimensions (B, 100, 3) f32, scores (B, 100) f32, labels (B, 100) i32,
     orientations (B, 100) i32], padded with -1 (filter_detections.py:170-177)."""
    f = np.float32
    boxes = np.asarray(boxes, dtype=f)
    dimensions = np.asarray(dimensions, dtype=f)
    classification = np.asarray(classification, dtype=f)
    B = boxes.shape[0]
    D = max_detections
    out_b = np.full((B, D, 12), -1, f)
    out_d = np.full((B, D, 3), -1, f)
    out_s = np.full((B, D), -1, f)
    out_l = np.full((B, D), -1, np.int32)
    out_o = np.full((B, D), -1, np.int32)
    for b in range(B):
        c = classification[b]
        c4 = np.maximum(c[:, :4], c[:, 4:])                       # max over the two halves (:66-67)
        orient = np.argmax(c4, axis=1)                            # first maximum (:118)
        scores = np.max(c4, axis=1)                               # (:119)
        idx = np.where(scores > f(score_threshold))[0]            # (:53)
        keep = idx[nms_ref(boxes[b][idx, :4], scores[idx], D, nms_threshold)]          # (:55-64)
        s = scores[keep]
        top = np.lexsort((np.arange(s.shape[0]), -s))[:D]         # top_k, ties by lower position (:160)
        sel = keep[top]
        n = sel.shape[0]
        out_b[b, :n] = boxes[b][sel]
        out_d[b, :n] = dimensions[b][sel]
        out_s[b, :n] = s[top]
        out_l[b, :n] = 0
        out_o[b, :n] = orient[sel]
    return [out_b, out_d, out_s, out_l, out_o]
