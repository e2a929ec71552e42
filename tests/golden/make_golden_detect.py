"""
Golden vectors for the decode (RegressBoxes / RegressDims) and FilterDetections steps, produced by EXECUTING THE
REFERENCE'S OWN FILES (keras_retinanet_3D/layers/_misc.py, backend/common.py, layers/filter_detections.py,
unmodified) over the numpy stand-ins of tf_numpy_shim.py.  Build container only:

    python tests/golden/make_golden_detect.py

The layers' default mean / std are passed explicitly as float32 arrays: TensorFlow converts the float64 numpy
constants to the tensors' float32 at graph construction, numpy would promote the whole expression to float64.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)

import tf_numpy_shim  # noqa: E402
from oracle.detect_ref import BOX_MEAN, BOX_STD, DIM_MEAN, DIM_STD  # noqa: E402


def synth_heads(B, A, seed, n_hot=40, tie=False, none_above=False):
    """Synthetic head outputs of the detector: anchors over a 1333x402 image, regression deltas ~ N(0, 1),
    sigmoid-like scores mostly below the 0.05 threshold with `n_hot` clusters of overlapping high scorers."""
    rng = np.random.default_rng(seed)
    f = np.float32
    cx, cy = rng.uniform(0, 1333, A), rng.uniform(0, 402, A)
    w, h = rng.uniform(16, 400, A), rng.uniform(16, 300, A)
    anchors = np.stack([cx - w / 2, cy - h / 2, cx + w / 2, cy + h / 2], 1).astype(f)
    regression = rng.normal(0, 1, (B, A, 12)).astype(f)
    regression_dim = rng.normal(0, 1, (B, A, 3)).astype(f)
    cls = rng.uniform(0.0, 0.04, (B, A, 8)).astype(f)
    if not none_above:
        for b in range(B):
            hot = rng.choice(A, size=n_hot, replace=False)
            for k in hot:
                near = np.argsort(np.abs(cx - cx[k]) + np.abs(cy - cy[k]))[:6]        # neighbours overlap -> NMS matters
                cls[b, near, rng.integers(0, 8, size=near.shape[0])] = rng.uniform(0.06, 0.99, size=near.shape[0])
        if tie:
            cls[:, : A // 4] = np.round(cls[:, : A // 4] * 16) / 16                   # many exactly equal scores
    return anchors, regression, cls, regression_dim


def cases():
    return {
        'detect_2x1500': synth_heads(2, 1500, 1),
        'detect_ties_1x1200': synth_heads(1, 1200, 2, tie=True),
        'detect_crowded_1x2500': synth_heads(1, 2500, 3, n_hot=300),           # > 100 survivors -> top-100 cut
        'detect_sparse_1x800': synth_heads(1, 800, 5, n_hot=6),                # fewer than 100 survivors -> padding
        'detect_empty_1x300': synth_heads(1, 300, 4, none_above=True),          # nothing above the threshold
    }


if __name__ == '__main__':
    misc = tf_numpy_shim.load_reference_module('/root/reference/keras_retinanet_3D/layers/_misc.py',
                                               'keras_retinanet_3D.layers._misc', with_common=True)
    filt = tf_numpy_shim.load_reference_module('/root/reference/keras_retinanet_3D/layers/filter_detections.py',
                                               'keras_retinanet_3D.layers.filter_detections')
    f32 = np.float32
    for name, (anchors, regression, cls, regression_dim) in cases().items():
        B = regression.shape[0]
        tiled = np.tile(anchors[None], (B, 1, 1))
        boxes = misc.RegressBoxes(mean=BOX_MEAN.astype(f32), std=BOX_STD.astype(f32)).call([tiled, regression, cls])
        dims = misc.RegressDims(mean=DIM_MEAN.astype(f32), std=DIM_STD.astype(f32)).call(regression_dim)
        assert boxes.dtype == f32 and dims.dtype == f32
        layer = filt.FilterDetections()
        out = layer.call([boxes, dims, cls])
        shapes = layer.compute_output_shape([boxes.shape, dims.shape, cls.shape])
        assert [tuple(s) for s in shapes] == [o.shape for o in out]
        np.savez_compressed(os.path.join(HERE, name + '.npz'), anchors=anchors, regression=regression,
                            classification=cls, regression_dim=regression_dim, boxes=np.asarray(boxes),
                            dimensions=np.asarray(dims), f_boxes=out[0], f_dimensions=out[1], f_scores=out[2],
                            f_labels=out[3], f_orientations=out[4])
        print('%-24s anchors %d kept %s' % (name, anchors.shape[0], [(int((o >= 0).sum())) for o in out[2]]))
