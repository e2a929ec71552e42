"""CPU: the decode / FilterDetections oracle against the golden vectors produced by executing the reference's own
_misc.py, backend/common.py and filter_detections.py over numpy op stand-ins (tests/golden/make_golden_detect.py)."""
import os

import numpy as np
import pytest

from conftest import GOLDEN
from oracle.detect_ref import decode_ref, filter_detections_ref, nms_ref, tf_iou_gt

CASES = sorted(f[:-4] for f in os.listdir(GOLDEN) if f.startswith('detect_') and f.endswith('.npz'))


def _eq(a, b):
    return a.shape == b.shape and a.dtype == b.dtype and np.array_equal(a, b)


@pytest.mark.parametrize('name', CASES)
def test_decode_oracle_matches_reference_layers(name):
    g = np.load(os.path.join(GOLDEN, name + '.npz'))
    boxes, dims = decode_ref(g['anchors'], g['regression'], g['classification'], g['regression_dim'])
    assert _eq(boxes, g['boxes']) and _eq(dims, g['dimensions'])


@pytest.mark.parametrize('name', CASES)
def test_filter_oracle_matches_reference_layer(name):
    g = np.load(os.path.join(GOLDEN, name + '.npz'))
    out = filter_detections_ref(g['boxes'], g['dimensions'], g['classification'])
    for o, k in zip(out, ('f_boxes', 'f_dimensions', 'f_scores', 'f_labels', 'f_orientations')):
        assert _eq(o, g[k]), k


def test_golden_set_covers_padding_topk_ties_and_empty():
    names = set(CASES)
    assert {'detect_sparse_1x800', 'detect_empty_1x300', 'detect_crowded_1x2500', 'detect_ties_1x1200'} <= names
    g = np.load(os.path.join(GOLDEN, 'detect_sparse_1x800.npz'))
    n = int((g['f_scores'][0] >= 0).sum())
    assert 0 < n < 100 and (g['f_boxes'][0, n:] == -1).all() and (g['f_labels'][0, n:] == -1).all()
    assert (np.load(os.path.join(GOLDEN, 'detect_empty_1x300.npz'))['f_scores'] == -1).all()
    s = np.load(os.path.join(GOLDEN, 'detect_ties_1x1200.npz'))['f_scores'][0]
    assert (np.diff(s[s >= 0]) <= 0).all() and len(np.unique(s[s >= 0])) < (s >= 0).sum()     # sorted, with ties


def test_nms_semantics():
    f = np.float32
    boxes = np.array([[0, 0, 10, 10], [1, 1, 11, 11], [20, 20, 30, 30], [0, 0, 10, 10], [5, 5, 5, 9]], f)
    scores = np.array([0.9, 0.8, 0.7, 0.9, 0.95], f)
    assert tf_iou_gt(boxes[0], boxes[1], 0.5) and not tf_iou_gt(boxes[0], boxes[2], 0.5)
    assert not tf_iou_gt(boxes[4], boxes[0], 0.0)                 # zero-area box never suppresses / is suppressed
    keep = nms_ref(boxes, scores, 100, 0.5)
    assert keep.tolist() == [4, 0, 2]                              # ties: lower index (0 before 3), 1 and 3 suppressed
    assert nms_ref(boxes, scores, 2, 0.5).tolist() == [4, 0]
