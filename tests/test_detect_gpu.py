"""GPU: decode (RegressBoxes / RegressDims) and FilterDetections kernels, through the C ABI, against the oracle
and the golden vectors; the composed device-resident pipeline heads -> boxes -> detections -> polling."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, load_planes
from oracle import c_oracle
from oracle.detect_ref import decode_ref, filter_detections_ref

pytestmark = pytest.mark.gpu

CASES = sorted(f[:-4] for f in os.listdir(GOLDEN) if f.startswith('detect_') and f.endswith('.npz'))


def _same(got, want):
    for g, w in zip(got, want):
        assert g.shape == w.shape and g.dtype == w.dtype
        assert np.array_equal(g, w)


@pytest.mark.parametrize('name', CASES)
def test_decode_and_filter_equal_golden_vectors(gpp, name):
    g = np.load(os.path.join(GOLDEN, name + '.npz'))
    boxes, dims = gpp.decode(g['anchors'], g['regression'], g['classification'], g['regression_dim'])
    _same([boxes, dims], [g['boxes'], g['dimensions']])
    out = gpp.FilterDetections().call([boxes, dims, g['classification']])
    _same(out, [g[k] for k in ('f_boxes', 'f_dimensions', 'f_scores', 'f_labels', 'f_orientations')])
    # the layer mirrors
    assert np.array_equal(gpp.RegressBoxes().call([g['anchors'], g['regression'], g['classification']]), g['boxes'])
    assert np.array_equal(gpp.RegressDims().call(g['regression_dim']), g['dimensions'])
    one = gpp.filter_detections(boxes[0], dims[0], g['classification'][0])
    _same(one, [g[k][0] for k in ('f_boxes', 'f_dimensions', 'f_scores', 'f_labels', 'f_orientations')])


def _heads(B, A, seed, hot):
    import importlib.util
    spec = importlib.util.spec_from_file_location('mgd', os.path.join(GOLDEN, 'make_golden_detect.py'))
    # only synth_heads is needed; the module's top level imports the shim, which is CPU-only and harmless
    mgd = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mgd)
    return mgd.synth_heads(B, A, seed, n_hot=hot)


def test_large_candidate_lists_and_many_images(gpp):
    """More candidates than the shared-memory sort holds (global-memory sort path), many images per call."""
    anchors, reg, cls, rdim = _heads(3, 20000, 11, 200)
    cls = cls.copy()
    cls[0] += 0.05                                        # image 0: every anchor above the threshold (20000 candidates)
    boxes, dims = gpp.decode(anchors, reg, cls, rdim)
    wb, wd = decode_ref(anchors, reg, cls, rdim)
    _same([boxes, dims], [wb, wd])
    out = gpp.filter_detections_batch(boxes, dims, cls)
    want = filter_detections_ref(wb, wd, cls)
    _same(out, want)
    with pytest.raises(NotImplementedError):
        gpp.FilterDetections(nms=False)
    with pytest.raises(ValueError):
        gpp.filter_detections_batch(boxes, dims[:, :, :2], cls)


def test_device_pipeline_heads_to_polled_detections(gpp):
    """models/retinanet.py:411-419 after the CNN, device-resident: decode -> filter -> poll; equals the oracles
    chained on the host."""
    import torch
    from gpp_b200.utils import synthetic
    planes = load_planes('1k')
    anchors, reg, cls, rdim = _heads(2, 4000, 21, 30)
    _, P_inv = synthetic.kitti_calibration()
    P_inv = np.tile(P_inv[None], (2, 1, 1))
    dev = torch.device('cuda', 0)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    out = gpp.detections_from_heads(t(anchors), t(reg), t(rdim), t(cls), t(P_inv.astype(np.float32)), planes)
    torch.cuda.synchronize()
    out = [o.cpu().numpy() for o in out]
    wb, wd = decode_ref(anchors, reg, cls, rdim)
    det = filter_detections_ref(wb, wd, cls)
    poll = c_oracle.fit_road_planes_c(det[0], det[1], det[4], P_inv, planes)
    for g, w in zip(out, det + poll):
        assert g.shape == w.shape and g.dtype == w.dtype and np.array_equal(g, w, equal_nan=True)
    assert out[0].shape == (2, 100, 12) and out[5].shape == (2, 100, 4, 3) and out[7].shape == (2, 100)
