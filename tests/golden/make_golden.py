"""
Generate the golden vectors under tests/golden/*.npz by EXECUTING THE REFERENCE'S OWN FILE
(/root/reference/keras_retinanet_3D/layers/fit_road_planes.py, unmodified) over the numpy stand-ins of
tf_numpy_shim.py.  Run in the build container only (the reference tree does not travel to the GPU box):

    python tests/golden/make_golden.py

Each .npz holds the inputs exactly as a reference caller would feed them (float32 boxes/dimensions, int32
orientations, float64 P_inv and planes -- Keras casts those to float32 at feed, run_network.py:105) and the
three outputs of ``FitRoadPlanes.call`` (keypoints, keyplanes, residuals).  Cases that use a shipped plane
database store its tag (``planes_db``) instead of the values.
"""
import importlib.util
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)

import tf_numpy_shim  # noqa: E402


def _load(path, name):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


synthetic = _load(os.path.join(ROOT, 'ground-plane-polling_b200', 'utils', 'synthetic.py'), 'synthetic')


def shipped(tag):
    return np.load(os.path.join(ROOT, 'road_planes_database', 'road_planes_database_%s.npy' % tag))


def edge_case():
    """Hand-built inputs that hit every special branch of the selection logic:
       duplicate planes (lowest index must win), a plane exactly parallel to a key-point ray
       (n.d == 0 -> inf/NaN), a plane with b == 0 (normalises to NaN), padding rows (-1), a detection whose
       every plane fails the z-check (all-sentinel), per-image DIFFERENT plane sets (planes (B, N, 4))."""
    rng = np.random.default_rng(7)
    base = shipped('100')[:24].copy()
    planes0 = base.copy()
    planes0[5] = planes0[3]                       # exact duplicates
    planes0[17] = planes0[3]
    planes0[9] = [1.0, -2.0, 0.0, 1.7]            # parallel to the ray (2,1,1) of detection 0
    planes0[11] = [0.3, 0.0, 0.9, 1.5]            # b == 0 -> direction 0 -> 0/0 = NaN plane
    planes1 = base[::-1].copy()                   # image 1 sees a different DB
    planes = np.stack([planes0, planes1], axis=0)
    boxes, dims, orient, P_inv = synthetic.synth_detections(2, 12, shipped('100'), seed=11, n_valid=9)
    # image 0 uses a trivial P_inv so that ray m of detection 0 is exactly (2, 1, 1)
    P_inv = P_inv.copy()
    P_inv[0] = np.array([[1.0, 0, 0], [0, 1.0, 0], [0, 0, 1.0], [0, 0, 0]])
    boxes[0, :, 4:] = (boxes[0, :, 4:] - np.tile([666.0, 186.0], 4)) / 774.0   # roughly normalised coords
    boxes[0, 0, 6:8] = [2.0, 1.0]
    # detection 1 of image 1: swap l and r key-points -> z_dir_check < 0 for (nearly) every plane
    boxes[1, 1, 4:6], boxes[1, 1, 8:10] = boxes[1, 1, 8:10].copy(), boxes[1, 1, 4:6].copy()
    # detection 2 of image 1: absurd dimensions -> no votes anywhere
    dims[1, 2] = [40.0, 50.0, 60.0]
    del rng
    return dict(boxes=boxes, dimensions=dims, orientations=orient, P_inv=P_inv, planes=planes)


def cases():
    out = {}
    for name, (B, D, tag, seed, n_valid) in {
        'c1_1x20x10': (1, 20, '10', 1, None),
        'pad_2x100x100': (2, 100, '100', 2, 60),
        'c2_1x100x1k': (1, 100, '1k', 2, None),
        'dup_1x16x10k': (1, 16, '10k', 3, None),
        'n22k_1x12x22k': (1, 12, '22k', 4, None),
    }.items():
        planes = shipped(tag)
        boxes, dims, orient, P_inv = synthetic.synth_detections(B, D, planes, seed=seed, n_valid=n_valid)
        out[name] = dict(boxes=boxes, dimensions=dims, orientations=orient, P_inv=P_inv, planes_db=tag)
    out['edge_2x12x24'] = edge_case()
    # every plane degenerate (b == 0 -> NaN plane): nothing compares below FLT_MAX -> index 0, NaN outputs
    boxes, dims, orient, P_inv = synthetic.synth_detections(1, 3, shipped('10'), seed=5)
    out['allnan_1x3x2'] = dict(boxes=boxes, dimensions=dims, orientations=orient, P_inv=P_inv,
                               planes=np.array([[[0.3, 0.0, 0.9, 1.5], [0.1, 0.0, 0.2, 1.0]]]))
    return out


def run_reference(ref, case):
    f32 = np.float32
    if 'planes_db' in case:
        planes = shipped(str(case['planes_db']))
        B = case['boxes'].shape[0]
        # every reference caller feeds the DB once per image (run_network.py:105, preprocessing/kitti.py:220)
        planes = np.tile(planes[None], (B, 1, 1))
    else:
        planes = case['planes']
    inputs = [case['boxes'].astype(f32), case['dimensions'].astype(f32), case['orientations'].astype(np.int32),
              case['P_inv'].astype(f32), planes.astype(f32)]
    layer = ref.FitRoadPlanes()
    keypoints, keyplanes, residuals = layer.call(inputs)
    shapes = layer.compute_output_shape([x.shape for x in inputs])
    assert [tuple(s) for s in shapes] == [keypoints.shape, keyplanes.shape, residuals.shape]
    return keypoints.astype(f32), keyplanes.astype(f32), residuals.astype(f32)


if __name__ == '__main__':
    ref = tf_numpy_shim.load_reference_module()
    for name, case in cases().items():
        kp, kpl, res = run_reference(ref, case)
        np.savez_compressed(os.path.join(HERE, name + '.npz'), keypoints=kp, keyplanes=kpl, residuals=res, **case)
        print('%-16s keypoints %s  nan rows %d  sentinel rows %d' % (
            name, kp.shape, int(np.isnan(res).sum()), int((res == np.float32(100.0) / np.float32(6.0)).sum())))
