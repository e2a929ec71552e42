"""CPU: the pose / KITTI-record / driver-selection oracles (oracle/pose_ref.py, kitti_ref.py, driver_ref.py) against
vectors produced by EXECUTING the reference driver's own source lines (tests/golden/make_golden_pose.py cuts
run_network.py:48-59, :113-135, :137-287, :291, :295-330 and utils/eval.py:96-118 out of the reference files and
exec's them).  With this the oracles the CUDA pose path is compared with are pinned to the reference itself."""
import os

import numpy as np
import pytest

from conftest import GOLDEN
from oracle.driver_ref import driver_image_ref
from oracle.kitti_ref import kitti_records_ref
from oracle.pose_ref import pose_ref

CASES = sorted(f[:-4] for f in os.listdir(GOLDEN) if f.startswith('pose_') and f.endswith('.npz'))
IN_KEYS = ('boxes', 'dimensions', 'scores', 'labels', 'orientations', 'keypoints', 'keyplanes', 'residuals')


def load_case(name):
    return dict(np.load(os.path.join(GOLDEN, name + '.npz')))


def test_the_golden_sets_cover_the_branches():
    assert set(CASES) >= {'pose_main0', 'pose_main1', 'pose_main2', 'pose_identity', 'pose_flip', 'pose_yaw', 'pose_empty'}
    seen = set()
    for name in CASES:
        g = load_case(name)
        seen |= set(int(o) for o in g['select_orientations'])
    assert seen == {0, 1, 2, 3}                       # every reachable branch of run_network.py:147-247
    assert len(load_case('pose_empty')['out_scores']) == 0
    ang = load_case('pose_identity')['out_angles']
    assert np.any(np.linalg.norm(ang, axis=1) < 1e-6) and np.any(np.linalg.norm(ang, axis=1) > 1e-3)
    assert np.any(np.linalg.norm(load_case('pose_flip')['out_angles'], axis=1) > 3.1)


@pytest.mark.parametrize('name', CASES)
def test_driver_oracle_equals_the_reference_lines(name):
    """selection (:113-135), pose loop (:137-287), outputs dict (:291), KITTI lines (:295-330): bit for bit"""
    g = load_case(name)
    outs = [g['in_' + k] for k in IN_KEYS]
    outputs, lines = driver_image_ref(outs, float(g['scale']), (1242, 375))
    for k in ('boxes', 'keypoints', 'labels', 'scores', 'locations', 'angles', 'dimensions', 'residuals'):
        want = g['out_' + k]
        assert outputs[k].shape == want.shape and outputs[k].dtype == want.dtype, k
        assert np.array_equal(outputs[k], want, equal_nan=True), k
    assert ''.join(lines) == str(g['kitti_lines'])


@pytest.mark.parametrize('name', CASES)
def test_pose_and_kitti_oracles_equal_the_reference_lines(name):
    g = load_case(name)
    kp, dims, orient = g['select_keypoints'], g['select_dimensions'].copy(), g['select_orientations']
    loc, ang, dout = pose_ref(kp, dims, orient)
    assert np.array_equal(loc, g['out_locations'], equal_nan=True)
    assert np.array_equal(ang, g['out_angles'], equal_nan=True)
    assert np.array_equal(dout, g['out_dimensions'], equal_nan=True)
    rec = kitti_records_ref(g['out_locations'], g['out_angles'], g['out_dimensions'])
    assert np.array_equal(rec, g['kitti_rec'], equal_nan=True)


def test_calibration_mirror_equals_the_reference_function(tmp_path, gpp):
    """utils/calibration.load_calibration against run_network.py:48-59 executed on the same file"""
    g = dict(np.load(os.path.join(GOLDEN, 'driver_calib.npz')))
    path = tmp_path / '000001.txt'
    path.write_text(str(g['calib_text']))
    P, P_inv = gpp.load_calibration(str(path), float(g['scale']))
    assert np.array_equal(P, g['P']) and np.array_equal(P_inv, g['P_inv'])


@pytest.mark.parametrize('name', CASES)
def test_host_selection_mirrors_equal_the_reference_lines(name, gpp):
    """the numpy-only host helpers of the product (no GPU involved): run_network.py:117-125 and utils/eval.py:96-118"""
    g = load_case(name)
    keep = gpp.select_detections(g['in_scores'][0])
    assert np.array_equal(g['in_scores'][0][keep], g['select_scores'])
    assert np.array_equal(g['in_residuals'][0][keep], g['select_residuals'])
    det = gpp.image_detections(g['in_boxes'], g['in_dimensions'], g['in_scores'], g['in_labels'], g['in_orientations'],
                               g['in_keypoints'], g['in_keyplanes'])
    assert det.shape == g['eval_detections'].shape and det.shape[1] == 34
    assert np.array_equal(det, g['eval_detections'], equal_nan=True)
