"""GPU: pose recovery / KITTI-record kernels against the reference driver's own lines (tests/golden/pose_*.npz, made by
exec'ing run_network.py:113-330) and against the restated loop (oracle/pose_ref.py, pinned to the same vectors)."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, load_planes
from gpp_b200.utils import synthetic
from oracle import c_oracle
from oracle.pose_ref import kitti_yaw, pose_ref

pytestmark = pytest.mark.gpu

RTOL = 1e-4   # north_star tolerance for location, dimensions and yaw


def _wrap(a):
    return (a + np.pi) % (2 * np.pi) - np.pi


def test_pose_matches_reference_loop(gpp):
    planes = load_planes('1k')
    boxes, dims, orient, P_inv = synthetic.synth_detections(4, 100, planes, seed=91, n_valid=90)
    kp, kpl, res = c_oracle.fit_road_planes_c(boxes, dims, orient, P_inv, planes)
    kp2 = kp.reshape(-1, 12)
    o = orient.reshape(-1)
    loc, ang, dout = gpp.recover_pose(kp2, dims.reshape(-1, 3), o)
    wloc, wang, wdims = pose_ref(kp2, dims.reshape(-1, 3).copy(), o)
    valid = (o >= 0) & np.isfinite(wang).all(1) & np.isfinite(wloc).all(1)
    assert valid.sum() > 300
    scale = np.abs(wloc[valid]).max(axis=1, keepdims=True)
    assert np.all(np.abs(loc[valid] - wloc[valid]) <= RTOL * scale + 1e-6)
    assert np.allclose(dout[valid], wdims[valid], rtol=RTOL, atol=0)
    # Rodrigues vectors: compare as rotations (axis*angle), then the KITTI yaw
    assert np.all(np.abs(ang[valid] - wang[valid]) <= RTOL * np.maximum(1.0, np.abs(wang[valid]).max(1, keepdims=True)))
    assert np.all(np.abs(_wrap(kitti_yaw(ang[valid]) - kitti_yaw(wang[valid]))) <= RTOL * np.pi)
    # padding rows untouched (zeros from the wrapper)
    assert np.all(loc[o < 0] == 0) and np.all(ang[o < 0] == 0)


def test_pose_device_entry(gpp):
    import torch
    planes = load_planes('100')
    boxes, dims, orient, P_inv = synthetic.synth_detections(2, 50, planes, seed=92)
    dev = torch.device('cuda', 0)
    out = gpp.fit_road_planes_torch(torch.from_numpy(boxes).to(dev), torch.from_numpy(dims).to(dev),
                                    torch.from_numpy(orient).to(dev), torch.from_numpy(P_inv).to(dev), planes)
    loc, ang, dout = gpp.recover_pose_torch(out[0], torch.from_numpy(dims).to(dev), torch.from_numpy(orient).to(dev))
    torch.cuda.synchronize()
    hl, ha, hd = gpp.recover_pose(out[0].cpu().numpy(), dims.reshape(-1, 3), orient.reshape(-1))
    assert np.array_equal(loc.cpu().numpy(), hl) and np.array_equal(ang.cpu().numpy(), ha)
    assert np.array_equal(dout.cpu().numpy(), hd)


def test_kitti_records_match_reference_writer(gpp):
    from oracle.kitti_ref import kitti_records_ref
    planes = load_planes('1k')
    boxes, dims, orient, P_inv = synthetic.synth_detections(3, 100, planes, seed=93)
    kp, kpl, res = c_oracle.fit_road_planes_c(boxes, dims, orient, P_inv, planes)
    wloc, wang, wdims = pose_ref(kp.reshape(-1, 12), dims.reshape(-1, 3).copy(), orient.reshape(-1))
    ok = np.isfinite(wang).all(1) & np.isfinite(wloc).all(1)
    got = gpp.kitti_records(wloc[ok], wang[ok], wdims[ok])
    want = kitti_records_ref(wloc[ok], wang[ok], wdims[ok])
    assert got.shape == want.shape
    # angles (alpha, r_y) compared on the circle, lengths (h, Y) relatively
    assert np.all(np.abs(_wrap(got[:, 0] - want[:, 0])) <= RTOL * np.pi)
    assert np.all(np.abs(_wrap(got[:, 3] - want[:, 3])) <= RTOL * np.pi)
    assert np.allclose(got[:, 1], want[:, 1], rtol=RTOL, atol=1e-5)
    assert np.allclose(got[:, 2], want[:, 2], rtol=RTOL, atol=1e-4)


def test_postprocess_image_mirrors_the_driver(gpp):
    """run_network.py:113-287 for one image: score filter + sort, unscale, pose, KITTI record."""
    from oracle.kitti_ref import kitti_records_ref
    planes = load_planes('1k')
    rng = np.random.default_rng(5)
    boxes, dims, orient, P_inv = synthetic.synth_detections(1, 100, planes, seed=94, n_valid=40)
    scores = np.where(np.arange(100) < 40, rng.uniform(0.0, 1.0, 100), -1.0).astype(np.float32)
    labels = np.where(np.arange(100) < 40, 0, -1).astype(np.int32)
    kp, kpl, res = gpp.fit_road_planes(boxes, dims, orient, P_inv, planes)
    scale = 1333.0 / 1242.0
    out = gpp.postprocess_image(boxes[0], dims[0], scores, labels, orient[0], kp[0], kpl[0], res[0], scale)
    keep = np.where(scores > 0.05)[0]
    keep = keep[np.argsort(-scores[keep])][:100]
    assert np.array_equal(out['scores'], scores[keep])
    assert np.allclose(out['boxes'], boxes[0][keep][:, :4] / np.float32(scale))
    wloc, wang, wdims = pose_ref(kp[0].reshape(-1, 12)[keep], dims[0][keep].copy(), orient[0][keep])
    assert np.allclose(out['locations'], wloc, rtol=1e-4, atol=1e-4)
    assert np.allclose(out['dimensions'], wdims, rtol=1e-4, atol=0)
    want = kitti_records_ref(wloc, wang, wdims)
    assert np.allclose(out['kitti'][:, 1:3], want[:, 1:3], rtol=1e-4, atol=1e-4)
    lines = gpp.kitti.format_kitti_lines(out['boxes'], out['dimensions'], out['locations'], out['scores'],
                                         out['kitti'], (1242, 375))
    assert len(lines) == len(keep) and lines[0].startswith('Car -1 -1 ')


def test_return_pose_extension(gpp):
    """fit_road_planes(..., return_pose=True) appends what recover_pose gives for every row; the default return list
    is unchanged; numpy and torch entries agree bit for bit."""
    import torch
    planes = load_planes('1k')
    boxes, dims, orient, P_inv = synthetic.synth_detections(3, 40, planes, seed=93, n_valid=31)
    base = gpp.fit_road_planes(boxes, dims, orient, P_inv, planes)
    assert len(base) == 3
    full = gpp.fit_road_planes(boxes, dims, orient, P_inv, planes, return_index=True, return_pose=True)
    assert len(full) == 7 and all(np.array_equal(a, b, equal_nan=True) for a, b in zip(full[:3], base))
    loc, ang, dout = gpp.recover_pose(base[0].reshape(-1, 12), dims.reshape(-1, 3), orient.reshape(-1))
    assert full[4].shape == full[5].shape == full[6].shape == (3, 40, 3)
    assert np.array_equal(full[4].reshape(-1, 3), loc, equal_nan=True)
    assert np.array_equal(full[5].reshape(-1, 3), ang, equal_nan=True)
    assert np.array_equal(full[6].reshape(-1, 3), dout, equal_nan=True)
    dev = torch.device('cuda', 0)
    t = gpp.fit_road_planes_torch(torch.from_numpy(boxes).to(dev), torch.from_numpy(dims).to(dev),
                                  torch.from_numpy(orient).to(dev), torch.from_numpy(P_inv.astype(np.float32)).to(dev),
                                  planes, return_pose=True)
    assert len(t) == 6
    for a, b in zip(t[3:], full[4:]):
        assert np.array_equal(a.cpu().numpy(), b, equal_nan=True)


POSE_CASES = sorted(f[:-4] for f in os.listdir(GOLDEN) if f.startswith('pose_') and f.endswith('.npz'))


def _rot(rv):
    import cv2
    return np.stack([cv2.Rodrigues(np.asarray(v, np.float64))[0] for v in rv]) if len(rv) else np.zeros((0, 3, 3))


@pytest.mark.parametrize('name', POSE_CASES)
def test_pose_and_kitti_kernels_equal_the_reference_lines(gpp, name):
    """locations, dimensions and yaw within 1e-4 relative (north_star) of what run_network.py:137-287 / :297-327
    themselves compute -- all four orientation branches, near-identity and near-pi rotations"""
    g = dict(np.load(os.path.join(GOLDEN, name + '.npz')))
    kp, dims, orient = g['select_keypoints'], g['select_dimensions'], g['select_orientations']
    loc, ang, dout = gpp.recover_pose(kp, dims, orient)
    wloc, wang, wdim = g['out_locations'], g['out_angles'], g['out_dimensions']
    assert loc.shape == wloc.shape and ang.shape == wang.shape and dout.shape == wdim.shape
    if len(kp) == 0:
        return
    scale = np.maximum(np.abs(wloc).max(axis=1, keepdims=True), 1.0)
    assert np.all(np.abs(loc - wloc) <= RTOL * scale)
    assert np.allclose(dout, wdim, rtol=RTOL, atol=0)
    # the Rodrigues vector as a rotation (near pi the vector may flip sign: same rotation), then the KITTI yaw
    assert np.abs(_rot(ang) - _rot(wang)).max() <= 2e-4
    near_pi = np.linalg.norm(wang, axis=1) > np.pi - 2e-3
    assert np.all(np.abs(ang[~near_pi] - wang[~near_pi]) <= RTOL * np.maximum(1.0, np.abs(wang[~near_pi]).max(1, keepdims=True)))
    assert np.all(np.abs(_wrap(kitti_yaw(ang[~near_pi]) - kitti_yaw(wang[~near_pi]))) <= RTOL * np.pi)
    rec = gpp.kitti_records(wloc, wang, wdim)
    want = g['kitti_rec']
    assert np.all(np.abs(_wrap(rec[:, 0] - want[:, 0])) <= RTOL * np.pi)
    assert np.all(np.abs(_wrap(rec[:, 3] - want[:, 3])) <= RTOL * np.pi)
    assert np.allclose(rec[:, 1], want[:, 1], rtol=RTOL, atol=1e-5)
    assert np.allclose(rec[:, 2], want[:, 2], rtol=RTOL, atol=1e-4)


@pytest.mark.parametrize('name', POSE_CASES)
def test_postprocess_image_equals_the_reference_driver(gpp, name):
    """one image through the product's driver tail (selection, unscale, pose, KITTI lines) against the reference's
    `outputs` dict (:291) and the text its KITTI writer produced (:295-330)"""
    g = dict(np.load(os.path.join(GOLDEN, name + '.npz')))
    out = gpp.postprocess_image(g['in_boxes'][0], g['in_dimensions'][0], g['in_scores'][0], g['in_labels'][0],
                                g['in_orientations'][0], g['in_keypoints'][0], g['in_keyplanes'][0],
                                g['in_residuals'][0], float(g['scale']))
    for k in ('boxes', 'keypoints', 'labels', 'scores', 'residuals'):
        assert np.array_equal(out[k], g['out_' + k]), k                     # selection / unscale: exact
    n = len(out['scores'])
    if n == 0:
        return
    assert np.allclose(out['locations'], g['out_locations'], rtol=RTOL, atol=1e-4)
    assert np.allclose(out['dimensions'], g['out_dimensions'], rtol=RTOL, atol=0)
    lines = gpp.kitti.format_kitti_lines(out['boxes'], out['dimensions'], out['locations'], out['scores'],
                                         out['kitti'], (1242, 375))
    want_lines = str(g['kitti_lines']).splitlines()
    assert len(lines) == len(want_lines) == n
    near_pi = np.linalg.norm(g['out_angles'], axis=1) > np.pi - 2e-3
    for a, b, skip in zip(lines, want_lines, near_pi):
        fa, fb = a.split(), b.split()
        assert fa[:3] == fb[:3] == ['Car', '-1', '-1']
        va, vb = np.array(fa[3:], dtype=np.float64), np.array(fb[3:], dtype=np.float64)
        if skip:
            continue
        d = np.abs(va - vb)
        d[[0, 11]] = np.abs(_wrap(d[[0, 11]]))                              # alpha, r_y live on the circle
        assert np.all(d <= 0.0100001 + 1e-4 * np.abs(vb))                   # one unit of the last printed place


def test_polled_keypoints_feed_the_second_caller(gpp):
    """utils/eval.py:96-118 (`_get_detections`): the GPU path's (1, 100, 4, 3) / (1, 100, 1, 4) outputs go through the
    reshapes of the second caller and give the reference's 34-column table bit for bit"""
    g = dict(np.load(os.path.join(GOLDEN, 'pose_main1.npz')))
    planes = load_planes(str(g['poll_planes_db']))
    kp, kpl, res = gpp.fit_road_planes(g['in_boxes'], g['in_dimensions'], g['in_orientations'], g['poll_P_inv'],
                                       np.expand_dims(planes, axis=0))          # the callers' (1, N, 4) feed
    assert np.array_equal(kp, g['in_keypoints']) and np.array_equal(kpl, g['in_keyplanes'])
    assert np.array_equal(res, g['in_residuals'])
    det = gpp.image_detections(g['in_boxes'], g['in_dimensions'], g['in_scores'], g['in_labels'], g['in_orientations'],
                               kp, kpl)
    assert np.array_equal(det, g['eval_detections'], equal_nan=True)
