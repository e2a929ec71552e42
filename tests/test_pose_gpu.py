"""GPU: pose recovery kernel against the restated reference loop (oracle/pose_ref.py, cv2.Rodrigues)."""
import numpy as np
import pytest

from conftest import load_planes
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
