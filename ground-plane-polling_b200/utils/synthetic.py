"""
Seeded synthetic KITTI-shaped detections -- the workload generator for bench.py,
smoke() and the parity tests.  No dataset and no network are needed: cars are placed on a plane drawn
from the road-plane database, their 3-D boxes are built with the reference's corner convention
(/root/reference/label_prep/computeBox3D.m:12-30), the (l, m, r, t) key-points and the orientation class
follow /root/reference/label_prep/create_mod_labels.m:57-101, and everything is projected with a
KITTI camera-2 matrix scaled like /root/reference/keras_retinanet_3D/bin/run_network.py:48-59
(image 1242x375 resized by 1333/1242, keras_retinanet_3D/utils/image.py:174-200).

Output layout equals what FilterDetections hands to FitRoadPlanes
(/root/reference/keras_retinanet_3D/layers/filter_detections.py:170-185): exactly D rows per image,
rows past ``n_valid`` padded with -1.
"""
import numpy as np

KITTI_IMAGE_WH = (1242, 375)
KITTI_SCALE = 1333.0 / 1242.0
KITTI_P2 = np.array([[721.5377, 0.0, 609.5593, 44.85728],
                     [0.0, 721.5377, 172.854, 0.2163791],
                     [0.0, 0.0, 1.0, 0.002745884]], dtype=np.float64)
DIMS_MEAN = np.array([1.6570, 1.7999, 4.2907])   # (h, w, l), layers/_misc.py:168
DIMS_STD = np.array([0.2681, 0.2243, 0.6281])    # layers/_misc.py:170

# 0-based corner ids of (l, m, r, t) per orientation class, create_mod_labels.m:57-101
KEYPOINT_CORNERS = np.array([[2, 1, 0, 5],
                             [1, 0, 3, 4],
                             [3, 2, 1, 6],
                             [0, 3, 2, 7]])


def kitti_calibration(scale=KITTI_SCALE, P2=KITTI_P2):
    """(P, P_inv) float64 exactly as ``load_calibration`` builds them (run_network.py:56-58)."""
    S = np.array([[scale, 0.0, 0.0], [0.0, scale, 0.0], [0.0, 0.0, 1.0]])
    P = S.dot(np.asarray(P2, dtype=np.float64).reshape(3, 4))
    return P, np.linalg.pinv(P)


def _normalise(planes):
    p = np.asarray(planes, dtype=np.float64)
    p = p * (-np.sign(p[:, 1:2]))
    return p / np.linalg.norm(p[:, :3], axis=1, keepdims=True)


def synth_detections(B, D, planes, seed, n_valid=None, kp_noise_px=1.5, dim_noise=0.05,
                     return_truth=False):
    """Synthetic FilterDetections output for B images.

    Returns boxes (B, D, 12) float32, dimensions (B, D, 3) float32, orientations (B, D) int32,
    P_inv (B, 4, 3) float64 (callers of the reference pass float64, run_network.py:105)
    [, truth dict with the generating plane index / location / yaw / dims].
    """
    rng = np.random.default_rng(seed)
    n_valid = D if n_valid is None else int(n_valid)
    P, P_inv = kitti_calibration()
    pn = _normalise(planes)
    ok = np.where((pn[:, 3] > 1.0) & (pn[:, 3] < 2.5) & (np.abs(pn[:, 1]) > 0.9))[0]
    if ok.size == 0:
        ok = np.arange(pn.shape[0])
    M = B * D
    pid = ok[rng.integers(0, ok.size, size=M)]
    a, b, c, d = pn[pid, 0], pn[pid, 1], pn[pid, 2], pn[pid, 3]
    dims = DIMS_MEAN + 0.5 * DIMS_STD * rng.standard_normal((M, 3))
    h, w, l = dims[:, 0], dims[:, 1], dims[:, 2]
    z = rng.uniform(6.0, 50.0, size=M)
    # keep the car inside the field of view: |x| limited by the frustum at depth z (and by 15 m)
    half_fov = 0.5 * KITTI_IMAGE_WH[0] / KITTI_P2[0, 0]
    xlim = np.minimum(15.0, 0.85 * half_fov * z)
    x = rng.uniform(-1.0, 1.0, size=M) * xlim
    y = -(a * x + c * z + d) / b
    ry = rng.uniform(-np.pi, np.pi, size=M)

    xc = np.stack([l / 2, l / 2, -l / 2, -l / 2, l / 2, l / 2, -l / 2, -l / 2], axis=1)
    yc = np.stack([0 * h, 0 * h, 0 * h, 0 * h, -h, -h, -h, -h], axis=1)
    zc = np.stack([w / 2, -w / 2, -w / 2, w / 2, w / 2, -w / 2, -w / 2, w / 2], axis=1)
    cr, sr = np.cos(ry)[:, None], np.sin(ry)[:, None]
    X = cr * xc + sr * zc + x[:, None]
    Y = yc + y[:, None]
    Z = -sr * xc + cr * zc + z[:, None]
    hom = np.stack([X, Y, Z, np.ones_like(X)], axis=-1)            # (M, 8, 4)
    pix = hom @ P.T                                                # (M, 8, 3)
    uv = pix[..., :2] / pix[..., 2:3]                              # (M, 8, 2)

    alpha = ry - np.arctan2(x, z)
    alpha = (alpha + np.pi) % (2 * np.pi) - np.pi
    adeg = np.degrees(alpha)
    o = np.where((adeg >= 0) & (adeg < 90), 0,
                 np.where(adeg >= 90, 1, np.where(adeg >= -90, 2, 3))).astype(np.int32)
    kp_ids = KEYPOINT_CORNERS[o]                                   # (M, 4)
    kp = np.take_along_axis(uv, kp_ids[:, :, None], axis=1)        # (M, 4, 2)
    kp = kp + kp_noise_px * rng.standard_normal(kp.shape)
    x1y1 = uv.min(axis=1)
    x2y2 = uv.max(axis=1)
    boxes = np.concatenate([x1y1, x2y2, kp.reshape(M, 8)], axis=1)
    dims_noisy = dims * (1.0 + dim_noise * rng.standard_normal(dims.shape))

    boxes = boxes.reshape(B, D, 12).astype(np.float32)
    dims_out = dims_noisy.reshape(B, D, 3).astype(np.float32)
    orient = o.reshape(B, D).copy()
    if n_valid < D:                                               # FilterDetections pads with -1
        boxes[:, n_valid:] = -1.0
        dims_out[:, n_valid:] = -1.0
        orient[:, n_valid:] = -1
    P_inv_b = np.ascontiguousarray(np.broadcast_to(P_inv, (B, 4, 3)))
    if return_truth:
        truth = dict(plane_index=pid.reshape(B, D), location=np.stack([x, y, z], 1).reshape(B, D, 3),
                     ry=ry.reshape(B, D), dims=dims.reshape(B, D, 3), P=P)
        return boxes, dims_out, orient, P_inv_b, truth
    return boxes, dims_out, orient, P_inv_b
