"""
Adversarial inputs for the VERIFIED mode: detections and plane databases far outside the geometry of the benchmark
generator (utils/synthetic.py), used by tests/test_adversarial_gpu.py and scripts/soak_adversarial.py to press on the
error margin the fast filter relies on (csrc/gpp_poll2.cuh).  The reference's operator accepts any float32 input
(keras_retinanet_3D/layers/fit_road_planes.py:49-61 has no validation), so 'verified' has to equal 'exact' on all of it.

Flavours of ``detections``:
  kitti     key-points of a real box under KITTI's camera with heavy pixel noise
  wild      four independent random pixels, dimensions from 1 cm to 100 m, any orientation class (incl. -1 and 7)
  near      objects 0.3 - 3 m from the camera (huge parallax), far: 80 - 2000 m (t_k -> 0, margins quadratic in depth)
  collapse  two or three key-points coincide or differ by one ulp (distances cancel), collinear key-points
  pinv      a random dense P_inv (rays in any direction, z of either sign), scaled by 1e-6 .. 1e6
  horizon   key-points on the vanishing line of a database plane (n . d_k within rounding noise of 0)
Flavours of ``planes``:
  road      the shipped databases (near-vertical normals)            steep   |b| down to 1e-4, any normal direction
  scale     |n| and d from 1e-3 to 1e3, un-normalised                degenerate  b = 0, all-zero rows, duplicates, d = 0,
                                                                              denormal and huge entries
"""
import numpy as np

from . import synthetic

__all__ = ['DET_FLAVOURS', 'PLANE_FLAVOURS', 'detections', 'planes']

DET_FLAVOURS = ('kitti', 'wild', 'near', 'far', 'collapse', 'pinv', 'horizon')
PLANE_FLAVOURS = ('road', 'steep', 'scale', 'degenerate')


def _kitti_pinv(n_img, rng, jitter=0.0):
    s = 1333.0 / 1242.0
    out = np.empty((n_img, 4, 3), np.float32)
    for b in range(n_img):
        P = np.dot(np.diag([s, s, 1.0]), synthetic.KITTI_P2)
        if jitter:
            P = P * (1.0 + jitter * rng.standard_normal(P.shape))
        out[b] = np.linalg.pinv(P)
    return out


def planes(flavour, n, rng, base=None):
    """(n, 4) float32 raw plane database of one flavour (``base``: a shipped database for 'road')."""
    if flavour == 'road':
        idx = rng.choice(base.shape[0], size=n, replace=n > base.shape[0])
        return np.ascontiguousarray(base[np.sort(idx)], dtype=np.float32)
    nrm = rng.standard_normal((n, 3))
    nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    d = rng.uniform(0.5, 3.0, (n, 1)) * rng.choice([-1.0, 1.0], (n, 1))
    p = np.concatenate([nrm, d], axis=1)
    if flavour == 'steep':
        k = n // 2
        p[:k, 1] = 10.0 ** rng.uniform(-4, -0.5, k) * rng.choice([-1.0, 1.0], k)
        p[:k, :3] /= np.linalg.norm(p[:k, :3], axis=1, keepdims=True)
    elif flavour == 'scale':
        p *= 10.0 ** rng.uniform(-3, 3, (n, 1))
        p[:, 3] *= 10.0 ** rng.uniform(-2, 2, n)
    elif flavour == 'degenerate':
        k = max(1, n // 16)
        p[0 * k:1 * k, 1] = 0.0                               # b = 0: all-zero then NaN after normalisation
        p[1 * k:2 * k] = 0.0
        p[2 * k:3 * k] = p[3 * k:4 * k]                       # duplicates
        p[4 * k:5 * k, 3] = 0.0                               # through the camera centre
        p[5 * k:6 * k] *= 1e-40                               # denormal after the float32 cast
        p[6 * k:7 * k] *= 1e30
        p[7 * k:8 * k, 3] *= 1e-30
    return np.ascontiguousarray(p, dtype=np.float32)


def detections(flavour, n_img, n_det, rng, base_planes):
    """boxes (B, D, 12), dimensions (B, D, 3), orientations (B, D) int32, P_inv (B, 4, 3) -- float32."""
    boxes, dims, orient, P_inv = synthetic.synth_detections(n_img, n_det, base_planes, seed=int(rng.integers(1 << 30)),
                                                            kp_noise_px=float(rng.choice([0.0, 1.5, 6.0, 25.0])))
    boxes, dims, orient = boxes.copy(), dims.copy(), orient.copy()
    P_inv = P_inv.astype(np.float32)
    shape = (n_img, n_det)
    if flavour == 'wild':
        boxes[..., 4:12] = rng.uniform(-300, 1700, shape + (8,))
        dims[...] = 10.0 ** rng.uniform(-2, 2, shape + (3,))
        orient[...] = rng.choice([-1, 0, 1, 2, 3, 7], shape)
    elif flavour in ('near', 'far'):
        # rescale the key-point pattern around the principal point: same box seen from much closer / farther
        c = np.array([609.5593 * 1333 / 1242, 172.854 * 1333 / 1242], np.float32)
        f = 10.0 ** rng.uniform(0.5, 1.3, shape + (1, 1)) if flavour == 'near' else 10.0 ** rng.uniform(-2.5, -0.7, shape + (1, 1))
        kp = boxes[..., 4:12].reshape(shape + (4, 2))
        boxes[..., 4:12] = ((kp - kp.mean(axis=-2, keepdims=True)) * f + c +
                            rng.uniform(-200, 200, shape + (1, 2))).reshape(shape + (8,))
    elif flavour == 'collapse':
        kp = boxes[..., 4:12].reshape(shape + (4, 2)).copy()
        which = rng.integers(0, 5, shape)
        kp[which == 0, 0] = kp[which == 0, 1]                                   # l = m
        kp[which == 1, 2] = np.nextafter(kp[which == 1, 1], np.float32(np.inf))   # r one ulp from m
        kp[which == 2, 3] = kp[which == 2, 1]                                   # t = m
        sel = which == 3                                                        # collinear l, m, r
        kp[sel, 2] = 2 * kp[sel, 1] - kp[sel, 0]
        sel = which == 4                                                        # all four within a pixel
        kp[sel] = kp[sel][:, :1] + rng.uniform(-0.5, 0.5, (int(sel.sum()), 4, 2))
        boxes[..., 4:12] = kp.reshape(shape + (8,))
    elif flavour == 'pinv':
        P_inv = (rng.standard_normal((n_img, 4, 3)) * 10.0 ** rng.uniform(-6, 6, (n_img, 1, 1))).astype(np.float32)
        k = n_img // 2
        P_inv[:k] = _kitti_pinv(k, rng, jitter=0.3) * (10.0 ** rng.uniform(-6, 6, (k, 1, 1))).astype(np.float32)
    elif flavour == 'horizon':
        # move the ground key-points onto (almost) the vanishing line of a database plane: n . d_k ~ 0
        pl = base_planes[rng.integers(0, base_planes.shape[0], n_img)].astype(np.float64)
        for b in range(n_img):
            line = P_inv[b, :3, :].astype(np.float64).T @ pl[b, :3]             # image line of the pixels with n . d = 0
            u = boxes[b, :, 4:10].reshape(n_det, 3, 2)[..., 0].astype(np.float64)
            v = -(line[0] * u + line[2]) / line[1]
            v = v + rng.choice([0.0, 1e-3, 0.05, 1.0], (n_det, 1)) * rng.standard_normal((n_det, 3))
            boxes[b, :, 5:10:2] = v.astype(np.float32)
    return (np.ascontiguousarray(boxes, np.float32), np.ascontiguousarray(dims, np.float32),
            np.ascontiguousarray(orient, np.int32), np.ascontiguousarray(P_inv, np.float32))
