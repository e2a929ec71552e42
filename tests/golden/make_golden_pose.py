"""
Generates tests/golden/pose_*.npz and driver_calib.npz by EXECUTING the reference driver's own source lines (run in
the build container, where /root/reference exists; the GPU box only sees the committed vectors).

/root/reference/keras_retinanet_3D/bin/run_network.py cannot be imported (it imports keras / tensorflow at the top),
but the per-image work after ``model.predict_on_batch`` is plain numpy + cv2.  The lines are cut out of the file by
line number, de-indented and exec'd in a namespace that holds nothing but numpy, cv2 and the synthetic stand-ins of
the network outputs -- no line of them is retyped here:

  :48-59     load_calibration(calib_path, image_scale)                         -> calib_*   (P, P_inv)
  :113-135   boxes /= scale, P rescale, score filter, argsort, top-100 select  -> select_*
  :137-287   the 6-DoF pose loop (all branches as written, dead ones included)   -> locations, angles, dimensions
  :291       the `outputs` dict of the .mat file                                 -> out_*
  :295-330   the KITTI writer (Rodrigues -> 8 corners -> Y / h / r_y / alpha, the text line)  -> kitti_lines, kitti_rec
  utils/eval.py:96-118   the second caller's selection + reshape contract        -> eval_detections

Cases: `pose_main` (3 images worth of polled detections, all four orientation classes, padding rows below the score
threshold), `pose_identity` (detections whose axes are almost the camera axes: the small-angle branch of
cv2.Rodrigues, incl. an exactly axis-aligned box), `pose_flip` (rotations by ~pi, the other special case),
`pose_yaw` (pure yaw sweeps over (-pi, pi) and random rotations: r_y / alpha wrapping), `pose_main1` also has equal
scores (argsort order), `pose_empty` (no score above the threshold).
"""
import os
import sys
import tempfile
import textwrap
import types

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF_DRIVER = '/root/reference/keras_retinanet_3D/bin/run_network.py'
REF_EVAL = '/root/reference/keras_retinanet_3D/utils/eval.py'


def ref_lines(path, first, last):
    """Source lines first..last (1-based, inclusive) of a reference file, de-indented."""
    with open(path) as f:
        lines = f.readlines()
    return textwrap.dedent(''.join(lines[first - 1:last]))


def run_reference_driver_lines(outs, scale, P, image_hw):
    """run_network.py:113-330 for one image.  `outs` = the 8 arrays model.predict_on_batch returns."""
    boxes, dimensions, scores, labels, orientations, keypoints, keyplanes, residuals = [np.array(o, copy=True) for o in outs]
    ns = {'np': np, 'cv2': cv2, 'boxes': boxes, 'dimensions': dimensions, 'scores': scores, 'labels': labels,
          'orientations': orientations, 'keypoints': keypoints, 'keyplanes': keyplanes, 'residuals': residuals,
          'scale': scale, 'P': np.array(P, copy=True)}
    exec(compile(ref_lines(REF_DRIVER, 113, 135), REF_DRIVER + ':113-135', 'exec'), ns)
    select = {k: np.array(ns[k], copy=True) for k in ('boxes', 'dimensions', 'scores', 'labels', 'orientations',
                                                       'keypoints', 'keyplanes', 'residuals', 'P')}
    # np.empty_like rows of orientations outside 0..3 stay uninitialised in the reference; make them determinate
    src = ref_lines(REF_DRIVER, 137, 287).replace('np.empty_like', 'np.zeros_like')
    with np.errstate(all='ignore'):
        exec(compile(src, REF_DRIVER + ':137-287', 'exec'), ns)
    exec(compile(ref_lines(REF_DRIVER, 291, 291), REF_DRIVER + ':291', 'exec'), ns)
    outputs = {k: np.array(v, copy=True) for k, v in ns['outputs'].items()}
    # KITTI writer: the same lines, once to a real file ...
    tmp = tempfile.mkdtemp()
    os.makedirs(os.path.join(tmp, 'outputs', 'kitti'))
    ns.update({'os': os, 'args': types.SimpleNamespace(kitti=True), 'output_dir': tmp, 'image_fp': 'frame.png',
               'raw_image': np.zeros(tuple(image_hw) + (3,), np.uint8)})
    with np.errstate(all='ignore'):
        exec(compile(ref_lines(REF_DRIVER, 295, 330), REF_DRIVER + ':295-330', 'exec'), ns)
    with open(os.path.join(tmp, 'outputs', 'kitti', 'frame.txt')) as f:
        lines = f.readlines()
    # ... and once per row with the loop header replaced, to read (alpha, h, Y, r_y) before they are formatted
    body = ref_lines(REF_DRIVER, 298, 327)
    rec = np.zeros((len(ns['scores']), 4), np.float64)
    for i in range(len(ns['scores'])):
        ns['i'] = i
        with np.errstate(all='ignore'):
            exec(compile(body, REF_DRIVER + ':298-327', 'exec'), ns)
        rec[i] = (ns['alpha'], ns['h'], ns['Y'], ns['r_y'])
    return select, outputs, lines, rec


def run_reference_eval_lines(outs):
    """utils/eval.py:96-118: selection + reshape contract of the second caller (`_get_detections`)."""
    boxes, dimensions, scores, labels, orientations, plane_pts, planes, residuals = [np.array(o, copy=True) for o in outs]
    ns = {'np': np, 'boxes': boxes, 'dimensions': dimensions, 'scores': scores, 'labels': labels,
          'orientations': orientations, 'plane_pts': plane_pts, 'planes': planes, 'residuals': residuals,
          'score_threshold': 0.05, 'max_detections': 100}
    with open(REF_EVAL) as f:
        text = f.readlines()
    first = next(i for i, l in enumerate(text) if 'indices = np.where(scores[0, :] > score_threshold)' in l) + 1
    last = next(i for i, l in enumerate(text) if 'image_detections   = np.concatenate' in l) + 2
    exec(compile(textwrap.dedent(''.join(text[first - 2:last])), REF_EVAL, 'exec'), ns)
    return np.array(ns['image_detections'], copy=True), (first - 1, last)


def polled_case(seed, n_img, planes_tag, n_valid, rng):
    from gpp_b200.utils import synthetic
    from oracle import c_oracle
    planes = np.load(os.path.join(ROOT, 'road_planes_database', 'road_planes_database_%s.npy' % planes_tag))
    boxes, dims, orient, P_inv = synthetic.synth_detections(n_img, 100, planes, seed=seed, n_valid=n_valid)
    kp, kpl, res = c_oracle.fit_road_planes_c(boxes, dims, orient, P_inv, planes)
    return boxes, dims, orient, kp, kpl, res, P_inv


def outs_for_image(b, boxes, dims, orient, kp, kpl, res, scores):
    labels = np.where(scores > 0, 0, -1).astype(np.int32)
    return [boxes[b:b + 1].copy(), dims[b:b + 1].copy(), scores[None].astype(np.float32), labels[None],
            orient[b:b + 1].copy(), kp[b:b + 1].copy(), kpl[b:b + 1].copy(), res[b:b + 1].copy()]


def box_keypoints(R, t, h, w, l, o):
    """3-D key-points (X_l, X_m, X_r, X_t) of a box with rotation R (columns x, y, z), centre-bottom t, for class o,
    inverting the four reachable branches of the pose loop."""
    x, y, z = R[:, 0], R[:, 1], R[:, 2]
    if o == 1:      # loc = (X_m + X_r)/2 - z w/2, x = (X_m - X_r)/l
        mid = t + z * w / 2
        X_m, X_r = mid + x * l / 2, mid - x * l / 2
        X_l = X_m - z * w
    elif o == 2:    # loc = (X_m + X_r)/2 + z w/2, x = (X_r - X_m)/l
        mid = t - z * w / 2
        X_m, X_r = mid - x * l / 2, mid + x * l / 2
        X_l = X_m + z * w
    elif o == 0:    # loc = (X_m + X_l)/2 + z w/2, x = (X_m - X_l)/l
        mid = t - z * w / 2
        X_m, X_l = mid + x * l / 2, mid - x * l / 2
        X_r = X_m + z * w
    else:           # o == 3: loc = (X_m + X_l)/2 - z w/2, x = (X_l - X_m)/l
        mid = t + z * w / 2
        X_m, X_l = mid - x * l / 2, mid + x * l / 2
        X_r = X_m - z * w
    X_t = X_m - y * h
    return np.concatenate([X_l, X_m, X_r, X_t]).astype(np.float32)


def synthetic_rotation_case(rotvecs, rng):
    """100 rows whose pose is the given Rodrigues vector (cycled), padded to a (1, 100, ...) network output."""
    n = len(rotvecs)
    kp = np.zeros((1, 100, 4, 3), np.float32)
    dims = -np.ones((1, 100, 3), np.float32)
    orient = -np.ones((1, 100), np.int32)
    boxes = -np.ones((1, 100, 12), np.float32)
    for i, rv in enumerate(rotvecs):
        R = cv2.Rodrigues(np.asarray(rv, np.float64))[0]
        o = i % 4
        h, w, l = rng.uniform(1.3, 1.9), rng.uniform(1.5, 2.0), rng.uniform(3.5, 5.0)
        t = np.array([rng.uniform(-10, 10), rng.uniform(1.2, 2.0), rng.uniform(8, 40)])
        kp[0, i] = box_keypoints(R, t, h, w, l, o).reshape(4, 3)
        dims[0, i] = (h, w, l)
        orient[0, i] = o
        boxes[0, i] = rng.uniform(0, 1200, 12)
    kpl = np.tile(np.array([0.0, -1.0, 0.0, 1.65], np.float32), (1, 100, 1, 1))
    res = rng.uniform(0, 0.2, (1, 100)).astype(np.float32)
    scores = np.where(np.arange(100) < n, np.linspace(0.95, 0.3, 100), -1.0).astype(np.float32)
    return boxes, dims, orient, kp, kpl, res, scores


def main():
    rng = np.random.default_rng(20261017)
    scale = 1333.0 / 1242.0
    P2 = np.array([[721.5377, 0, 609.5593, 44.85728], [0, 721.5377, 172.854, 0.2163791], [0, 0, 1, 0.002745884]])
    P_scaled = np.dot(np.diag([scale, scale, 1.0]), P2)
    out = {}

    # ---- load_calibration (:48-59) on a KITTI-format calibration file
    ns = {'np': np}
    exec(compile(ref_lines(REF_DRIVER, 48, 59), REF_DRIVER + ':48-59', 'exec'), ns)
    tmp = tempfile.mkdtemp()
    calib = os.path.join(tmp, '000001.txt')
    with open(calib, 'w') as f:
        for k in range(4):
            M = P2 if k == 2 else P2 + k
            f.write('P%d: %s\n' % (k, ' '.join('%.12e' % v for v in M.reshape(-1))))
    P_ref, P_inv_ref = ns['load_calibration'](calib, scale)
    with open(calib) as f:
        calib_text = f.read()
    np.savez_compressed(os.path.join(HERE, 'driver_calib.npz'), calib_text=np.array(calib_text), scale=np.float64(scale),
                        P=P_ref, P_inv=P_inv_ref)

    cases = {}
    boxes, dims, orient, kp, kpl, res, P_inv = polled_case(301, 3, '1k', 83, rng)
    extra = {}
    for b in range(3):
        extra['pose_main%d' % b] = {'poll_P_inv': P_inv[b:b + 1], 'poll_planes_db': np.array('1k')}
        scores = np.where(np.arange(100) < 83, rng.uniform(0.0, 1.0, 100), -1.0).astype(np.float32)
        if b == 1:
            scores[5:9] = scores[5]                        # equal scores: argsort order of the reference
        cases['pose_main%d' % b] = outs_for_image(b, boxes, dims, orient, kp, kpl, res, scores)
    small = [rng.normal(0, s, 3) for s in (1e-9, 1e-7, 1e-5, 1e-4, 1e-3, 1e-2) for _ in range(4)] + [np.zeros(3)] * 4
    c = synthetic_rotation_case(small, rng)
    cases['pose_identity'] = outs_for_image(0, c[0], c[1], c[2], c[3], c[4], c[5], c[6])
    flips = []
    for eps in (0.0, 1e-7, 1e-5, 1e-3, 1e-2):
        for axis in ((1, 0, 0), (0, 1, 0), (0, 0, 1), (0.6, 0.8, 0), (0.3, -0.5, 0.81)):
            a = np.asarray(axis, np.float64)
            flips.append(a / np.linalg.norm(a) * (np.pi - eps))
    c = synthetic_rotation_case(flips, rng)
    cases['pose_flip'] = outs_for_image(0, c[0], c[1], c[2], c[3], c[4], c[5], c[6])
    yaws = [np.array([0.0, y, 0.0]) for y in np.linspace(-3.1, 3.1, 40)] + [rng.normal(0, 1.2, 3) for _ in range(40)]
    c = synthetic_rotation_case(yaws, rng)
    cases['pose_yaw'] = outs_for_image(0, c[0], c[1], c[2], c[3], c[4], c[5], c[6])
    e = [a.copy() for a in cases['pose_main0']]
    e[2][...] = 0.01                                        # nothing above the score threshold
    cases['pose_empty'] = e

    for name, outs in cases.items():
        select, outputs, lines, rec = run_reference_driver_lines(outs, scale, P_scaled, (375, 1242))
        det, eval_lines = run_reference_eval_lines(outs)
        rec_d = {'in_%s' % k: v for k, v in zip(('boxes', 'dimensions', 'scores', 'labels', 'orientations', 'keypoints',
                                                  'keyplanes', 'residuals'), outs)}
        rec_d.update({'select_%s' % k: v for k, v in select.items()})
        rec_d.update({'out_%s' % k: v for k, v in outputs.items()})
        rec_d['kitti_lines'] = np.array(''.join(lines))
        rec_d['kitti_rec'] = rec
        rec_d['eval_detections'] = det
        rec_d['scale'] = np.float64(scale)
        rec_d['P_scaled'] = P_scaled
        rec_d.update(extra.get(name, {}))
        np.savez_compressed(os.path.join(HERE, name + '.npz'), **rec_d)
        print('%-14s kept %3d rows, %d KITTI lines, eval.py lines %d-%d' % (name, len(outputs['scores']), len(lines),
                                                                          eval_lines[0], eval_lines[1]))


if __name__ == '__main__':
    main()
