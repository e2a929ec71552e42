#!/usr/bin/env python
"""
Driver with the command line and the output files of the reference's
/root/reference/keras_retinanet_3D/bin/run_network.py (``retinanet-3D-run-network``), with everything after the CNN
on the GPU through libgpp: box / dimension decode, FilterDetections, ground-plane polling, pose recovery and the
KITTI record arithmetic.  The CNN is out of scope (Keras / TensorFlow weights), so ``model_path`` names a stand-in:

    standin[:SEED]     random-init torch detector with the reference's head layout (utils/standin_detector.py)
    heads:DIR          head tensors dumped by any detector: DIR/<image stem>.npz with ``regression`` (A, 12),
                       ``regression_dim`` (A, 3), ``classification`` (A, 8) for the RESIZED image (402 x 1333 for KITTI)

Per image it writes ``<output_dir>/<model name>/outputs/full/<stem>.mat`` (run_network.py:291-292) and, with
``--kitti``, ``outputs/kitti/<stem>.txt`` (:295-326).  ``--save-images`` (visualisation) is not provided.
"""
import argparse
import os
import shutil
import sys
import time

import numpy as np

if __name__ == '__main__' and __package__ is None:            # allow running the file directly
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', '..'))
    import gpp_b200  # noqa: F401
    __package__ = 'gpp_b200.bin'

from ..pipeline import detections_from_heads
from ..utils.anchors import cached_anchors
from ..utils.calibration import load_calibration, load_road_planes
from ..utils.image import preprocess_image, read_image_bgr, resize_image
from ..utils.kitti import format_kitti_lines, postprocess_image


def parse_args(args):
    """The reference's arguments (run_network.py:27-41) plus --mode / --dump-heads / --device."""
    parser = argparse.ArgumentParser(description='Run detector heads + GPU ground-plane polling on a directory of images.')
    parser.add_argument('model_path', help="Detector: 'standin[:SEED]' or 'heads:DIR' (see module docstring).", type=str)
    parser.add_argument('image_dir', help='Path to directory of input images.', type=str)
    parser.add_argument('calib_dir', help='Path to directory of calibration files.', type=str)
    parser.add_argument('plane_params_path', help='Path to .MAT (or .npy) file containing road planes.', type=str)
    parser.add_argument('output_dir', help='Path to output directory', type=str)
    parser.add_argument('--kitti', help='Include to save results in KITTI format.', action='store_true')
    parser.add_argument('--save-images', help='Not provided (visualisation is out of scope).', action='store_true')
    parser.add_argument('--backbone', help='Accepted for compatibility; ignored.', default='resnet50')
    parser.add_argument('--mode', help='Polling arithmetic: verified (default), exact, fast, f64.', default=None)
    parser.add_argument('--dump-heads', help='Also write the head tensors of every image to this directory.', default=None)
    parser.add_argument('--device', help='CUDA device index.', type=int, default=0)
    return parser.parse_args(args)


def model_name(model_path):
    """Name of the per-model output directory: the reference strips the '.h5' of the file name (:78)."""
    base = os.path.basename(model_path.rstrip('/')) or model_path
    base = base.replace(':', '_')
    return os.path.splitext(base)[0] if os.path.splitext(base)[1] in ('.h5', '.pt', '.npz') else base


def load_detector(model_path, device):
    """Returns f(image (rows, cols, 3) float32, stem) -> (regression, regression_dim, classification) CUDA tensors
    with a leading batch axis of 1."""
    import torch
    if model_path.startswith('heads:'):
        root = model_path[len('heads:'):]

        def from_files(image, stem):
            with np.load(os.path.join(root, stem + '.npz')) as z:
                return tuple(torch.from_numpy(np.ascontiguousarray(z[k], dtype=np.float32)).to(device)[None]
                             for k in ('regression', 'regression_dim', 'classification'))
        return from_files
    if model_path == 'standin' or model_path.startswith('standin:'):
        from ..utils.standin_detector import StandInDetector
        seed = int(model_path.split(':', 1)[1]) if ':' in model_path else 0
        net = StandInDetector(seed).to(device).eval()
        return lambda image, stem: net(torch.from_numpy(image).to(device)[None])
    raise ValueError("model_path must be 'standin[:SEED]' or 'heads:DIR' (Keras .h5 models need the reference's "
                     "TensorFlow stack, which this library replaces only after the CNN)")


_DEVICE_ANCHORS = {}


def device_anchors(rows, cols, device):
    """The (A, 4) float32 anchors of a resized image as a CUDA tensor, uploaded once per shape."""
    import torch
    key = (rows, cols, str(device))
    if key not in _DEVICE_ANCHORS:
        _DEVICE_ANCHORS[key] = torch.tensor(cached_anchors(rows, cols), device=device)
    return _DEVICE_ANCHORS[key]


def process_image(detector, image_fp, calib_fp, planes_dev, device, mode=None, dump_heads=None):
    """One iteration of the reference's loop (run_network.py:91-287): returns (outputs dict, raw image (W, H))."""
    import torch
    stem = os.path.splitext(os.path.basename(image_fp))[0]
    raw_image = read_image_bgr(image_fp)
    image, scale = resize_image(preprocess_image(raw_image))
    P, P_inv = load_calibration(calib_fp, scale)
    regression, regression_dim, classification = detector(image, stem)
    anchors = device_anchors(image.shape[0], image.shape[1], device)
    if regression.shape[1] != anchors.shape[0]:
        raise ValueError('%s: detector produced %d anchors, the image shape %r has %d' % (
            stem, regression.shape[1], image.shape[:2], anchors.shape[0]))
    if dump_heads:
        np.savez(os.path.join(dump_heads, stem + '.npz'), regression=regression[0].cpu().numpy(),
                 regression_dim=regression_dim[0].cpu().numpy(), classification=classification[0].cpu().numpy())
    p_inv = torch.from_numpy(P_inv.astype(np.float32)).to(device)[None]
    outs = detections_from_heads(anchors, regression, regression_dim, classification, p_inv, planes_dev, mode=mode)
    boxes, dimensions, scores, labels, orientations, keypoints, keyplanes, residuals = [t[0].cpu().numpy() for t in outs]
    post = postprocess_image(boxes, dimensions, scores, labels, orientations, keypoints, keyplanes, residuals, scale,
                             device=device.index)
    return post, (raw_image.shape[1], raw_image.shape[0])


def main(args=None):
    args = parse_args(sys.argv[1:] if args is None else args)
    if args.save_images:
        raise SystemExit('--save-images: visualisation is not part of this library (see DESIGN.md, out of scope)')
    import scipy.io
    import torch
    device = torch.device('cuda', args.device)
    detector = load_detector(args.model_path, device)
    plane_params = load_road_planes(args.plane_params_path)
    planes_dev = torch.from_numpy(np.ascontiguousarray(plane_params, dtype=np.float32)).to(device)

    output_dir = os.path.join(args.output_dir, model_name(args.model_path))
    if os.path.isdir(output_dir):
        shutil.rmtree(output_dir)
    os.makedirs(os.path.join(output_dir, 'outputs', 'full'))
    if args.kitti:
        os.mkdir(os.path.join(output_dir, 'outputs', 'kitti'))
    if args.dump_heads:
        os.makedirs(args.dump_heads, exist_ok=True)

    for j, fn in enumerate(sorted(os.listdir(args.calib_dir))):
        calib_fp = os.path.join(args.calib_dir, fn)
        image_fp = os.path.join(args.image_dir, fn.replace('.txt', '.png'))
        start = time.time()
        post, image_wh = process_image(detector, image_fp, calib_fp, planes_dev, device, args.mode, args.dump_heads)
        print('Image {}: frame rate: {:.2f}'.format(j, 1.0 / (time.time() - start)))
        stem = os.path.basename(image_fp)[:-3]
        outputs = {k: post[k] for k in ('boxes', 'keypoints', 'labels', 'scores', 'locations', 'angles', 'dimensions',
                                        'residuals')}                                  # run_network.py:291
        scipy.io.savemat(os.path.join(output_dir, 'outputs', 'full', stem + 'mat'), outputs)
        if args.kitti:
            full_boxes = np.concatenate([post['boxes'], post['keypoints']], axis=1)
            with open(os.path.join(output_dir, 'outputs', 'kitti', stem + 'txt'), 'w') as f:
                f.writelines(format_kitti_lines(full_boxes, post['dimensions'], post['locations'], post['scores'],
                                                post['kitti'], image_wh))
    return output_dir


if __name__ == '__main__':
    main()
