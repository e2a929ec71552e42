"""CPU tests of the driver-side utilities against vectors produced by the reference's own
utils/anchors.py and utils/image.py (tests/golden/make_golden_driver.py)."""
import hashlib
import os

import numpy as np
import pytest

import gpp_b200
from gpp_b200.utils import anchors as A
from gpp_b200.utils import image as I

GOLD = os.path.join(os.path.dirname(__file__), 'golden', 'driver_utils.npz')


@pytest.fixture(scope='module')
def gold():
    with np.load(GOLD) as z:
        return {k: z[k] for k in z.files}


def test_anchors_equal_the_reference_bit_for_bit(gold):
    small = A.anchors_for_shape((96, 160, 3))
    assert small.dtype == np.float64 and np.array_equal(small, gold['anchors_small'])
    big = A.anchors_for_shape((402, 1333, 3))
    assert big.shape == (int(gold['anchors_kitti_rows']), 4) == (137256, 4)
    digest = np.frombuffer(hashlib.sha256(np.ascontiguousarray(big).tobytes()).digest(), dtype=np.uint8)
    assert np.array_equal(digest, gold['anchors_kitti_sha256'])
    assert np.array_equal(big[::997], gold['anchors_kitti_sample'])


def test_anchor_layout_is_cell_major_with_12_anchors_per_cell():
    a = A.anchors_for_shape((402, 1333))
    cell0 = a[:12]
    centres = np.stack([(cell0[:, 0] + cell0[:, 2]) / 2, (cell0[:, 1] + cell0[:, 3]) / 2], axis=1)
    assert np.allclose(centres, 4.0)                            # first P3 cell: (0 + 0.5) * stride 8
    assert np.allclose((a[12:24, 0] + a[12:24, 2]) / 2, 12.0)   # next cell along x
    ratios = (cell0[:, 3] - cell0[:, 1]) / (cell0[:, 2] - cell0[:, 0])
    assert np.allclose(ratios, np.repeat([0.5, 1.0, 2.0], 4))
    cached = A.cached_anchors(402, 1333)
    assert cached.dtype == np.float32 and not cached.flags.writeable and cached is A.cached_anchors(402, 1333)


def test_generate_anchors_custom_ratios_and_scales():
    g = A.generate_anchors(32, ratios=[1.0], scales=[1.0, 2.0])
    assert np.allclose(g, [[-16, -16, 16, 16], [-32, -32, 32, 32]])
    shapes = A.guess_shapes((402, 1333, 3), [3, 4, 5, 6, 7])
    assert [tuple(s) for s in shapes] == [(51, 167), (26, 84), (13, 42), (7, 21), (4, 11)]


def test_image_preprocessing_equals_the_reference(gold):
    pre = I.preprocess_image(gold['image_raw'])
    assert pre.dtype == np.float32 and np.array_equal(pre, gold['image_preprocessed'])
    assert gold['image_raw'].dtype == np.uint8                   # the input is not modified
    resized, scale = I.resize_image(pre)
    assert scale == float(gold['image_scale']) and np.array_equal(resized.shape, gold['image_resized_shape'])
    digest = np.frombuffer(hashlib.sha256(np.ascontiguousarray(resized).tobytes()).digest(), dtype=np.uint8)
    assert np.array_equal(digest, gold['image_resized_sha256'])
    assert np.array_equal(resized.reshape(-1)[::4099], gold['image_resized_sample'])
    small, s2 = I.resize_image(pre, min_side=24, max_side=64)
    assert s2 == float(gold['image_small_scale']) and np.array_equal(small, gold['image_small'])
    _, ks = I.resize_image(np.zeros((375, 1242, 3), np.float32))
    assert ks == float(gold['kitti_scale']) == 1333 / 1242


def test_read_image_bgr_round_trip(tmp_path):
    import cv2
    img = np.zeros((8, 12, 3), np.uint8)
    img[..., 0], img[..., 2] = 10, 200                          # B = 10, R = 200
    fp = str(tmp_path / 'a.png')
    cv2.imwrite(fp, img)
    back = I.read_image_bgr(fp)
    assert back.dtype == np.uint8 and np.array_equal(back, img)
    with pytest.raises(ValueError):
        I.read_image_bgr(str(tmp_path / 'missing.png'))


def test_standin_detector_shapes_and_determinism():
    import torch
    from gpp_b200.utils.standin_detector import StandInDetector
    x = torch.from_numpy(np.random.default_rng(0).normal(0, 50, size=(1, 96, 160, 3)).astype(np.float32))
    r, d, c = StandInDetector(3)(x)
    n = A.anchors_for_shape((96, 160)).shape[0]
    assert r.shape == (1, n, 12) and d.shape == (1, n, 3) and c.shape == (1, n, 8)
    assert float(c.min()) >= 0.0 and float(c.max()) <= 1.0
    r2, _, c2 = StandInDetector(3)(x)
    assert torch.equal(r, r2) and torch.equal(c, c2)
    assert not torch.equal(StandInDetector(4)(x)[0], r)


def test_driver_argument_parsing_and_model_names():
    from gpp_b200.bin import run_network as rn
    a = rn.parse_args(['standin:3', 'img', 'cal', 'planes.mat', 'out', '--kitti'])
    assert (a.model_path, a.image_dir, a.calib_dir, a.plane_params_path, a.output_dir) == ('standin:3', 'img', 'cal', 'planes.mat', 'out')
    assert a.kitti and not a.save_images and a.backbone == 'resnet50'
    assert rn.model_name('/x/resnet50_kitti_01.h5') == 'resnet50_kitti_01'      # run_network.py:78
    assert rn.model_name('standin:3') == 'standin_3'
    with pytest.raises(ValueError):
        rn.load_detector('/x/model.h5', None)
