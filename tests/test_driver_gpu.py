"""GPU test of the run_network driver: files written for a directory of synthetic KITTI-sized
images equal what the oracle chain (oracle/driver_ref.py) produces from the same head tensors."""
import os

import numpy as np
import pytest

import gpp_b200
from gpp_b200.utils import anchors as A
from gpp_b200.utils import calibration, synthetic

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _write_scene(tmp_path, n_images=2):
    import cv2
    img_dir, cal_dir = tmp_path / 'image_2', tmp_path / 'calib'
    img_dir.mkdir(), cal_dir.mkdir()
    yy, xx = np.mgrid[0:375, 0:1242]
    for k in range(n_images):
        img = np.stack([(128 + 100 * np.sin(xx / (30.0 + 7 * k) + c) * np.cos(yy / (20.0 + 3 * c))) for c in range(3)],
                       axis=-1).astype(np.uint8)
        cv2.imwrite(str(img_dir / ('%06d.png' % k)), img)
        P2 = synthetic.KITTI_P2.reshape(-1)
        with open(cal_dir / ('%06d.txt' % k), 'w') as f:
            for cam in range(4):
                row = P2 if cam == 2 else np.arange(12.0)
                f.write('P%d: %s\n' % (cam, ' '.join('%.12e' % v for v in row)))
    return str(img_dir), str(cal_dir)


def test_driver_writes_the_files_the_reference_driver_would(tmp_path):
    import scipy.io
    from gpp_b200.bin import run_network as rn
    from oracle import driver_ref
    img_dir, cal_dir = _write_scene(tmp_path)
    planes_fp = os.path.join(ROOT, 'road_planes_database', 'road_planes_database_1k.npy')
    heads_dir = str(tmp_path / 'heads')
    out_root = str(tmp_path / 'out')
    os.mkdir(out_root)
    out_dir = rn.main(['standin:1', img_dir, cal_dir, planes_fp, out_root, '--kitti', '--dump-heads', heads_dir])
    assert out_dir == os.path.join(out_root, 'standin_1')
    planes = np.load(planes_fp)
    anchors = A.anchors_for_shape((402, 1333))
    total = 0
    for k in range(2):
        stem = '%06d' % k
        with np.load(os.path.join(heads_dir, stem + '.npz')) as z:
            heads = [z[n][None] for n in ('regression', 'regression_dim', 'classification')]
        assert heads[0].shape == (1, 137256, 12)
        scale = 1333 / 1242
        P, P_inv = calibration.load_calibration(os.path.join(cal_dir, stem + '.txt'), scale)
        outs = driver_ref.predict_on_batch_ref(anchors, heads[0], heads[1], heads[2], P_inv[None], planes)
        want, want_lines = driver_ref.driver_image_ref(outs, scale, (1242, 375))
        got = scipy.io.loadmat(os.path.join(out_dir, 'outputs', 'full', stem + '.mat'))
        n = want['scores'].shape[0]
        total += n
        assert np.array_equal(got['scores'].reshape(-1), want['scores'])
        assert np.array_equal(got['labels'].reshape(-1), want['labels'])
        assert np.array_equal(got['boxes'].reshape(n, 4), want['boxes'])
        assert np.array_equal(got['keypoints'].reshape(n, 8), want['keypoints'])
        assert np.array_equal(got['residuals'].reshape(-1), want['residuals'])
        for key in ('locations', 'angles', 'dimensions'):
            g, w = got[key].reshape(n, 3), want[key]
            ok = np.isfinite(w)
            assert np.array_equal(np.isfinite(g), ok)
            assert np.allclose(g[ok], w[ok], rtol=1e-4, atol=1e-4), key      # north-star tolerance
        with open(os.path.join(out_dir, 'outputs', 'kitti', stem + '.txt')) as f:
            lines = f.readlines()
        assert len(lines) == n
        for a, b in zip(lines, want_lines):
            fa, fb = a.split(), b.split()
            assert fa[:3] == fb[:3] == ['Car', '-1', '-1']
            va, vb = np.array(fa[3:], dtype=np.float64), np.array(fb[3:], dtype=np.float64)
            ok = np.isfinite(vb)
            assert np.array_equal(np.isfinite(va), ok)
            # two-decimal text of values that agree to 1e-4: at most one unit in the last printed place
            assert np.all(np.abs(va[ok] - vb[ok]) <= 0.0100001 + 1e-4 * np.abs(vb[ok]))
    assert total > 0                                              # the stand-in produced detections at all


def test_driver_runs_from_dumped_heads(tmp_path):
    from gpp_b200.bin import run_network as rn
    img_dir, cal_dir = _write_scene(tmp_path, n_images=1)
    planes_fp = os.path.join(ROOT, 'road_planes_database', 'road_planes_database_100.npy')
    heads_dir = str(tmp_path / 'heads')
    out_root = str(tmp_path / 'out')
    os.mkdir(out_root)
    a = rn.main(['standin:2', img_dir, cal_dir, planes_fp, out_root, '--kitti', '--dump-heads', heads_dir])
    b = rn.main(['heads:' + heads_dir, img_dir, cal_dir, planes_fp, out_root, '--kitti', '--mode', 'exact'])
    with open(os.path.join(a, 'outputs', 'kitti', '000000.txt')) as f, open(os.path.join(b, 'outputs', 'kitti', '000000.txt')) as g:
        assert f.read() == g.read()                               # verified == exact, heads round-trip through .npz
    with pytest.raises(SystemExit):
        rn.main(['standin', img_dir, cal_dir, planes_fp, out_root, '--save-images'])
