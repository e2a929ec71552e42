"""CPU: host-side glue that mirrors the reference driver (selection, KITTI text lines, calibration) and the
oracles of the post-processing steps."""
import numpy as np

from gpp_b200.utils import kitti, synthetic
from oracle.kitti_ref import kitti_records_ref
from oracle.pose_ref import kitti_yaw, pose_ref


def test_select_detections_matches_the_driver():
    rng = np.random.default_rng(0)
    scores = np.concatenate([rng.uniform(0, 1, 60), -np.ones(40)]).astype(np.float32)
    keep = kitti.select_detections(scores, 0.05, 100)
    # run_network.py:117-125
    indices = np.where(scores > 0.05)[0]
    order = np.argsort(-scores[indices])[:100]
    assert np.array_equal(keep, indices[order])
    assert (np.diff(scores[keep]) <= 0).all() and (scores[keep] > 0.05).all()
    assert len(kitti.select_detections(scores, 0.05, 10)) == 10


def test_kitti_line_format():
    boxes = np.array([[-5.0, 10.0, 1300.0, 400.0]], np.float32)
    dims = np.array([[1.5, 1.6, 4.0]], np.float32)
    loc = np.array([[1.0, 1.7, 20.0]], np.float32)
    rec = np.array([[-1.57, 1.5, 1.7, 0.1]], np.float32)
    line = kitti.format_kitti_lines(boxes, dims, loc, np.array([0.9]), rec, (1242, 375))[0]
    f = line.split()
    assert f[0] == 'Car' and f[1] == '-1' and f[2] == '-1' and len(f) == 16
    assert f[4:8] == ['0.00', '10.00', '1242.00', '375.00']          # box clipped to the image (:325-326)
    assert f[8:11] == ['1.50', '1.60', '4.00'] and f[11:14] == ['1.00', '1.70', '20.00']


def test_pose_oracle_recovers_the_synthetic_truth():
    """End-to-end sanity of the oracles: polling + pose on zero-noise synthetic cars gives back location,
    yaw and dimensions (validates the orientation tables and the sign conventions of all four branches)."""
    from oracle import c_oracle
    planes = np.load(__import__('os').path.join(__import__('conftest').ROOT, 'road_planes_database',
                                                'road_planes_database_1k.npy'))
    boxes, dims, orient, P_inv, truth = synthetic.synth_detections(2, 100, planes, seed=7, kp_noise_px=0.0,
                                                                   dim_noise=0.0, return_truth=True)
    kp, kpl, res = c_oracle.fit_road_planes_c(boxes, dims, orient, P_inv, planes)
    loc, ang, dout = pose_ref(kp.reshape(-1, 12), dims.reshape(-1, 3).copy(), orient.reshape(-1))
    t = truth['location'].reshape(-1, 3)
    assert np.median(np.linalg.norm(loc - t, axis=1)) < 0.5
    dyaw = (kitti_yaw(ang) - truth['ry'].reshape(-1) + np.pi) % (2 * np.pi) - np.pi
    assert np.median(np.abs(dyaw)) < 0.05
    assert set(np.unique(orient)) == {0, 1, 2, 3}
    rec = kitti_records_ref(loc, ang, dout)
    assert np.median(np.abs(rec[:, 1] - dims.reshape(-1, 3)[:, 0])) < 0.2     # box height ~ h
    assert np.all(rec[:, 3] >= -np.pi) and np.all(rec[:, 3] < np.pi)


def test_synthetic_generator_layout():
    planes = np.load(__import__('os').path.join(__import__('conftest').ROOT, 'road_planes_database',
                                                'road_planes_database_100.npy'))
    b, d, o, p = synthetic.synth_detections(3, 100, planes, seed=1, n_valid=17)
    assert b.shape == (3, 100, 12) and b.dtype == np.float32
    assert d.shape == (3, 100, 3) and o.shape == (3, 100) and o.dtype == np.int32
    assert p.shape == (3, 4, 3) and p.dtype == np.float64          # callers feed float64 P_inv
    assert (b[:, 17:] == -1).all() and (d[:, 17:] == -1).all() and (o[:, 17:] == -1).all()
    b2 = synthetic.synth_detections(3, 100, planes, seed=1, n_valid=17)[0]
    assert np.array_equal(b, b2)                                    # seeded
    P, P_inv = synthetic.kitti_calibration()
    assert np.allclose(P @ P_inv, np.eye(3), atol=1e-9)


def test_calibration_and_plane_loading(tmp_path):
    import gpp_b200
    calib = tmp_path / '000001.txt'
    P2 = synthetic.KITTI_P2.reshape(-1)
    lines = ['P%d: %s\n' % (i, ' '.join('%.12e' % (v + (i - 2)) for v in P2)) for i in range(4)]
    calib.write_text(''.join(lines))
    P, P_inv = gpp_b200.load_calibration(str(calib), synthetic.KITTI_SCALE)
    P_ref, P_inv_ref = synthetic.kitti_calibration()
    assert np.allclose(P, P_ref, rtol=1e-10) and np.allclose(P_inv, P_inv_ref, rtol=1e-8)
    db = gpp_b200.load_road_planes(__import__('os').path.join(__import__('conftest').ROOT, 'road_planes_database',
                                                              'road_planes_database_10.npy'))
    assert db.shape == (10, 4) and db.dtype == np.float64
    import scipy.io
    mat = tmp_path / 'db.mat'
    scipy.io.savemat(str(mat), {'road_planes_database': db})
    assert np.array_equal(gpp_b200.load_road_planes(str(mat)), db)
