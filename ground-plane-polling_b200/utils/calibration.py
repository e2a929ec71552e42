"""Calibration and plane-database loading, as the reference's driver does it
(/root/reference/keras_retinanet_3D/bin/run_network.py:48-59 and :75)."""
import numpy as np

__all__ = ['load_calibration', 'scale_projection', 'load_road_planes']


def scale_projection(P, image_scale):
    """P <- diag(s, s, 1) . P and its pseudo-inverse (run_network.py:56-58); float64 like the reference."""
    P = np.asarray(P, dtype=np.float64).reshape(3, 4)
    P = np.dot(np.array([[image_scale, 0.0, 0.0], [0.0, image_scale, 0.0], [0.0, 0.0, 1.0]]), P)
    return P, np.linalg.pinv(P)


def load_calibration(calib_path, image_scale):
    """ Load inverse of camera projection matrix from a KITTI calibration file (camera 2 = third line). """
    cam_id = 2
    with open(calib_path, 'r') as f:
        line = f.readlines()[cam_id]
    key, value = line.split(':', 1)
    return scale_projection([float(x) for x in value.split()], image_scale)


def load_road_planes(path):
    """The (N, 4) float64 road-plane database: the reference's MATLAB files
    (``scipy.io.loadmat(path)['road_planes_database']``, run_network.py:75) or the .npy copies of this repo."""
    if str(path).endswith('.npy'):
        db = np.load(path)
    else:
        import scipy.io
        db = scipy.io.loadmat(path)['road_planes_database']
    db = np.asarray(db)
    if db.ndim != 2 or db.shape[1] != 4:
        raise ValueError('road plane database must have shape (N, 4), got %r' % (db.shape,))
    return db
