"""
ORACLE (test infrastructure, NOT product code) -- restatement of the KITTI-record arithmetic of
/root/reference/keras_retinanet_3D/bin/run_network.py:297-323 (numpy + cv2.Rodrigues like the reference).
Parity status: PINNED -- equal, bit for bit, to (alpha, h, Y, r_y) as the reference's own writer lines compute them
(tests/golden/pose_*.npz `kitti_rec`, made by exec'ing :298-327; tests/test_pose_oracle_golden.py).  Tolerance against
the CUDA path 1e-4 (north_star).
"""
import numpy as np

try:
    import cv2
except Exception:  # pragma: no cover
    cv2 = None


def kitti_records_ref(locations, angles, dimensions):
    n = len(locations)
    out = np.zeros((n, 4), np.float64)
    for i in range(n):
        h = dimensions[i, 0]
        w = dimensions[i, 1]
        l = dimensions[i, 2]
        x_corners = np.array([l / 2, l / 2, -l / 2, -l / 2, l / 2, l / 2, -l / 2, -l / 2])
        y_corners = np.array([0, 0, 0, 0, -h, -h, -h, -h])
        z_corners = np.array([w / 2, -w / 2, -w / 2, w / 2, w / 2, -w / 2, -w / 2, w / 2])
        R = cv2.Rodrigues(angles[i, :])[0]
        X_all = np.matmul(R, np.stack([x_corners, y_corners, z_corners], axis=0))
        X_all[0, :] = X_all[0, :] + locations[i, 0]
        X_all[1, :] = X_all[1, :] + locations[i, 1]
        X_all[2, :] = X_all[2, :] + locations[i, 2]
        r_y = angles[i, 1] % (2 * np.pi)
        if r_y < -np.pi:
            r_y = r_y + 2 * np.pi
        elif r_y >= np.pi:
            r_y = r_y - 2 * np.pi
        Y = np.amax(X_all[1, :])
        h = Y - np.amin(X_all[1, :])
        alpha = r_y + np.arctan2(locations[i, 2], locations[i, 0]) + 1.5 * np.pi
        alpha = alpha % (2 * np.pi)
        if alpha < -np.pi:
            alpha = alpha + 2 * np.pi
        elif alpha >= np.pi:
            alpha = alpha - 2 * np.pi
        out[i] = (alpha, h, Y, r_y)
    return out
