"""
ORACLE (test infrastructure, NOT product code) -- restatement of the 6-DoF pose loop of
/root/reference/keras_retinanet_3D/bin/run_network.py:137-247 (numpy float32 + cv2.Rodrigues exactly like
the reference).  Only the branches that loop can reach are restated: ``outlier`` is 2 for orientation 0/3
and 0 for orientation 1/2 (:147-150), so orientation 1 -> :167-177, 2 -> :178-188, 0 -> :204-214,
3 -> :237-247; the other branches (:156-166, :189-199, :215-236, :248-287) are dead code.
Parity status: PINNED -- tests/golden/pose_*.npz hold what the reference's own lines (:137-287, cut out of the file and
exec'd by tests/golden/make_golden_pose.py) produce, and this restatement equals them bit for bit
(tests/test_pose_oracle_golden.py).  Tolerance against the CUDA path: 1e-4 relative (BASELINE.json north_star).
"""
import numpy as np

try:
    import cv2
except Exception:  # pragma: no cover
    cv2 = None


def pose_ref(keypoints, dimensions, orientations, locations=None, angles=None):
    """keypoints (n, 12) float32, dimensions (n, 3) float32 (MODIFIED IN PLACE like the reference),
    orientations (n,) int.  Returns (locations (n, 3), angles (n, 3), dimensions (n, 3))."""
    keypoints = np.asarray(keypoints, dtype=np.float32)
    n = keypoints.shape[0]
    if angles is None:
        angles = np.zeros_like(dimensions)       # reference: np.empty_like (:137); zeros keep tests determinate
    if locations is None:
        locations = np.zeros_like(dimensions)    # reference: np.empty_like (:138)
    for i in range(n):
        X_l = keypoints[i, 0:3]
        X_m = keypoints[i, 3:6]
        X_r = keypoints[i, 6:9]
        X_t = keypoints[i, 9:12]
        o = orientations[i]
        if o == 1:                                                    # :167-177
            dimensions[i, 0] = np.linalg.norm(X_t - X_m)
            dimensions[i, 2] = np.linalg.norm(X_r - X_m)
            x_dir = (X_m - X_r) / dimensions[i, 2]
            y_dir = (X_m - X_t) / dimensions[i, 0]
            z_dir = np.cross(x_dir, y_dir)
            locations[i, :] = (X_m + X_r) / 2 - z_dir * dimensions[i, 1] / 2
        elif o == 2:                                                  # :178-188
            dimensions[i, 0] = np.linalg.norm(X_t - X_m)
            dimensions[i, 2] = np.linalg.norm(X_r - X_m)
            x_dir = (X_r - X_m) / dimensions[i, 2]
            y_dir = (X_m - X_t) / dimensions[i, 0]
            z_dir = np.cross(x_dir, y_dir)
            locations[i, :] = (X_m + X_r) / 2 + z_dir * dimensions[i, 1] / 2
        elif o == 0:                                                  # :204-214
            dimensions[i, 0] = np.linalg.norm(X_t - X_m)
            dimensions[i, 2] = np.linalg.norm(X_l - X_m)
            x_dir = (X_m - X_l) / dimensions[i, 2]
            y_dir = (X_m - X_t) / dimensions[i, 0]
            z_dir = np.cross(x_dir, y_dir)
            locations[i, :] = (X_m + X_l) / 2 + z_dir * dimensions[i, 1] / 2
        elif o == 3:                                                  # :237-247
            dimensions[i, 0] = np.linalg.norm(X_t - X_m)
            dimensions[i, 2] = np.linalg.norm(X_l - X_m)
            x_dir = (X_l - X_m) / dimensions[i, 2]
            y_dir = (X_m - X_t) / dimensions[i, 0]
            z_dir = np.cross(x_dir, y_dir)
            locations[i, :] = (X_m + X_l) / 2 - z_dir * dimensions[i, 1] / 2
        else:
            continue                                                  # no branch runs for padding rows
        angles[i, :] = cv2.Rodrigues(np.stack([x_dir, y_dir, z_dir], axis=-1))[0][:, 0]
    return locations, angles, dimensions


def kitti_yaw(angles):
    """r_y of the KITTI writer (run_network.py:312-316): angles[:, 1] wrapped to [-pi, pi)."""
    r_y = np.asarray(angles)[:, 1] % (2 * np.pi)
    r_y = np.where(r_y >= np.pi, r_y - 2 * np.pi, r_y)
    return r_y
