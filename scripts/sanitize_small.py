"""Small run of every kernel path for compute-sanitizer (memcheck / racecheck / synccheck):
    compute-sanitizer --tool memcheck python scripts/sanitize_small.py"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gpp_b200
from gpp_b200.utils import synthetic
from oracle import c_oracle

planes = np.load(os.path.join(ROOT, 'road_planes_database', 'road_planes_database_10k.npy'))[:2500]
poller = gpp_b200.get_poller(0)
for (B, D, nv, force) in ((2, 20, 15, 0), (40, 50, None, 200), (3, 9, 7, 100)):
    boxes, dims, orient, P_inv = synthetic.synth_detections(B, D, planes, seed=B, n_valid=nv)
    dims = dims.copy(); dims[0, :3, 1] *= 1.7
    want = c_oracle.fit_road_planes_c(boxes, dims, orient, P_inv, planes, return_index=True)
    poller.debug_set_config(force, 0)
    for mode in ('verified', 'exact', 'fast', 'f64'):
        got = gpp_b200.fit_road_planes(boxes, dims, orient, P_inv, planes, mode=mode, return_index=True)
        if mode in ('verified', 'exact'):
            assert all(np.array_equal(g, w, equal_nan=True) for g, w in zip(got, want)), (mode, B, D)
    poller.debug_set_config(0, 0)
kp = got[0].astype(np.float32).reshape(-1, 12)
loc, ang, dd = gpp_b200.recover_pose(kp, dims.reshape(-1, 3), orient.reshape(-1))
gpp_b200.kitti_records(loc, ang, dd)
poller.debug_scores(boxes[0, 0], dims[0, 0], orient[0, 0], P_inv[0], which=2, with_margin=True)
print('sanitize_small: all paths ran, exact/verified == oracle')
