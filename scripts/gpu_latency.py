"""Development aid: per-call latency of the numpy entry for the reference's own calling pattern
(1 image x 100 rows, ~15 valid, 22k planes re-fed each call).  Run under gpurun."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gpp_b200
from gpp_b200.utils import synthetic
planes = np.load(os.path.join(ROOT, 'road_planes_database', 'road_planes_database_22k.npy'))
for nv in (100, 15):
    boxes, dims, orient, P_inv = synthetic.synth_detections(1, 100, planes, seed=1, n_valid=nv)
    for mode in ('verified', 'exact', 'fast'):
        for i in range(5):
            gpp_b200.fit_road_planes(boxes, dims, orient, P_inv, planes[None], mode=mode)
        t = time.time()
        for i in range(50):
            gpp_b200.fit_road_planes(boxes, dims, orient, P_inv, planes[None], mode=mode)
        dt = (time.time() - t) / 50
        print('valid rows %3d mode %-8s: %.3f ms per call (1 x 100 x 21634)' % (nv, mode, dt * 1e3))
