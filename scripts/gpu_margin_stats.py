"""Development aid: distribution of |fast - exact| / margin over many hypotheses (run under gpurun)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gpp_b200
from gpp_b200.utils import synthetic
poller = gpp_b200.get_poller(0)
for tag in ('22k', '10k', '1k'):
    planes = np.load(os.path.join(ROOT, 'road_planes_database', 'road_planes_database_%s.npy' % tag))
    poller.set_planes(planes)
    boxes, dims, orient, P_inv = synthetic.synth_detections(4, 100, planes, seed=77)
    worst, worst_r, nrel = 0.0, 0.0, 0
    loose_viol = 0
    for b in range(4):
        for d in range(100):
            fv, fr, fz, fm, _, _ = poller.debug_scores(boxes[b, d], dims[b, d], orient[b, d], P_inv[b], which=2, with_margin=True)
            ev, er, ez = poller.debug_scores(boxes[b, d], dims[b, d], orient[b, d], P_inv[b], which=0)
            rel = np.isfinite(er) & (ev == 6)
            if rel.any():
                ratio = np.abs(fr[rel] - er[rel]) / fm[rel]
                worst = max(worst, float(ratio.max()))
                nrel += int(rel.sum())
    print('%s: exact six-vote hypotheses %d, worst |fast-exact|/margin = %.4f (K = 64 -> deviation <= %.1f u-scale units)' % (tag, nrel, worst, worst * 64))
