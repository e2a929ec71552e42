import sys, numpy as np
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
from conftest import load_planes
import gpp_b200
from gpp_b200.utils import synthetic
poller = gpp_b200.get_poller(0)
planes = load_planes('10k')[:4000]
boxes, dims, orient, P_inv = synthetic.synth_detections(2, 60, planes, seed=505)
dims = dims.copy(); dims[0, :10] *= 12.0; dims[0, 10:20] *= 0.05
boxes = boxes.copy(); boxes[1, :10, 4:] = boxes[1, :10, 4:] * 0.02 + 650.0
poller.set_planes(planes)
for b, d in ((0, 0), (0, 5), (0, 12), (1, 3), (1, 30), (0, 40)):
    for which in (1, 2):
        fv, fr, fz, fm, vhi, zok = poller.debug_scores(boxes[b, d], dims[b, d], orient[b, d], P_inv[b], which=which, with_margin=True)
        ev, er, ez = poller.debug_scores(boxes[b, d], dims[b, d], orient[b, d], P_inv[b], which=0)
        fin = np.isfinite(er) & np.isfinite(fr)
        ratio = np.abs(fr[fin] - er[fin]) / fm[fin]
        j = np.argmax(ratio); jj = np.flatnonzero(fin)[j]
        print((b, d), which, 'max ratio %.3f at plane %d: exact R %.6g fast R %.6g margin %.3g; n viol %d of %d' % (ratio[j], jj, er[jj], fr[jj], fm[jj], (ratio > 1).sum(), fin.sum()), 'plane', planes[jj])
