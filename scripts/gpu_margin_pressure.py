"""How close does |fast - exact| come to the margin the VERIFIED filter relies on?  Per-hypothesis scores of adversarial
detections (gpp_b200.utils.adversarial) from the search loops' own device functions (gpp_debug_scores): the largest
|fast residual sum - exact residual sum| / margin over all finite hypotheses, per flavour, for the general form
(which = 1) and the all-six form with the merged reciprocal (which = 2), plus how often the filters' vote / z-check
bounds hold.  A ratio above 1 would be a plane the filter may drop wrongly.

    python scripts/gpu_margin_pressure.py [DETECTIONS_PER_CASE]"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gpp_b200  # noqa: E402
from gpp_b200.utils import adversarial as adv  # noqa: E402


def main():
    n_det = int(sys.argv[1]) if len(sys.argv) > 1 else 60
    poller = gpp_b200.get_poller(0)
    base = np.load(os.path.join(ROOT, 'road_planes_database', 'road_planes_database_10k.npy'))
    rng = np.random.default_rng(4242)
    out = {}
    for pf in adv.PLANE_FLAVOURS:
        db = adv.planes(pf, 4000, rng, base=base)
        poller.set_planes(db)
        for df in adv.DET_FLAVOURS:
            boxes, dims, orient, P_inv = adv.detections(df, 2, max(1, n_det // 2), rng, base)
            worst = {1: 0.0, 2: 0.0}
            n_fin = vote_viol = z_viol = 0
            for b in range(boxes.shape[0]):
                for d in range(boxes.shape[1]):
                    ev, er, ez = poller.debug_scores(boxes[b, d], dims[b, d], orient[b, d], P_inv[b], which=0)
                    for which in (1, 2):
                        fv, fr, fz, fm, vhi, zok = poller.debug_scores(boxes[b, d], dims[b, d], orient[b, d], P_inv[b],
                                                                       which=which, with_margin=True)
                        fin = np.isfinite(er) & np.isfinite(fr) & np.isfinite(fm) & (fm > 0)
                        if fin.any():
                            worst[which] = max(worst[which], float((np.abs(fr[fin] - er[fin]) / fm[fin]).max()))
                        if which == 1:
                            n_fin += int(fin.sum())
                            vote_viol += int((vhi[fin] < ev[fin]).sum())
                            z_viol += int((~zok[fin] & ~ez[fin]).sum())
            out['%s/%s' % (pf, df)] = {'finite_hypotheses': n_fin, 'worst_ratio_general': worst[1], 'worst_ratio_all_six': worst[2],
                                       'vote_bound_violations': vote_viol, 'z_bound_violations': z_viol}
            print(pf, df, json.dumps(out['%s/%s' % (pf, df)]), file=sys.stderr, flush=True)
    out['_worst'] = max(max(v['worst_ratio_general'], v['worst_ratio_all_six']) for v in out.values())
    print(json.dumps(out))


if __name__ == '__main__':
    main()
