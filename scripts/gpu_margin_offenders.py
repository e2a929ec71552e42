"""Development aid: the hypotheses with the largest |fast - exact| / margin (see gpu_margin_pressure.py), dumped with
their inputs to gpurun_out/margin_offenders.npz for analysis on the CPU."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gpp_b200  # noqa: E402
from gpp_b200.utils import adversarial as adv  # noqa: E402

N_DET = int(sys.argv[1]) if len(sys.argv) > 1 else 30
position = 0
poller = gpp_b200.get_poller(0)
base = np.load(os.path.join(ROOT, 'road_planes_database', 'road_planes_database_10k.npy'))
rng = np.random.default_rng(4242)
rows = []
for pf in adv.PLANE_FLAVOURS:
    db = adv.planes(pf, 4000, rng, base=base)
    poller.set_planes(db)
    norm = poller.normalised_planes()
    for df in adv.DET_FLAVOURS:
        boxes, dims, orient, P_inv = adv.detections(df, 2, N_DET, rng, base)
        for b in range(2):
            for d in range(N_DET):
                ev, er, ez = poller.debug_scores(boxes[b, d], dims[b, d], orient[b, d], P_inv[b], which=0)
                fv, fr, fz, fm, vhi, zok = poller.debug_scores(boxes[b, d], dims[b, d], orient[b, d], P_inv[b], which=1, with_margin=True)
                fin = np.isfinite(er) & np.isfinite(fr) & np.isfinite(fm) & (fm > 0)
                ratio = np.where(fin, np.abs(fr - er) / np.where(fin, fm, 1), 0)
                j = int(np.argmax(ratio))
                if ratio[j] > 0.5:
                    rows.append(dict(flavour='%s/%s' % (pf, df), ratio=ratio[j], box=boxes[b, d], dims=dims[b, d], orient=orient[b, d],
                                     pinv=P_inv[b], plane=norm[j], raw=db[j], fast=fr[j], exact=er[j], margin=fm[j], ev=ev[j], fv=fv[j]))
rows.sort(key=lambda r: -r['ratio'])
rows = rows[:60]
if not rows:
    print('no hypothesis above half its margin')
    sys.exit(0)
np.savez(os.path.join(ROOT, 'gpurun_out', 'margin_offenders.npz'),
         **{k: np.array([r[k] for r in rows]) for k in rows[0]})
for r in rows[:20]:
    print('%-22s ratio %.3f fast %.6g exact %.6g margin %.3g votes %d/%d' % (r['flavour'], r['ratio'], r['fast'], r['exact'], r['margin'], r['fv'], r['ev']))
