"""Development aid: a VERIFIED kernel variant (debug_set_config code, 2 = 2 CTAs per SM, 3 = 3 CTAs per SM) against
EXACT, bit for bit, then its timing against the default variant."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import gpp_b200
from gpp_b200.utils import synthetic
variant = int(sys.argv[1]) if len(sys.argv) > 1 else 2
poller = gpp_b200.get_poller(0)
dev = torch.device('cuda', 0)
bad = 0
for db, B, nv, noise in (('22k', 64, None, 3.0), ('22k', 300, 17, 3.0), ('22k', 256, None, 20.0), ('10k', 128, 60, 1.5),
                         ('1k', 512, None, 3.0), ('100', 600, 33, 3.0), ('10', 700, None, 3.0)):
    planes = np.load(os.path.join(ROOT, 'road_planes_database', 'road_planes_database_%s.npy' % db))
    poller.set_planes(planes)
    boxes, dims, orient, P_inv = synthetic.synth_detections(B, 100, planes, seed=7, n_valid=nv, kp_noise_px=noise)
    t = [torch.from_numpy(a).to(dev) for a in (boxes, dims, orient, P_inv.astype(np.float32))]
    poller.debug_set_config(0, 0)
    ex = [o.cpu().numpy() for o in poller.fit_torch(*t, mode='exact', return_index=True)]
    poller.debug_set_config(200 + variant, 0)          # batch kernel forced
    ve = [o.cpu().numpy() for o in poller.fit_torch(*t, mode='verified', return_index=True)]
    torch.cuda.synchronize()
    same = all(np.array_equal(a, b, equal_nan=True) for a, b in zip(ex, ve))
    bad += 0 if same else 1
    print('db %s B %d n_valid %s noise %.1f: %s' % (db, B, nv, noise, 'identical' if same else 'MISMATCH %d' % int((ex[3] != ve[3]).sum())))
planes = np.load(os.path.join(ROOT, 'road_planes_database', 'road_planes_database_22k.npy'))
poller.set_planes(planes)
for noise in (0.0, 1.5, 4.0, 10.0):
    boxes, dims, orient, P_inv = synthetic.synth_detections(2048, 100, planes, seed=3, kp_noise_px=noise)
    t = [torch.from_numpy(a).to(dev) for a in (boxes, dims, orient, P_inv.astype(np.float32))]
    line = 'noise %.1f:' % noise
    for v in (3, variant):
        poller.debug_set_config(v, 0)
        best = 1e9
        for i in range(4):
            poller.fit_torch(*t, mode='verified')
            torch.cuda.synchronize()
            if i: best = min(best, poller.last_kernel_ms())
        line += '  variant %d %.3f ms %.3e' % (v, best, 2048 * 100 * planes.shape[0] / best * 1e3)
    print(line)
poller.debug_set_config(0, 0)
print('mismatching cases:', bad)
