import os, sys
import numpy as np
ROOT = '/root/repo'
sys.path.insert(0, ROOT)
import torch
import gpp_b200
from gpp_b200.utils import synthetic
poller = gpp_b200.get_poller(0)
dev = torch.device('cuda', 0)
for db in ('10k', '22k'):
    planes = np.load(os.path.join(ROOT, 'road_planes_database', 'road_planes_database_%s.npy' % db))
    poller.set_planes(planes)
    for B in (1, 4, 8, 17, 18, 35, 36, 64, 71, 72, 107, 128, 256, 512):
        boxes, dims, orient, P_inv = synthetic.synth_detections(B, 100, planes, seed=3)
        t = [torch.from_numpy(a).to(dev) for a in (boxes, dims, orient, P_inv.astype(np.float32))]
        line = 'db %s B %4d:' % (db, B)
        for mode in ('verified', 'fast', 'exact'):
            best = 1e9
            for i in range(5):
                poller.fit_torch(*t, mode=mode)
                torch.cuda.synchronize()
                if i: best = min(best, poller.last_kernel_ms())
            line += '  %s %.3f ms %.2e' % (mode, best, B * 100 * planes.shape[0] / best * 1e3)
        print(line)
