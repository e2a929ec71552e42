"""Development aid: time the FAST kernel variants on C4-like slices (run under gpurun)."""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import gpp_b200
from gpp_b200.utils import synthetic
poller = gpp_b200.get_poller(0)
dev = torch.device('cuda', 0)
planes = np.load(os.path.join(ROOT, 'road_planes_database', 'road_planes_database_22k.npy'))
poller.set_planes(planes)
boxes, dims, orient, P_inv = synthetic.synth_detections(256, 100, planes, seed=3)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
rep = (B + 255) // 256
tile = lambda a: np.tile(a, (rep,) + (1,) * (a.ndim - 1))[:B]
tb, td, to, tp = [torch.from_numpy(tile(a)).to(dev) for a in (boxes, dims, orient, P_inv.astype(np.float32))]
for mode in sys.argv[2:] or ['fast']:
    for variant in (2, 3, 4):
        for cps in (0,):
            poller.debug_set_config(variant, cps)
            best = 1e9
            for i in range(4):
                poller.fit_torch(tb, td, to, tp, mode=mode)
                torch.cuda.synchronize()
                if i: best = min(best, poller.last_kernel_ms())
            print('%s variant %d cps %d: %.3f ms  %.4e hyp/s' % (mode, variant, cps, best, B * 100 * planes.shape[0] / best * 1e3))
poller.debug_set_config(0, 0)
