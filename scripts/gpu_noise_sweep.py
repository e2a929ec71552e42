"""Development aid: how the VERIFIED / FAST / EXACT kernel times depend on the detection noise (run under gpurun)."""
import os, sys
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
B = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
modes = sys.argv[2:] or ['verified', 'fast']
for noise in (0.0, 0.5, 1.5, 4.0, 10.0, 40.0):
    boxes, dims, orient, P_inv = synthetic.synth_detections(B, 100, planes, seed=3, kp_noise_px=noise)
    tb, td, to, tp = [torch.from_numpy(a).to(dev) for a in (boxes, dims, orient, P_inv.astype(np.float32))]
    line = 'noise %5.1f px:' % noise
    for mode in modes:
        best = 1e9
        for i in range(4):
            out = poller.fit_torch(tb, td, to, tp, mode=mode)
            torch.cuda.synchronize()
            if i: best = min(best, poller.last_kernel_ms())
        line += '  %s %.3f ms %.3e hyp/s' % (mode, best, B * 100 * planes.shape[0] / best * 1e3)
    res = out[2].cpu().numpy()
    line += '  | sentinel rows %.1f%%  median residual %.3f' % (100.0 * np.mean(res > 16.0), float(np.median(res)))
    print(line)
