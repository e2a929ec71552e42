"""Development aid for ncu captures: a few launches of one mode on a C4-like slice.
    python scripts/gpu_one.py IMAGES MODE [NOISE_PX] [REPS]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import gpp_b200
from gpp_b200.utils import synthetic
B = int(sys.argv[1]); mode = sys.argv[2]
noise = float(sys.argv[3]) if len(sys.argv) > 3 else 1.5
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
poller = gpp_b200.get_poller(0)
dev = torch.device('cuda', 0)
planes = np.load(os.path.join(ROOT, 'road_planes_database', 'road_planes_database_22k.npy'))
poller.set_planes(planes)
boxes, dims, orient, P_inv = synthetic.synth_detections(B, 100, planes, seed=4, kp_noise_px=noise)
t = [torch.from_numpy(a).to(dev) for a in (boxes, dims, orient, P_inv.astype(np.float32))]
for i in range(reps):
    poller.fit_torch(*t, mode=mode)
    torch.cuda.synchronize()
    print('%s %d x 100 x %d: %.3f ms' % (mode, B, planes.shape[0], poller.last_kernel_ms()))
