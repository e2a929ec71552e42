"""Development aid for ncu captures: a few launches of one mode on one (images, database) configuration.
    python scripts/gpu_one_cfg.py IMAGES DB MODE [REPS] [N_VALID]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gpp_b200  # noqa: E402
from gpp_b200.utils import synthetic  # noqa: E402

B, tag, mode = int(sys.argv[1]), sys.argv[2], sys.argv[3]
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
nv = int(sys.argv[5]) if len(sys.argv) > 5 else 100
poller = gpp_b200.get_poller(0)
dev = torch.device('cuda', 0)
planes = np.load(os.path.join(ROOT, 'road_planes_database', 'road_planes_database_%s.npy' % tag))
poller.set_planes(planes)
boxes, dims, orient, P_inv = synthetic.synth_detections(B, 100, planes, seed=11, n_valid=nv)
t = [torch.from_numpy(a).to(dev) for a in (boxes, dims, orient, P_inv.astype(np.float32))]
for i in range(reps):
    poller.fit_torch(*t, mode=mode)
    torch.cuda.synchronize()
    print('%s %d x 100 x %d: %.4f ms' % (mode, B, planes.shape[0], poller.last_kernel_ms()))
