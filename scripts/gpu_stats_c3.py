"""Development aid: filter statistics of C3 under different segment counts (needs a -DGPP_STATS build via GPP_LIB_PATH)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gpp_b200  # noqa: E402
from gpp_b200.utils import synthetic  # noqa: E402

dev = torch.device('cuda', 0)
poller = gpp_b200.get_poller(0)
pl = np.load(os.path.join(ROOT, 'road_planes_database', 'road_planes_database_10k.npy'))
boxes, dims, orient, P_inv = synthetic.synth_detections(64, 100, pl, seed=11)
args = [torch.from_numpy(a).to(dev) for a in (boxes, dims, orient, P_inv.astype(np.float32))]
poller.set_planes(pl)
for n_seg in (1, 2, 3, 6):
    poller.debug_set_schedule(n_seg, -1)
    print('== n_seg', n_seg, file=sys.stderr, flush=True)
    poller.fit_torch(*args, mode='verified')
    torch.cuda.synchronize()
poller.debug_set_schedule(0, -1)
