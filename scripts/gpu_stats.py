"""Development aid: filter statistics of the resident VERIFIED kernel (needs a -DGPP_STATS build passed through GPP_LIB_PATH)."""
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
for B, tag, nv in ((512, '22k', 100), (64, '10k', 100), (512, '10k', 100), (512, '1k', 100), (1, '22k', 100), (1, '22k', 15)):
    pl = np.load(os.path.join(ROOT, 'road_planes_database', 'road_planes_database_%s.npy' % tag))
    boxes, dims, orient, P_inv = synthetic.synth_detections(min(B, 256), 100, pl, seed=3, n_valid=nv)
    rep = (B + 255) // 256
    tile = lambda a: np.ascontiguousarray(np.tile(a, (rep,) + (1,) * (a.ndim - 1))[:B])  # noqa: E731
    args = [torch.from_numpy(tile(a)).to(dev) for a in (boxes, dims, orient, P_inv.astype(np.float32))]
    poller.set_planes(pl)
    print('== %d x 100 x %s (valid %d)' % (B, tag, nv), file=sys.stderr, flush=True)
    poller.fit_torch(*args, mode='verified')
    torch.cuda.synchronize()
