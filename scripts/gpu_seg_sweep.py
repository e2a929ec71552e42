"""Development aid: kernel time of small / mid-size batches against the number of plane segments per detection."""
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
CASES = ((64, '10k'), (16, '22k'), (4, '22k'), (1, '22k'), (128, '10k'))
if len(sys.argv) > 1:          # e.g. 512:22k 256:22k
    CASES = tuple((int(a.split(':')[0]), a.split(':')[1]) for a in sys.argv[1:])
SEGS = (0, 5, 6, 7, 8) if len(sys.argv) > 1 else (0, 1, 2, 3, 4, 6, 8, 12, 16, 24, 32)
for B, tag in CASES:
    pl = np.load(os.path.join(ROOT, 'road_planes_database', 'road_planes_database_%s.npy' % tag))
    boxes, dims, orient, P_inv = synthetic.synth_detections(B, 100, pl, seed=11)
    args = [torch.from_numpy(a).to(dev) for a in (boxes, dims, orient, P_inv.astype(np.float32))]
    poller.set_planes(pl)
    line = []
    for mode in ('verified', 'fast'):
        for n_seg in SEGS:
            poller.debug_set_schedule(n_seg, -1)
            ms = []
            for i in range(7):
                poller.fit_torch(*args, mode=mode)
                torch.cuda.synchronize()
                if i > 1:
                    ms.append(poller.last_kernel_ms())
            line.append('%s/%d: %.4f' % (mode[0], n_seg, float(np.median(ms))))
    poller.debug_set_schedule(0, -1)
    print('%d x 100 x %s  ' % (B, tag) + '  '.join(line), flush=True)
