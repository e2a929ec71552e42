"""Development aid: where the end-to-end time of a C4 host call goes (pinned buffers): wall clock, sum of the chunk kernels,
one launch on resident data."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import gpp_b200  # noqa: E402

planes = bench.load_planes('22k')
B, D = 4096, 100
boxes, dims, orient, P_inv = bench.make_workload(B, D, planes, seed=3)
poller = gpp_b200.get_poller(0)
poller.set_planes(planes)
dev = torch.device('cuda', 0)


def pinned(a):
    t = torch.empty(a.shape, dtype=torch.from_numpy(a).dtype, pin_memory=True)
    t.numpy()[...] = a
    return t


pin = [pinned(a) for a in (boxes, dims, orient, P_inv)]
outs = [torch.empty(s, dtype=torch.float32, pin_memory=True).numpy() for s in ((B, D, 4, 3), (B, D, 1, 4), (B, D))]
ins = [t.numpy() for t in pin]
for _ in range(3):
    poller.fit(*ins, out=outs)
wall, ksum = [], []
for _ in range(10):
    t0 = time.perf_counter()
    poller.fit(*ins, out=outs)
    wall.append(1e3 * (time.perf_counter() - t0))
    ksum.append(poller.last_kernel_ms())
t = [torch.from_numpy(a).to(dev) for a in (boxes, dims, orient, P_inv)]
one = []
for _ in range(5):
    poller.fit_torch(*t)
    torch.cuda.synchronize()
    one.append(poller.last_kernel_ms())
print('wall %.3f ms   sum of chunk kernels %.3f ms   one launch %.3f ms' % (np.median(wall), np.median(ksum), np.median(one)))
