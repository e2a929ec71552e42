"""Development aid (round 2): kernel-only timings of the polling kernel on the bench configurations, for schedule sweeps
(profiles/r02a_* were made with an earlier version that also timed the round-1 ring kernels, since removed).  Run on the GPU box:  python scripts/gpu_r02_schedules.py > gpurun_out/x.json"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gpp_b200  # noqa: E402
from gpp_b200.utils import synthetic  # noqa: E402


def planes_of(tag):
    return np.load(os.path.join(ROOT, 'road_planes_database', 'road_planes_database_%s.npy' % tag))


def time_mode(poller, args, mode, reps=5):
    ms = []
    for i in range(reps + 1):
        poller.fit_torch(*args, mode=mode)
        torch.cuda.synchronize()
        if i:
            ms.append(poller.last_kernel_ms())
    return float(np.median(ms)), float(np.min(ms))


def main():
    dev = torch.device('cuda', 0)
    poller = gpp_b200.get_poller(0)
    out = {'lib': os.path.basename(gpp_b200._lib.LIB_PATH)}
    quick = '--quick' in sys.argv
    cases = [('C4_4096x100x22k', 4096, '22k'), ('C3_64x100x10k', 64, '10k'), ('1x100x22k', 1, '22k'),
             ('C2_1x100x1k', 1, '1k'), ('16x100x22k', 16, '22k'), ('512x100x22k', 512, '22k'), ('1024x100x10k', 1024, '10k'),
             ('1024x100x1k', 1024, '1k')]
    for name, B, tag in cases:
        pl = planes_of(tag)
        pool = min(B, 256)
        boxes, dims, orient, P_inv = synthetic.synth_detections(pool, 100, pl, seed=3)
        rep = (B + pool - 1) // pool
        tile = lambda a: np.ascontiguousarray(np.tile(a, (rep,) + (1,) * (a.ndim - 1))[:B])  # noqa: E731
        args = [torch.from_numpy(tile(a)).to(dev) for a in (boxes, dims, orient, P_inv.astype(np.float32))]
        poller.set_planes(pl)
        hyp = B * 100.0 * pl.shape[0]
        res = {}
        for mode in ('verified', 'fast', 'exact'):
            med, best = time_mode(poller, args, mode)
            res[mode + '_auto'] = {'ms': med, 'ms_min': best, 'hyp_per_s': hyp / (med * 1e-3)}
        if not quick:
            sweeps = [(1, 0), (1, 64), (1, 128), (1, -1)] if B >= 512 else [(1, 0), (2, 0), (4, 0), (8, 0), (16, 0), (32, 0), (2, -1), (4, -1), (32, -1)]
            for n_seg, resid in sweeps:
                poller.debug_set_schedule(n_seg, resid)
                med, best = time_mode(poller, args, 'verified', reps=3)
                poller.debug_set_schedule(0, -1)
                res['verified_seg%d_res%d' % (n_seg, resid)] = {'ms': med, 'hyp_per_s': hyp / (med * 1e-3)}
        out[name] = res
        print(name, json.dumps(res), file=sys.stderr, flush=True)
    print(json.dumps(out))


if __name__ == '__main__':
    main()
