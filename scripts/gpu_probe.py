"""GPU probe (development aid, run under gpurun): FP32 pipe microbenchmarks + kernel timing sweep.
Writes gpurun_out/probe.json."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import gpp_b200  # noqa: E402
from gpp_b200.utils import synthetic  # noqa: E402

out = {}
poller = gpp_b200.get_poller(0)
kinds = {0: 'ffma', 1: 'ffma2', 2: 'fmul+fadd', 3: 'mufu.rcp', 4: 'mufu.rsq', 5: 'ffma+alu', 6: 'sqrt.approx',
         7: 'fmul2', 8: 'ffma+rcp(4:1)', 9: 'fadd2', 10: 'fmul2|fadd2', 11: 'ffma2|fadd2', 12: 'ffma2 3x64b regs', 13: 'ffma2 bcast'}
mb = {}
for k, name in kinds.items():
    r = poller.microbench(k)
    mb[name] = r
    print('%-14s %8.2f Gops/s  %6.1f ops/clk/SM  %.3f ms' % (name, r['ops_per_s'] / 1e9, r['ops_per_clk_sm'], r['ms']))
out['microbench'] = mb

dev = torch.device('cuda', 0)


def timeit(B, D, tag, mode, dpw=0, cps=0, reps=3):
    planes = np.load(os.path.join(ROOT, 'road_planes_database', 'road_planes_database_%s.npy' % tag))
    boxes, dims, orient, P_inv = synthetic.synth_detections(min(B, 256), D, planes, seed=3)
    rep = (B + boxes.shape[0] - 1) // boxes.shape[0]
    tile = lambda a: np.tile(a, (rep,) + (1,) * (a.ndim - 1))[:B]  # noqa: E731
    tb, td, to, tp = [torch.from_numpy(tile(a)).to(dev) for a in (boxes, dims, orient, P_inv.astype(np.float32))]
    poller.set_planes(planes)
    poller.debug_set_config(dpw, cps)
    best = 1e30
    for i in range(reps + 1):
        poller.fit_torch(tb, td, to, tp, mode=mode)
        torch.cuda.synchronize()
        ms = poller.last_kernel_ms()
        if i > 0:
            best = min(best, ms)
    poller.debug_set_config(0, 0)
    hyp = B * D * planes.shape[0]
    return dict(B=B, D=D, N=int(planes.shape[0]), mode=mode, dpw=dpw, ctas_per_sm=cps, ms=best, hyp_per_s=hyp / (best * 1e-3))


runs = []
for mode in ('exact', 'fast', 'f64'):
    for (B, tag) in ((64, '10k'), (512, '22k')):
        for dpw in ((2, 3, 4) if mode != 'f64' else (0,)):
            for cps in (0,):
                if mode == 'f64' and cps not in (0, 2):
                    continue
                r = timeit(B if mode != 'f64' else B // 4, 100, tag, mode, dpw, cps)
                runs.append(r)
                print(json.dumps(r))
r = timeit(4096, 100, '22k', 'fast', 0, 0, reps=2)
runs.append(r)
print('C4 fast', json.dumps(r))
r = timeit(4096, 100, '22k', 'exact', 0, 0, reps=2)
runs.append(r)
print('C4 exact', json.dumps(r))
out['runs'] = runs
os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
with open(os.path.join(ROOT, 'gpurun_out', 'probe.json'), 'w') as f:
    json.dump(out, f, indent=1)
