"""Soak test (run under gpurun): VERIFIED vs EXACT on the GPU, bit for bit, over many seeds, databases and input
flavours (noise levels, padding, far / near objects).  Writes gpurun_out/soak_verified.json."""
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

dev = torch.device('cuda', 0)
total_det = total_hyp = mismatches = 0
t0 = time.time()
report = []
for tag, B in (('22k', 1024), ('10k', 1024), ('1k', 2048), ('100', 2048)):
    planes = np.load(os.path.join(ROOT, 'road_planes_database', 'road_planes_database_%s.npy' % tag))
    for seed in range(int(sys.argv[1]) if len(sys.argv) > 1 else 8):
        kp_noise = (0.0, 1.5, 4.0, 10.0)[seed % 4]
        dim_noise = (0.0, 0.05, 0.2, 0.5)[(seed // 2) % 4]
        nv = (None, 60, 17)[seed % 3]
        boxes, dims, orient, P_inv = synthetic.synth_detections(B, 100, planes, seed=1000 + seed, n_valid=nv,
                                                                kp_noise_px=kp_noise, dim_noise=dim_noise)
        if seed % 5 == 4:                                   # shrink the key-point spread: far-away geometry
            boxes = boxes.copy()
            boxes[:, :, 4:] = (boxes[:, :, 4:] - 650.0) * 0.1 + 650.0
        args = [torch.from_numpy(a).to(dev) for a in (boxes, dims, orient, P_inv.astype(np.float32))]
        ex = gpp_b200.fit_road_planes_torch(*args, planes, mode='exact', return_index=True)
        ve = gpp_b200.fit_road_planes_torch(*args, planes, mode='verified', return_index=True)
        torch.cuda.synchronize()
        bad = int((ex[3] != ve[3]).sum().item())
        same_vals = all(np.array_equal(a.cpu().numpy(), b.cpu().numpy(), equal_nan=True) for a, b in zip(ex[:3], ve[:3]))
        mismatches += bad + (0 if same_vals else 1)
        total_det += B * 100
        total_hyp += B * 100 * planes.shape[0]
        report.append(dict(db=tag, seed=seed, kp_noise=kp_noise, dim_noise=dim_noise, n_valid=nv, index_mismatches=bad,
                           values_identical=same_vals))
res = dict(detections=total_det, hypotheses=total_hyp, mismatches=mismatches, seconds=time.time() - t0, runs=report)
os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
with open(os.path.join(ROOT, 'gpurun_out', 'soak_verified.json'), 'w') as f:
    json.dump(res, f, indent=1)
print('soak: %d detections, %.3e hypotheses, %d mismatches, %.1f s' % (total_det, total_hyp, mismatches, res['seconds']))
