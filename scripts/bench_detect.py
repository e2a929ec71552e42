"""Measurement of the decode / FilterDetections rows and of the whole post-CNN tail at
the reference's own sizes: 137,268 anchors per 1333x402 image (12 anchors x P3..P7), B images, synthetic heads.
Run under gpurun; writes gpurun_out/bench_detect.json."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import gpp_b200  # noqa: E402

A = sum(((1333 + s - 1) // s) * ((402 + s - 1) // s) for s in (8, 16, 32, 64, 128)) * 12
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
rng = np.random.default_rng(0)
dev = torch.device('cuda', 0)
cx, cy = rng.uniform(0, 1333, A), rng.uniform(0, 402, A)
w, h = rng.uniform(16, 400, A), rng.uniform(16, 300, A)
anchors = torch.from_numpy(np.stack([cx - w / 2, cy - h / 2, cx + w / 2, cy + h / 2], 1).astype(np.float32)).to(dev)
reg = torch.randn((B, A, 12), device=dev)
rdim = torch.randn((B, A, 3), device=dev)
cls = torch.rand((B, A, 8), device=dev) * 0.04
hot = torch.randint(0, A, (B, 300), device=dev)
for b in range(B):
    cls[b, hot[b], torch.randint(0, 8, (300,), device=dev)] = torch.rand(300, device=dev) * 0.9 + 0.06
planes = np.load(os.path.join(ROOT, 'road_planes_database', 'road_planes_database_22k.npy'))
_, P_inv = gpp_b200.synthetic.kitti_calibration()
P_inv = torch.from_numpy(np.tile(P_inv[None].astype(np.float32), (B, 1, 1))).to(dev)
gpp_b200.get_poller(0).set_planes(planes)


def timed(fn, reps=10):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); out = fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best, out


peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json'))) if os.path.exists(os.path.join(ROOT, 'MEASURED_PEAKS.json')) else {}
hbm = peaks.get('hbm_gbs', 6650.0)
ms_dec, (boxes, dims) = timed(lambda: gpp_b200.decode_torch(anchors, reg, cls, rdim))
bytes_dec = B * A * 4 * (12 + 8 + 3 + 12 + 3) + A * 16
ms_fil, det = timed(lambda: gpp_b200.filter_detections_torch(boxes, dims, cls))
bytes_fil = B * A * 4 * 8 + B * A            # classification read + orientation byte written (the dominant traffic)
ms_all, out = timed(lambda: gpp_b200.detections_from_heads(anchors, reg, rdim, cls, P_inv, planes))
ms_pose, out = timed(lambda: gpp_b200.detections_from_heads(anchors, reg, rdim, cls, P_inv, planes, return_pose=True,
                                                          return_kitti=True))
res = {
    'anchors_per_image': A, 'images': B,
    'decode': {'ms': ms_dec, 'anchors_per_s': B * A / ms_dec * 1e3, 'algorithmic_bytes': bytes_dec,
               'achieved_gbs': bytes_dec / ms_dec / 1e6, 'peak_gbs': hbm, 'frac': bytes_dec / ms_dec / 1e6 / hbm},
    'filter': {'ms': ms_fil, 'images_per_s': B / ms_fil * 1e3, 'kept_per_image': float((det[2] >= 0).sum().item()) / B,
               'score_pass_algorithmic_bytes': bytes_fil, 'score_pass_gbs_if_alone': bytes_fil / ms_fil / 1e6},
    'tail_heads_to_polled_detections': {'ms': ms_all, 'images_per_s': B / ms_all * 1e3,
                                        'note': 'decode + filter + polling of 100 rows/image x 21634 planes (verified mode)'},
    'tail_with_pose_and_kitti_record': {'ms': ms_pose, 'images_per_s': B / ms_pose * 1e3,
                                        'note': 'the same with pose recovery and the KITTI record in the polling epilogue (4 launches)'},
    'peak_source': 'MEASURED_PEAKS.json hbm_gbs' if peaks else 'fallback 6650 GB/s',
}
os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
with open(os.path.join(ROOT, 'gpurun_out', 'bench_detect_%d.json' % B), 'w') as f:
    json.dump(res, f, indent=1)
print(json.dumps(res, indent=1))
