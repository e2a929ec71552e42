"""Development aid: wall-clock throughput of the single-process multi-GPU numpy entry (fit_road_planes_multi) against
the one-GPU entry, host arrays in / out.  Run under `gpurun --gpus N`."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gpp_b200
from gpp_b200.utils import synthetic
n = gpp_b200._lib.load().gpp_device_count()
per_gpu = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
planes = np.load(os.path.join(ROOT, 'road_planes_database', 'road_planes_database_22k.npy'))
base = synthetic.synth_detections(256, 100, planes, seed=5)
res = {'gpus': n, 'images_per_gpu': per_gpu, 'planes': int(planes.shape[0])}
for g in sorted({1, n}):
    B = per_gpu * g
    boxes, dims, orient, P_inv = [np.ascontiguousarray(np.tile(a, (B // 256,) + (1,) * (a.ndim - 1))) for a in base]
    best = 1e9
    for i in range(4):
        t = time.time()
        out = gpp_b200.fit_road_planes_multi(boxes, dims, orient, P_inv, planes, devices=list(range(g)))
        dt = time.time() - t
        if i: best = min(best, dt)
    res['multi_%d' % g] = {'ms': best * 1e3, 'hyp_per_s': B * 100 * planes.shape[0] / best}
    print('%d GPU(s), %d images: %.1f ms per call, %.3e hyp/s (numpy in/out, pageable host memory)' % (g, B, best * 1e3, res['multi_%d' % g]['hyp_per_s']))
os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, 'gpurun_out', 'multi_driver.json'), 'w'), indent=1)
