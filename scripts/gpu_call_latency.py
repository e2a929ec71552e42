"""Development aid: wall-clock latency of the reference's own call pattern -- ONE image, 100 rows, the 22k database
re-fed with every call as a float64 Fortran-ordered (1, N, 4) array (run_network.py:75,105-110) -- through the public
numpy entry, with its parts timed separately."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gpp_b200  # noqa: E402
from gpp_b200.utils import synthetic  # noqa: E402


def bench(fn, n=300, warm=20):
    for _ in range(warm):
        fn()
    t = []
    for _ in range(n):
        t0 = time.perf_counter()
        fn()
        t.append(time.perf_counter() - t0)
    t = np.array(t) * 1e3
    return {'median_ms': float(np.median(t)), 'p10_ms': float(np.percentile(t, 10)), 'p90_ms': float(np.percentile(t, 90))}


def main():
    out = {}
    for tag in ('22k', '1k'):
        db = np.load(os.path.join(ROOT, 'road_planes_database', 'road_planes_database_%s.npy' % tag))
        planes = np.asfortranarray(db.astype(np.float64))              # what scipy.io.loadmat returns
        for n_valid in (100, 15):
            boxes, dims, orient, P_inv = synthetic.synth_detections(1, 100, db, seed=7, n_valid=n_valid)
            P64 = P_inv.astype(np.float64)
            poller = gpp_b200.get_poller(0)
            feed = np.expand_dims(planes, axis=0)
            r = {}
            r['fit_road_planes'] = bench(lambda: gpp_b200.fit_road_planes(boxes, dims, orient, P64, feed))
            r['set_planes_only'] = bench(lambda: poller.set_planes(planes))
            b32, d32, o32, p32 = boxes.astype(np.float32), dims.astype(np.float32), orient.astype(np.int32), P_inv.astype(np.float32)
            r['poller_fit_only'] = bench(lambda: poller.fit(b32, d32, o32, p32))
            r['kernel_ms'] = poller.last_kernel_ms()
            r['return_pose'] = bench(lambda: gpp_b200.fit_road_planes(boxes, dims, orient, P64, feed, return_pose=True))
            out['%s_valid%d' % (tag, n_valid)] = r
            print(tag, n_valid, json.dumps(r), file=sys.stderr, flush=True)
    print(json.dumps(out))


if __name__ == '__main__':
    main()
