"""VERIFIED == EXACT on adversarial inputs (gpp_b200.utils.adversarial): every detection flavour x plane flavour, many
seeds, bit-for-bit comparison of index, key-points, key-planes and residuals on the GPU.  Prints one JSON summary; the
inputs of mismatching detections are saved to gpurun_out/adversarial_failures.npz for replay against the oracle.

    python scripts/soak_adversarial.py [SEEDS] [IMAGES_PER_CASE] [N_PLANES]"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gpp_b200  # noqa: E402
from gpp_b200.utils import adversarial as adv  # noqa: E402


def main():
    seeds = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    n_img = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    n_planes = int(sys.argv[3]) if len(sys.argv) > 3 else 3000
    dev = torch.device('cuda', 0)
    poller = gpp_b200.get_poller(0)
    bases = {t: np.load(os.path.join(ROOT, 'road_planes_database', 'road_planes_database_%s.npy' % t)) for t in ('10k', '22k')}
    summary, failures = {}, []
    total_det = total_hyp = total_bad = 0
    t0 = time.time()
    for seed in range(seeds):
        rng = np.random.default_rng(1000 + seed)
        base = bases['22k' if seed % 2 else '10k']
        for pf in adv.PLANE_FLAVOURS:
            db = adv.planes(pf, n_planes, rng, base=base)
            poller.set_planes(db)
            for df in adv.DET_FLAVOURS:
                boxes, dims, orient, P_inv = adv.detections(df, n_img, 100, rng, base)
                t = [torch.from_numpy(a).to(dev) for a in (boxes, dims, orient, P_inv)]
                ve = poller.fit_torch(*t, mode='verified', return_index=True)
                ex = poller.fit_torch(*t, mode='exact', return_index=True)
                torch.cuda.synchronize()
                bad = np.zeros(boxes.shape[:2], bool)
                for a, b in zip(ve, ex):
                    a, b = a.cpu().numpy(), b.cpu().numpy()
                    same = (a == b) | ((a != a) & (b != b))
                    bad |= ~same.reshape(same.shape[0], same.shape[1], -1).all(axis=2)
                key = '%s/%s' % (pf, df)
                s = summary.setdefault(key, {'detections': 0, 'mismatches': 0})
                s['detections'] += bad.size
                s['mismatches'] += int(bad.sum())
                total_det += bad.size
                total_hyp += bad.size * db.shape[0]
                total_bad += int(bad.sum())
                for b, d in zip(*np.nonzero(bad)):
                    if len(failures) < 200:
                        failures.append({'key': key, 'seed': seed, 'box': boxes[b, d], 'dims': dims[b, d], 'orient': orient[b, d],
                                         'pinv': P_inv[b], 'planes': db, 'idx_verified': int(ve[3][b, d]), 'idx_exact': int(ex[3][b, d])})
    out = {'detections': total_det, 'hypotheses': total_hyp, 'mismatches': total_bad, 'seconds': time.time() - t0,
           'seeds': seeds, 'planes_per_database': n_planes, 'by_case': summary}
    if failures:
        os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
        np.savez_compressed(os.path.join(ROOT, 'gpurun_out', 'adversarial_failures.npz'),
                            **{'%s_%d' % (k, i): np.asarray(f[k]) for i, f in enumerate(failures[:40]) for k in f if k not in ('key',)},
                            keys=np.array([f['key'] for f in failures[:40]]))
    print(json.dumps(out))


if __name__ == '__main__':
    main()
