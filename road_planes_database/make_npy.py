"""Regenerate the .npy copies of the reference's road-plane databases.

The reference ships five MATLAB v5 files under /root/reference/road_planes_database/ (variable
``road_planes_database``, float64 (N, 4) rows [a, b, c, d]; loaded by
keras_retinanet_3D/bin/run_network.py:75 and preprocessing/kitti.py:88).  /root/reference does not exist on
the GPU box, so the raw values are re-saved here, bit for bit (float64, C order), as .npy data fixtures.
Run in the build container only:  python road_planes_database/make_npy.py
"""
import hashlib
import os
import sys

import numpy as np
import scipy.io

SRC = '/root/reference/road_planes_database'
DST = os.path.dirname(os.path.abspath(__file__))

if __name__ == '__main__':
    lines = []
    for tag in ('10', '100', '1k', '10k', '22k'):
        src = os.path.join(SRC, 'road_planes_database_%s.mat' % tag)
        with open(src, 'rb') as f:
            sha = hashlib.sha256(f.read()).hexdigest()
        db = scipy.io.loadmat(src)['road_planes_database']
        assert db.dtype == np.float64 and db.ndim == 2 and db.shape[1] == 4
        out = np.ascontiguousarray(db)
        np.save(os.path.join(DST, 'road_planes_database_%s.npy' % tag), out)
        lines.append('%s  N=%d  mat_sha256=%s  values_sha256=%s' % (
            tag, out.shape[0], sha, hashlib.sha256(out.tobytes()).hexdigest()))
    with open(os.path.join(DST, 'MANIFEST.txt'), 'w') as f:
        f.write('\n'.join(lines) + '\n')
    sys.stdout.write('\n'.join(lines) + '\n')
