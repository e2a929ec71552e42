"""CPU, world_size 2, gloo: the host-side sharding logic of the multi-GPU path (no CUDA involved -- the
compute function is injected; here the C oracle stands in as the checker-side compute)."""
import os
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

from conftest import ROOT, load_planes


def test_shard_bounds_partition():
    from gpp_b200.sharding import shard_bounds
    for n in (0, 1, 7, 64, 4096, 4097):
        for w in (1, 2, 3, 8):
            spans = [shard_bounds(n, w, r) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    import torch.distributed as dist
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from gpp_b200.sharding import fit_road_planes_sharded
        from gpp_b200.utils import synthetic
        from oracle import c_oracle
        planes = load_planes('100')
        boxes, dims, orient, P_inv = synthetic.synth_detections(5, 20, planes, seed=17)

        def fit_fn(b, d, o, p, pl, mode=None):
            return c_oracle.fit_road_planes_c(b, d, o, p, pl, nthreads=1)

        out = fit_road_planes_sharded(boxes, dims, orient, P_inv, planes, fit_fn=fit_fn)
        b0, b1, part = fit_road_planes_sharded(boxes, dims, orient, P_inv, planes, gather=False, fit_fn=fit_fn)
        if rank == 0:
            want = c_oracle.fit_road_planes_c(boxes, dims, orient, P_inv, planes, nthreads=1)
            ok = all(np.array_equal(a, b, equal_nan=True) for a, b in zip(out, want))
            q.put(('full', ok))
        else:
            q.put(('none', out is None))
        q.put(('span', (rank, b0, b1, part[0].shape[0])))
    finally:
        dist.destroy_process_group()


def test_sharded_equals_single_process_gloo_world2():
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=180)
        assert p.exitcode == 0
    got = [q.get(timeout=5) for _ in range(4)]
    d = {}
    spans = []
    for k, v in got:
        if k == 'span':
            spans.append(v)
        else:
            d[k] = v
    assert d == {'full': True, 'none': True}
    spans.sort()
    assert [(s[1], s[2]) for s in spans] == [(0, 3), (3, 5)]
    assert all(s[3] == s[2] - s[1] for s in spans)
