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


def test_single_process_multi_device_driver_host_logic():
    """fit_road_planes_multi with an injected compute function (the C oracle) and three pretend devices: shards are
    contiguous, each is handed its own device id and writes into its slice of the shared result arrays."""
    import threading
    from gpp_b200.sharding import fit_road_planes_multi
    from gpp_b200.utils import synthetic
    from oracle import c_oracle
    planes = load_planes('100')
    boxes, dims, orient, P_inv = synthetic.synth_detections(7, 12, planes, seed=23, n_valid=9)
    calls = []
    lock = threading.Lock()

    def fit_fn(b, d, o, p, pl, mode=None, return_index=False, device=None, out=None):
        res = c_oracle.fit_road_planes_c(b, d, o, p, pl, nthreads=1, return_index=return_index)
        for dst, src in zip(out, res):
            assert dst.flags['C_CONTIGUOUS'] and dst.shape == src.shape
            dst[...] = src
        with lock:
            calls.append((device, b.shape[0], threading.current_thread().name))

    got = fit_road_planes_multi(boxes, dims, orient, P_inv, np.tile(planes[None], (7, 1, 1)), devices=[4, 5, 6],
                                return_index=True, fit_fn=fit_fn)
    want = c_oracle.fit_road_planes_c(boxes, dims, orient, P_inv, planes, nthreads=1, return_index=True)
    assert all(np.array_equal(a, b, equal_nan=True) for a, b in zip(got, want))
    assert sorted((c[0], c[1]) for c in calls) == [(4, 3), (5, 2), (6, 2)]
    assert sorted(c[2] for c in calls) == ['gpp-dev4', 'gpp-dev5', 'gpp-dev6']      # one host thread per device
    # more devices than images: the empty shards are not dispatched
    calls.clear()
    got = fit_road_planes_multi(boxes[:2], dims[:2], orient[:2], P_inv[:2], planes, devices=[0, 1, 2, 3], fit_fn=fit_fn)
    assert sorted(c[0] for c in calls) == [0, 1] and len(got) == 3
    assert all(np.array_equal(a[:2], b, equal_nan=True) for a, b in zip(want[:3], got))
    # errors of a worker surface in the caller; per-image databases are refused
    def bad(*a, **k):
        raise RuntimeError('device lost')
    with pytest.raises(RuntimeError, match='device lost'):
        fit_road_planes_multi(boxes, dims, orient, P_inv, planes, devices=[0, 1], fit_fn=bad)
    per_image = np.tile(planes[None], (7, 1, 1))
    per_image[3, 0, 3] += 1.0
    with pytest.raises(ValueError):
        fit_road_planes_multi(boxes, dims, orient, P_inv, per_image, devices=[0], fit_fn=fit_fn)


def test_multi_device_driver_needs_a_gpu():
    import torch
    import gpp_b200
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    assert gpp_b200._lib.load().gpp_device_count() == 0
    planes = load_planes('10')
    with pytest.raises(RuntimeError):
        gpp_b200.fit_road_planes_multi(np.zeros((1, 1, 12), np.float32), np.ones((1, 1, 3), np.float32),
                                       np.zeros((1, 1), np.int32), np.zeros((1, 4, 3), np.float32), planes)
