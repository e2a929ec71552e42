"""
Multi-GPU sharding of the polling path: the image axis is split into contiguous shards, the
plane database is replicated on every GPU, and there is NO collective on the data path -- every (image, detection)
is independent.

  fit_road_planes_sharded   one process per GPU (torchrun); ``torch.distributed`` only provides rank / world size and
                            the optional result gather to rank 0
  fit_road_planes_multi     one process, one host thread + one libgpp handle per GPU; results land in disjoint slices
                            of one set of host arrays (the numpy drop-in for a multi-GPU box)
"""
import numpy as np

__all__ = ['shard_bounds', 'fit_road_planes_sharded', 'fit_road_planes_multi']


def shard_bounds(n_images, world_size, rank):
    """Contiguous, balanced split of [0, n_images): the first (n_images % world_size) ranks get one extra."""
    base, extra = divmod(int(n_images), int(world_size))
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def fit_road_planes_sharded(boxes, dimensions, orientations, P_inv, planes, mode=None, gather=True, fit_fn=None,
                            group=None):
    """Every rank passes the SAME full batch; each polls its own shard of images on its own GPU.

    gather=True : rank 0 returns the full [keypoints, keyplanes, residuals] (other ranks return None);
    gather=False: every rank returns (start, stop, [its shard's outputs]).
    ``fit_fn`` defaults to ``gpp_b200.fit_road_planes`` (tests inject a stand-in to exercise the host logic
    on CPU with the gloo backend).
    """
    import torch
    import torch.distributed as dist
    if fit_fn is None:
        from .layers.fit_road_planes import fit_road_planes as fit_fn
    if dist.is_available() and dist.is_initialized():
        rank, world = dist.get_rank(group), dist.get_world_size(group)
    else:
        rank, world = 0, 1
    boxes = np.asarray(boxes)
    B = boxes.shape[0]
    b0, b1 = shard_bounds(B, world, rank)
    planes = np.asarray(planes)
    pl = planes if planes.ndim == 2 or planes.shape[0] == 1 else planes[b0:b1]
    outs = fit_fn(boxes[b0:b1], np.asarray(dimensions)[b0:b1], np.asarray(orientations)[b0:b1],
                  np.asarray(P_inv)[b0:b1], pl, mode=mode)
    if not gather:
        return b0, b1, outs
    if world == 1:
        return outs
    gathered = [None] * world if rank == 0 else None
    dist.gather_object((b0, b1, outs), gathered, dst=0, group=group)
    if rank != 0:
        return None
    gathered.sort(key=lambda t: t[0])
    return [np.concatenate([g[2][i] for g in gathered], axis=0) for i in range(len(outs))]


def fit_road_planes_multi(boxes, dimensions, orientations, P_inv, planes, devices=None, mode=None, return_index=False,
                          fit_fn=None, out=None):
    """``fit_road_planes`` over several GPUs of one box from ONE process: device ``devices[i]`` polls the i-th
    contiguous shard of images (its own handle, stream and copy of the database) and writes straight into the shard's
    slice of the result arrays.  The fan-out runs inside libgpp (``gpp_fit_host_multi``: one C++ host thread per
    device); an injected ``fit_fn`` is driven by Python threads instead.

    ``devices`` defaults to every visible GPU; ``out`` may hold preallocated result arrays for the WHOLE batch (e.g.
    pinned host memory), each device then writes its shard's slice of them.  ``planes`` is one database shared by the batch ((N, 4) or (1, N, 4) or
    a (B, N, 4) tile of one database).  ``fit_fn(boxes, dims, orient, P_inv, planes, mode=, return_index=, device=,
    out=)`` defaults to ``gpp_b200.fit_road_planes`` (tests inject a stand-in to exercise the host logic on CPU).
    """
    import threading
    if devices is None:
        from . import _lib
        devices = list(range(_lib.load().gpp_device_count()))
    devices = list(devices)
    if not devices:
        raise RuntimeError('fit_road_planes_multi: no CUDA device; libgpp has no CPU fallback')
    boxes = np.asarray(boxes)
    if boxes.ndim != 3:
        raise ValueError('boxes must have shape (B, D, 12), got %r' % (boxes.shape,))
    dimensions, orientations, P_inv = np.asarray(dimensions), np.asarray(orientations), np.asarray(P_inv)
    planes = np.asarray(planes)
    if planes.ndim == 3:
        from .layers.fit_road_planes import _plane_groups
        if len(_plane_groups(planes, boxes.shape[0])) != 1:
            raise ValueError('fit_road_planes_multi takes one plane database shared by the batch')
        planes = planes[0]
    B, D = boxes.shape[:2]
    out_t = np.float64 if mode == 'f64' else np.float32
    want = [((B, D, 4, 3), out_t), ((B, D, 1, 4), out_t), ((B, D), out_t)] + ([((B, D), np.int64)] if return_index else [])
    if out is None:
        out = [np.empty(shape, dt) for shape, dt in want]
    else:
        out = list(out)
        if len(out) != len(want) or any(not isinstance(a, np.ndarray) or a.shape != shape or a.dtype != dt or
                                        not a.flags['C_CONTIGUOUS'] for a, (shape, dt) in zip(out, want)):
            raise ValueError('out must hold C-contiguous arrays of shapes %r' % ([w[0] for w in want],))
    if fit_fn is None and mode != 'f64':
        return _fit_multi_native(boxes, dimensions, orientations, P_inv, planes, devices, mode, return_index, out)
    if fit_fn is None:
        from .layers.fit_road_planes import fit_road_planes as fit_fn
    shards = [(dev,) + shard_bounds(B, len(devices), i) for i, dev in enumerate(devices)]
    shards = [s for s in shards if s[2] > s[1]]

    def run(shard):
        dev, b0, b1 = shard
        fit_fn(boxes[b0:b1], dimensions[b0:b1], orientations[b0:b1], P_inv[b0:b1], planes, mode=mode,
               return_index=return_index, device=dev, out=[o[b0:b1] for o in out])

    if len(shards) == 1:
        run(shards[0])
    elif shards:
        errors = []

        def guarded(shard):
            try:
                run(shard)
            except BaseException as e:  # noqa: B902 -- re-raised in the caller's thread below
                errors.append(e)

        threads = [threading.Thread(target=guarded, args=(s,), name='gpp-dev%d' % s[0]) for s in shards]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        if errors:
            raise errors[0]
    return out


def _fit_multi_native(boxes, dimensions, orientations, P_inv, planes, devices, mode, return_index, out):
    """The fan-out in C (gpp_fit_host_multi): one ctypes call for the whole batch."""
    import ctypes
    from . import _lib
    from .layers.fit_road_planes import DEFAULT_MODE, _f32, get_poller
    mode = DEFAULT_MODE if mode is None else mode
    if mode not in _lib.MODES:
        raise ValueError('unknown mode %r (expected one of %s)' % (mode, sorted(_lib.MODES)))
    B, D = boxes.shape[:2]
    if boxes.shape[2] != 12 or dimensions.shape != (B, D, 3) or orientations.shape != (B, D) or P_inv.shape != (B, 4, 3):
        raise ValueError('inconsistent shapes: boxes %r dimensions %r orientations %r P_inv %r' % (
            boxes.shape, dimensions.shape, orientations.shape, P_inv.shape))
    pollers = [get_poller(d) for d in devices]
    # the database goes along as the caller holds it (float64 / Fortran order from loadmat included): every device compares
    # it with what it holds -- and uploads it if it differs -- from its shard's thread
    pl = np.asarray(planes)
    if pl.ndim != 2 or pl.shape[1] != 4 or pl.shape[0] < 1:
        raise ValueError('planes must have shape (N, 4) with N >= 1, got %r' % (pl.shape,))
    if pl.dtype not in (np.float32, np.float64) or not (pl.flags['C_CONTIGUOUS'] or pl.flags['F_CONTIGUOUS']):
        pl = _f32(pl)
    b, d, p_inv = _f32(boxes), _f32(dimensions), _f32(P_inv)
    o = np.ascontiguousarray(orientations, dtype=np.int32)
    handles = (ctypes.c_void_p * len(pollers))(*[p._h for p in pollers])
    lib = _lib.load()
    rc = lib.gpp_fit_host_multi_planes(handles, len(pollers), pl.ctypes.data, pl.shape[0], 1 if pl.dtype == np.float64 else 0,
                                       0 if pl.flags['C_CONTIGUOUS'] else 1, _lib.ptr(b), _lib.ptr(d), _lib.ptr(o),
                                       _lib.ptr(p_inv), B, D, _lib.ptr(out[0]), _lib.ptr(out[1]), _lib.ptr(out[2]),
                                       _lib.ptr(out[3]) if return_index else None, _lib.MODES[mode])
    for p in pollers:
        p._dev_planes = None
    _lib.check(rc, 'gpp_fit_host_multi_planes')
    return out
