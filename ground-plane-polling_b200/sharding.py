"""
Multi-GPU sharding of the polling path: one process per GPU, the image axis is split
into contiguous shards, the plane database is replicated on every GPU, and there is NO collective on the
data path -- every (image, detection) is independent.  ``torch.distributed`` only provides rank / world size
and the optional result gather to rank 0.
"""
import numpy as np

__all__ = ['shard_bounds', 'fit_road_planes_sharded']


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
