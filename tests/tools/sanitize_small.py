"""Small run of every kernel path for compute-sanitizer (memcheck / racecheck / synccheck):
    compute-sanitizer --tool memcheck python tests/tools/sanitize_small.py   (it checks against the oracle, hence under tests/)"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import gpp_b200
from gpp_b200.utils import synthetic
from oracle import c_oracle

planes = np.load(os.path.join(ROOT, 'road_planes_database', 'road_planes_database_10k.npy'))[:2500]
poller = gpp_b200.get_poller(0)
for (B, D, nv, seg, resid) in ((2, 20, 15, 0, -1), (40, 50, None, 1, 7), (3, 9, 7, 5, 0), (200, 30, 25, 0, -1)):
    boxes, dims, orient, P_inv = synthetic.synth_detections(B, D, planes, seed=B, n_valid=nv)
    dims = dims.copy(); dims[0, :3, 1] *= 1.7
    want = c_oracle.fit_road_planes_c(boxes, dims, orient, P_inv, planes, return_index=True)
    poller.debug_set_schedule(seg, resid)
    for mode in ('verified', 'exact', 'fast', 'f64'):
        got = gpp_b200.fit_road_planes(boxes, dims, orient, P_inv, planes, mode=mode, return_index=True)
        if mode in ('verified', 'exact'):
            assert all(np.array_equal(g, w, equal_nan=True) for g, w in zip(got, want)), (mode, B, D)
    gpp_b200.fit_road_planes(boxes, dims, orient, P_inv, planes, return_pose=True, return_kitti=True)
    poller.debug_set_schedule(0, -1)
poller.audit_set(3)
gpp_b200.fit_road_planes(boxes, dims, orient, P_inv, planes)
poller.audit_set(0)
assert poller.audit_counts()[1] == 0
kp = got[0].astype(np.float32).reshape(-1, 12)
loc, ang, dd = gpp_b200.recover_pose(kp, dims.reshape(-1, 3), orient.reshape(-1))
gpp_b200.kitti_records(loc, ang, dd)
poller.debug_scores(boxes[0, 0], dims[0, 0], orient[0, 0], P_inv[0], which=2, with_margin=True)
# the steps before polling and the device pipeline (decode, FilterDetections, heads -> polled detections)
import torch
sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden'))
from gpp_b200.utils.anchors import anchors_for_shape
from oracle import detect_ref
rng = np.random.default_rng(3)
anchors = anchors_for_shape((96, 160)).astype(np.float32)
A = anchors.shape[0]
reg = rng.normal(0, 1, (2, A, 12)).astype(np.float32)
rdim = rng.normal(0, 1, (2, A, 3)).astype(np.float32)
cls = (rng.random((2, A, 8)) ** 6).astype(np.float32)
b_, d_ = gpp_b200.decode(anchors, reg, cls, rdim)
wb, wd = detect_ref.decode_ref(anchors, reg, cls, rdim)
assert np.array_equal(b_, wb) and np.array_equal(d_, wd)
got = gpp_b200.filter_detections_batch(b_, d_, cls)
want = detect_ref.filter_detections_ref(b_, d_, cls)
assert all(np.array_equal(g, w) for g, w in zip(got, want))
dev = torch.device('cuda', 0)
outs = gpp_b200.detections_from_heads(*[torch.from_numpy(a).to(dev) for a in (anchors, reg, rdim, cls)],
                                      torch.from_numpy(P_inv[:2].astype(np.float32)).to(dev), planes)
torch.cuda.synchronize()
assert outs[5].shape == (2, 100, 4, 3)
# chunked host path through the pinned staging blocks (pageable memory, > 65536 detections)
big = synthetic.synth_detections(700, 100, planes[:64], seed=9)
gpp_b200.fit_road_planes(*big, planes[:64], mode='verified', return_index=True)
print('sanitize_small: all paths ran, exact/verified/decode/filter == oracle')
