"""GPU: the polling kernel (csrc/gpp_poll3.cuh, all four modes) under every schedule it can take -- detections cut into
plane segments, the pair database resident in shared memory / streamed from L2 / half and half, rows that repeat the
previous row written by the warp that polled the first of them -- must return what the oracle returns, bit for bit
('verified', 'exact', 'f64' against the FP64 oracle) or up to rounding-noise ties ('fast'); the fused pose / KITTI
epilogue must equal the stand-alone kernels bit for bit."""
import numpy as np
import pytest

from conftest import golden_cases, load_golden, load_planes
from gpp_b200.utils import synthetic
from oracle import c_oracle

pytestmark = pytest.mark.gpu


def _same(got, want):
    for g, w in zip(got, want):
        assert g.shape == w.shape and g.dtype == w.dtype
        assert np.array_equal(g, w, equal_nan=True)


def _hard_batch(planes, seed=77):
    """padding rows, runs of identical rows inside an image, detections without six votes, noisy key-points"""
    boxes, dims, orient, P_inv = synthetic.synth_detections(5, 48, planes, seed=seed, n_valid=37, kp_noise_px=3.0)
    boxes, dims, orient = boxes.copy(), dims.copy(), orient.copy()
    dims[1, :9, 1] *= 1.7                                   # max-votes < 6
    for b, (lo, hi) in ((2, (5, 11)), (3, (0, 4)), (4, (20, 21))):        # identical rows in the middle / at the start
        boxes[b, lo:hi], dims[b, lo:hi], orient[b, lo:hi] = boxes[b, lo], dims[b, lo], orient[b, lo]
    boxes[0, 47], dims[0, 47], orient[0, 47] = boxes[1, 0], dims[1, 0], orient[1, 0]   # equal rows in DIFFERENT images
    return boxes, dims, orient, P_inv


@pytest.mark.parametrize('n_seg,resident', [(1, 0), (1, 3), (1, -1), (2, -1), (5, 0), (7, 11), (32, -1), (32, 0)])
def test_every_schedule_equals_the_oracle(gpp, poller, n_seg, resident):
    planes = load_planes('10k')[:7001]                     # ragged last row, duplicate planes of the 10k database
    boxes, dims, orient, P_inv = _hard_batch(planes)
    want = c_oracle.fit_road_planes_c(boxes, dims, orient, P_inv, planes, return_index=True)
    want64 = c_oracle.fit_road_planes_c(boxes, dims, orient, P_inv, planes, return_index=True, dtype=np.float64)
    poller.debug_set_schedule(n_seg, resident)
    try:
        got = gpp.fit_road_planes(boxes, dims, orient, P_inv, planes, mode='verified', return_index=True)
        fast = gpp.fit_road_planes(boxes, dims, orient, P_inv, planes, mode='fast', return_index=True)
        exact = gpp.fit_road_planes(boxes, dims, orient, P_inv, planes, mode='exact', return_index=True)
        f64 = gpp.fit_road_planes(boxes, dims, orient, P_inv, planes, mode='f64', return_index=True)
        few = gpp.fit_road_planes(boxes[:1, :3], dims[:1, :3], orient[:1, :3], P_inv[:1], planes, mode='verified',
                                  return_index=True)
    finally:
        poller.debug_set_schedule(0, -1)
    _same(got, want)
    _same(exact, want)
    assert np.array_equal(f64[3], want64[3])
    assert np.allclose(f64[0], want64[0], rtol=1e-12, atol=0, equal_nan=True)
    assert np.array_equal(f64[2], want64[2], equal_nan=True)
    _same(few, c_oracle.fit_road_planes_c(boxes[:1, :3], dims[:1, :3], orient[:1, :3], P_inv[:1], planes,
                                          return_index=True))
    valid = orient >= 0                      # padding rows are degenerate: pure rounding-noise ties
    assert np.mean(fast[3][valid] == want[3][valid]) > 0.97


@pytest.mark.parametrize('name', golden_cases())
@pytest.mark.parametrize('n_seg', [3, 32])
def test_segmented_schedule_on_the_golden_vectors(gpp, poller, name, n_seg):
    """sentinel-100 winners, NaN planes, all-masked rows, per-image databases: the merge of the segments' partial
    results must keep every selection rule (max votes over ALL planes, lowest index on ties)"""
    g = load_golden(name)
    poller.debug_set_schedule(n_seg, 1)
    try:
        got = gpp.fit_road_planes(g['boxes'], g['dimensions'], g['orientations'], g['P_inv'], g['planes_raw'],
                                  mode='verified')
    finally:
        poller.debug_set_schedule(0, -1)
    _same(got, [g['keypoints'], g['keyplanes'], g['residuals']])


def test_repeated_calls_leave_the_counters_clean(gpp, poller):
    """the kernel resets its own claim / segment counters: many back-to-back calls of different shapes on one handle"""
    planes = load_planes('1k')
    rng = np.random.default_rng(5)
    for it in range(12):
        B, D = int(rng.integers(1, 5)), int(rng.integers(1, 60))
        boxes, dims, orient, P_inv = synthetic.synth_detections(B, D, planes, seed=100 + it, n_valid=max(1, D - 3))
        want = c_oracle.fit_road_planes_c(boxes, dims, orient, P_inv, planes, return_index=True)
        poller.debug_set_schedule(int(rng.integers(0, 9)), int(rng.integers(-1, 9)))
        try:
            got = gpp.fit_road_planes(boxes, dims, orient, P_inv, planes, mode='verified', return_index=True)
        finally:
            poller.debug_set_schedule(0, -1)
        _same(got, want)


def test_mid_size_batch_segment_major_order_and_shared_constants(gpp, poller):
    """Between one and four detections per resident warp the segments of a detection come in segment-major order, later
    segments adopt the first one's bound and copy its constants from the detection's scratch block: 96 padded images x
    100 rows x 10k planes (and the same call twice: the scratch is left clean) against the oracle, verified and exact;
    the fused pose epilogue on the same schedule against the plain one."""
    planes = load_planes('10k')
    boxes, dims, orient, P_inv = synthetic.synth_detections(96, 100, planes, seed=77, n_valid=61, kp_noise_px=2.5)
    want = c_oracle.fit_road_planes_c(boxes, dims, orient, P_inv, planes, return_index=True)
    for mode in ('verified', 'verified', 'exact'):
        got = gpp.fit_road_planes(boxes, dims, orient, P_inv, planes, mode=mode, return_index=True)
        for g, w in zip(got, want):
            assert np.array_equal(g, w, equal_nan=True), mode
    posed = gpp.fit_road_planes(boxes, dims, orient, P_inv, planes, return_index=True, return_pose=True)
    for g, w in zip(posed[:4], want):
        assert np.array_equal(g, w, equal_nan=True)
    # (the FAST mode is not compared across schedules: a plane is evaluated with one reciprocal for two rays once the scan
    # knows that max-votes is 6 and with separate ones before, so near-ties may fall differently -- its own rule is tested
    # in test_parity_gpu.py)
    fast = gpp.fit_road_planes(boxes, dims, orient, P_inv, planes, mode='fast', return_index=True)
    assert float(np.mean(fast[3][:, :61] == want[3][:, :61])) > 0.97          # (padding rows: every plane ties within rounding noise)


def test_large_batch_automatic_schedule_equals_exact(gpp):
    """automatic schedule of a large batch (one segment, 212 resident rows + streamed rest) against the EXACT mode on
    300 images x 100 rows x 21634 planes"""
    import torch
    planes = load_planes('22k')
    boxes, dims, orient, P_inv = synthetic.synth_detections(300, 100, planes, seed=9, n_valid=93)
    dev = torch.device('cuda', 0)
    args = [torch.from_numpy(a).to(dev) for a in (boxes, dims, orient, P_inv.astype(np.float32))]
    new = gpp.fit_road_planes_torch(*args, planes, mode='verified', return_index=True)
    ex = gpp.fit_road_planes_torch(*args, planes, mode='exact', return_index=True)
    torch.cuda.synchronize()
    for a, c in zip(new, ex):
        assert np.array_equal(a.cpu().numpy(), c.cpu().numpy(), equal_nan=True)


@pytest.mark.parametrize('mode', ['verified', 'exact', 'fast'])
@pytest.mark.parametrize('n_seg', [0, 4])
def test_fused_pose_epilogue_equals_the_stand_alone_kernels(gpp, poller, mode, n_seg):
    """return_pose / return_kitti (pose recovery and the KITTI record in the polling kernel's epilogue, one launch)
    against recover_pose / kitti_records on the polled key-points: the same device functions, the same bits -- incl.
    padding rows (orientation -1: zeros, input dimensions) and rows that are copies of their predecessor"""
    import torch
    planes = load_planes('1k')
    boxes, dims, orient, P_inv = _hard_batch(planes, seed=5)
    poller.set_planes(planes)                                         # (the database upload has launches of its own,
    poller.audit_set(0)                                               # and so has the audit pass if GPP_AUDIT is set)
    poller.debug_set_schedule(n_seg, -1)
    try:
        launches = poller.launch_count()
        full = gpp.fit_road_planes(boxes, dims, orient, P_inv, planes, mode=mode, return_index=True, return_pose=True,
                                   return_kitti=True)
        assert poller.launch_count() - launches == 1                  # one kernel for polling + pose + KITTI record
        dev = torch.device('cuda', 0)
        t = gpp.fit_road_planes_torch(torch.from_numpy(boxes).to(dev), torch.from_numpy(dims).to(dev),
                                      torch.from_numpy(orient).to(dev), torch.from_numpy(P_inv.astype(np.float32)).to(dev),
                                      planes, mode=mode, return_pose=True, return_kitti=True)
    finally:
        poller.debug_set_schedule(0, -1)
    base = gpp.fit_road_planes(boxes, dims, orient, P_inv, planes, mode=mode, return_index=True)
    assert len(full) == 8 and len(t) == 7
    _same(full[:4], base)
    loc, ang, dout = gpp.recover_pose(base[0].reshape(-1, 12), dims.reshape(-1, 3), orient.reshape(-1))
    rec = gpp.kitti_records(loc, ang, dout)
    shape = orient.shape
    for got, want in zip(full[4:], (loc, ang, dout, rec)):
        assert got.shape == shape + (want.shape[1],) and got.dtype == np.float32
        assert np.array_equal(got.reshape(want.shape), want, equal_nan=True)
    for a, b in zip(t[3:], full[4:]):
        assert np.array_equal(a.cpu().numpy(), b, equal_nan=True)
    assert np.all(full[4][orient < 0] == 0) and np.all(full[5][orient < 0] == 0)
    assert np.array_equal(full[6][orient < 0], dims[orient < 0])


def test_runtime_audit_counts(gpp, poller):
    """gpp_audit_set: every n-th detection of a 'verified' call is re-polled by the EXACT kernel on the device"""
    planes = load_planes('10k')
    boxes, dims, orient, P_inv = synthetic.synth_detections(7, 100, planes, seed=21, n_valid=90)
    c0, b0 = poller.audit_counts()
    poller.audit_set(3)
    try:
        got = gpp.fit_road_planes(boxes, dims, orient, P_inv, planes, mode='verified', return_index=True)
        again = gpp.fit_road_planes(boxes, dims, orient, P_inv, planes, mode='verified')       # without an index array
    finally:
        poller.audit_set(0)
    c1, b1 = poller.audit_counts()
    assert c1 - c0 == 2 * ((700 + 2) // 3) and b1 == b0
    want = c_oracle.fit_road_planes_c(boxes, dims, orient, P_inv, planes, return_index=True)
    _same(got, want)
    _same(again, want[:3])
    gpp.fit_road_planes(boxes, dims, orient, P_inv, planes, mode='verified')                  # audit off again
    assert poller.audit_counts() == (c1, b1)


def test_identical_rays_take_the_cheap_exact_form_bit_for_bit(gpp, poller):
    """Rows whose l / m / r rays are bitwise identical (FilterDetections' -1 padding; here also whole images, through a
    P_inv that maps every pixel to one ray) are polled in the exact arithmetic directly, in a form that computes the one
    point on the plane once.  The database holds planes parallel to that ray (point at infinity -> NaN differences: the
    general form must take over) and a b = 0 plane (NaN after normalisation)."""
    base = load_planes('1k')[:300].astype(np.float32)
    odd = np.array([[0.3, -0.5, 0.25, 1.3], [-0.2, 0.5, -0.25, -1.1], [1.0, 0.0, 0.0, 2.0], [0.0, -1.0, 0.0, 1.65]], np.float32)
    planes = np.concatenate([base[:100], odd, base[100:], odd[:2]], axis=0)
    boxes, dims, orient, P_inv = synthetic.synth_detections(3, 40, base, seed=515, n_valid=22)
    P_inv = P_inv.astype(np.float32).copy()
    P_inv[1] = np.array([[0, 0, 0], [0, 0, 0.5], [0, 0, 1], [0, 0, 0]], np.float32)      # every pixel -> ray (0, 0.5, 1)
    P_inv[2] = np.array([[0, 0, 0.1], [0, 0, 0.5], [0, 0, -1], [0, 0, 0]], np.float32)   # sign flip of the ray (z < 0)
    want = c_oracle.fit_road_planes_c(boxes, dims, orient, P_inv, planes, return_index=True)
    for n_seg in (0, 1, 5):
        poller.debug_set_schedule(n_seg, -1)
        try:
            got = gpp.fit_road_planes(boxes, dims, orient, P_inv, planes, mode='verified', return_index=True)
        finally:
            poller.debug_set_schedule(0, -1)
        _same(got, want)
    _same(gpp.fit_road_planes(boxes, dims, orient, P_inv, planes, mode='exact', return_index=True), want)
