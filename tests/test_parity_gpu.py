"""GPU: the CUDA path, called through the C ABI (ctypes -> libgpp.so), against the oracle.

Bars (BASELINE.json north_star):
  * EXACT mode: bit-identical to the oracle -- index, key-points, key-planes, residuals (== on the bits,
    NaN == NaN);
  * FAST mode: the selected plane equals the oracle's wherever the oracle's best and second-best masked
    residuals differ by more than NEAR_TIE_REL relative in FP64; values of agreeing rows within 1e-4 rel;
  * F64 mode: equal index, values within 1e-12 rel of the FP64 oracle.
"""
import numpy as np
import pytest

from conftest import golden_cases, load_golden, load_planes
from gpp_b200.utils import synthetic
from oracle import c_oracle

pytestmark = pytest.mark.gpu

NEAR_TIE_REL = 1e-6      # north_star: indices must match when best/second-best differ by more than this
VALUE_RTOL = 1e-4        # north_star: location / dimensions / yaw tolerance


def _assert_identical(got, want, names=('keypoints', 'keyplanes', 'residuals', 'index')):
    for g, w, n in zip(got, want, names):
        assert g.shape == w.shape, (n, g.shape, w.shape)
        assert g.dtype == w.dtype, (n, g.dtype, w.dtype)
        assert np.array_equal(g, w, equal_nan=True), '%s differs in %d entries' % (
            n, int((~((g == w) | (np.isnan(g.astype(np.float64)) & np.isnan(w.astype(np.float64))))).sum()))


@pytest.mark.parametrize('name', golden_cases())
def test_exact_mode_equals_golden_vectors(gpp, name):
    g = load_golden(name)
    planes = g['planes_raw']
    if planes.ndim == 2:
        planes = planes[None]                  # reference callers feed (1, N, 4), run_network.py:105
    got = gpp.fit_road_planes(g['boxes'], g['dimensions'], g['orientations'], g['P_inv'], planes, mode='exact')
    _assert_identical(got, [g['keypoints'], g['keyplanes'], g['residuals']])


@pytest.mark.parametrize('tag,B,D,seed', [('10', 1, 20, 1), ('1k', 1, 100, 2), ('100', 3, 37, 5),
                                         ('10k', 4, 100, 3), ('22k', 6, 100, 4), ('22k', 1, 1, 6)])
def test_exact_mode_is_bit_identical_to_the_oracle(gpp, tag, B, D, seed):
    planes = load_planes(tag)
    boxes, dims, orient, P_inv = synthetic.synth_detections(B, D, planes, seed=seed, n_valid=max(1, D - D // 5))
    got = gpp.fit_road_planes(boxes, dims, orient, P_inv, planes, mode='exact', return_index=True)
    want = c_oracle.fit_road_planes_c(boxes, dims, orient, P_inv, planes, return_index=True)
    _assert_identical(got, want)


def test_exact_mode_config3_full(gpp, poller):
    """Config 3 shape (64 x 100 x 10k), every row checked; automatic schedule and one segment per detection."""
    planes = load_planes('10k')
    boxes, dims, orient, P_inv = synthetic.synth_detections(64, 100, planes, seed=33)
    want = c_oracle.fit_road_planes_c(boxes, dims, orient, P_inv, planes, return_index=True)
    for n_seg in (0, 1):
        poller.debug_set_schedule(n_seg, -1)
        try:
            got = gpp.fit_road_planes(boxes, dims, orient, P_inv, planes, mode='exact', return_index=True)
        finally:
            poller.debug_set_schedule(0, -1)
        _assert_identical(got, want)


def test_max_votes_below_six_and_late_six(gpp):
    """Detections whose maximum vote count stays below 6 (general path for the whole database) and
    detections whose only 6-vote plane comes late (switch to the all-six-votes path mid-stream)."""
    planes = load_planes('1k')
    boxes, dims, orient, P_inv = synthetic.synth_detections(2, 100, planes, seed=34)
    dims = dims.copy()
    dims[0, :50, 1] *= 1.6                     # wrong widths: some polls can never vote -> max votes < 6
    dims[1, :50, 0] *= 0.5
    want = c_oracle.fit_road_planes_c(boxes, dims, orient, P_inv, planes, return_index=True)
    got = gpp.fit_road_planes(boxes, dims, orient, P_inv, planes, mode='exact', return_index=True)
    _assert_identical(got, want)
    # a database whose good planes are at the very end
    rev = planes[::-1].copy()
    want = c_oracle.fit_road_planes_c(boxes, dims, orient, P_inv, rev, return_index=True)
    got = gpp.fit_road_planes(boxes, dims, orient, P_inv, rev, mode='exact', return_index=True)
    _assert_identical(got, want)


def test_f64_mode_matches_fp64_oracle(gpp):
    planes = load_planes('22k')
    boxes, dims, orient, P_inv = synthetic.synth_detections(3, 100, planes, seed=12)
    got = gpp.fit_road_planes(boxes, dims, orient, P_inv, planes, mode='f64', return_index=True)
    want = c_oracle.fit_road_planes_c(boxes, dims, orient, P_inv, planes, return_index=True, dtype=np.float64)
    assert np.array_equal(got[3], want[3])
    for g, w in zip(got[:3], want[:3]):
        assert g.dtype == np.float64
        assert np.allclose(g, w, rtol=1e-12, atol=0, equal_nan=True)


def _check_fast_against_near_tie_rule(gpp, planes, boxes, dims, orient, P_inv):
    fast = gpp.fit_road_planes(boxes, dims, orient, P_inv, planes, mode='fast', return_index=True)
    want = c_oracle.fit_road_planes_c(boxes, dims, orient, P_inv, planes, return_index=True)
    mism = np.argwhere(fast[3] != want[3])
    # every mismatch must be a near-tie of the reference's own scores (FP64 verify mode on the GPU)
    if len(mism):
        f64 = gpp.fit_road_planes(boxes, dims, orient, P_inv, planes, mode='f64', return_index=True)
        from oracle.fit_road_planes_ref import second_best_gap
        for b, d in mism:
            sl = (slice(b, b + 1), slice(d, d + 1))
            best, v1, v2, getter = second_best_gap(boxes[sl], dims[sl], orient[sl], P_inv[b:b + 1], planes,
                                                   dtype=np.float32)
            r_ref = float(getter(0, 0, int(want[3][b, d])))
            r_fast = float(getter(0, 0, int(fast[3][b, d])))
            # the plane FAST picked scores within NEAR_TIE_REL of the reference's best in the reference's own
            # fp32 scores, or the fp64 verify mode sides with FAST's plane / calls it a tie at that level
            gap32 = abs(r_fast - r_ref) / max(abs(r_ref), 1e-30)
            best64, w1, w2, getter64 = second_best_gap(boxes[sl], dims[sl], orient[sl], P_inv[b:b + 1], planes,
                                                       dtype=np.float64)
            q_ref = float(getter64(0, 0, int(want[3][b, d])))
            q_fast = float(getter64(0, 0, int(fast[3][b, d])))
            gap64 = abs(q_fast - q_ref) / max(abs(q_ref), 1e-30)
            assert min(gap32, gap64) <= 64 * NEAR_TIE_REL or int(f64[3][b, d]) == int(fast[3][b, d]), (
                'fast mode picked plane %d, oracle %d, fp32 gap %.3e, fp64 gap %.3e' % (
                    fast[3][b, d], want[3][b, d], gap32, gap64))
    same = fast[3] == want[3]
    assert same.mean() > 0.97
    # rows that agree on the plane are recomputed in exact arithmetic: identical outputs
    assert np.array_equal(fast[0][same], want[0][same], equal_nan=True)
    assert np.array_equal(fast[1][same], want[1][same], equal_nan=True)
    fin = same & np.isfinite(want[2])
    assert np.allclose(fast[2][fin], want[2][fin], rtol=VALUE_RTOL, atol=0)
    return same.mean()


@pytest.mark.parametrize('tag,B,seed', [('1k', 4, 21), ('10k', 8, 22), ('22k', 8, 23)])
def test_fast_mode_obeys_the_near_tie_rule(gpp, tag, B, seed):
    planes = load_planes(tag)
    boxes, dims, orient, P_inv = synthetic.synth_detections(B, 100, planes, seed=seed)
    _check_fast_against_near_tie_rule(gpp, planes, boxes, dims, orient, P_inv)


def test_callers_dtypes_and_layouts(gpp):
    """Reference callers pass float64 P_inv / planes, Fortran-ordered planes from loadmat
    (run_network.py:75,105) and np.tile'd (B, N, 4) planes (preprocessing/kitti.py:220)."""
    planes = load_planes('100')
    boxes, dims, orient, P_inv = synthetic.synth_detections(3, 50, planes, seed=31)
    base = gpp.fit_road_planes(boxes, dims, orient, P_inv, planes, mode='exact')
    variants = [np.asfortranarray(planes), planes[None], np.tile(planes[None], (3, 1, 1)),
                np.broadcast_to(planes, (3,) + planes.shape), planes.astype(np.float32)]
    for pv in variants:
        out = gpp.fit_road_planes(boxes.astype(np.float64), dims.astype(np.float64), orient.astype(np.int64),
                                  np.asfortranarray(P_inv), pv, mode='exact')
        _assert_identical(out, base)
    layer = gpp.FitRoadPlanes(mode='exact')
    _assert_identical(layer.call([boxes, dims, orient, P_inv, planes[None]]), base)


def test_per_image_plane_databases(gpp):
    planes = load_planes('1k')
    boxes, dims, orient, P_inv = synthetic.synth_detections(3, 20, planes, seed=41)
    per_image = np.stack([planes[:300], planes[300:600], planes[:300]], axis=0)
    got = gpp.fit_road_planes(boxes, dims, orient, P_inv, per_image, mode='exact', return_index=True)
    want = c_oracle.fit_road_planes_c(boxes, dims, orient, P_inv, per_image, return_index=True)
    _assert_identical(got, want)


def test_empty_and_ragged_inputs(gpp):
    planes = load_planes('10')
    boxes, dims, orient, P_inv = synthetic.synth_detections(2, 7, planes, seed=51)
    got = gpp.fit_road_planes(boxes[:0], dims[:0], orient[:0], P_inv[:0], planes)
    assert got[0].shape == (0, 7, 4, 3) and got[1].shape == (0, 7, 1, 4) and got[2].shape == (0, 7)
    got = gpp.fit_road_planes(boxes[:, :0], dims[:, :0], orient[:, :0], P_inv, planes)
    assert got[0].shape == (2, 0, 4, 3)
    # N not a multiple of the warp width / tile, D = 1
    for n in (1, 31, 33, 1023, 1025):
        db = load_planes('10k')[:n]
        got = gpp.fit_road_planes(boxes[:, :1], dims[:, :1], orient[:, :1], P_inv, db, mode='exact', return_index=True)
        want = c_oracle.fit_road_planes_c(boxes[:, :1], dims[:, :1], orient[:, :1], P_inv, db, return_index=True)
        _assert_identical(got, want)
    with pytest.raises(ValueError):
        gpp.fit_road_planes(boxes, dims[:, :3], orient, P_inv, planes)
    with pytest.raises(ValueError):
        gpp.fit_road_planes(boxes, dims, orient, P_inv, planes[:, :3])
    with pytest.raises(ValueError):
        gpp.fit_road_planes(boxes, dims, orient, P_inv, planes, mode='bogus')


def test_device_entry_and_dlpack(gpp):
    import torch
    planes = load_planes('1k')
    boxes, dims, orient, P_inv = synthetic.synth_detections(5, 100, planes, seed=61)
    want = c_oracle.fit_road_planes_c(boxes, dims, orient, P_inv, planes, return_index=True)
    dev = torch.device('cuda', 0)
    tb, td = torch.from_numpy(boxes).to(dev), torch.from_numpy(dims).to(dev)
    to, tp = torch.from_numpy(orient).to(dev), torch.from_numpy(P_inv).to(dev)
    tpl = torch.from_numpy(planes).to(dev)
    side = torch.cuda.Stream(device=dev)
    side.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(side):
        out = gpp.fit_road_planes_torch(tb, td, to, tp, tpl, mode='exact', return_index=True)
    side.synchronize()
    _assert_identical([o.cpu().numpy() for o in out], want)
    out2 = gpp.fit_road_planes_dlpack(tb, td, to, tp, tpl, mode='exact', return_index=True)
    torch.cuda.synchronize()
    _assert_identical([o.cpu().numpy() for o in out2], want)


def test_torch_database_is_never_stale(gpp, poller):
    """A fresh CUDA tensor per call with different content must be polled against ITS content, also when the caching
    allocator hands it the block (data_ptr, shape, version 0) of the tensor before it; an in-place update of the same
    tensor must be seen; a database updated on one stream must be seen by a fit on another."""
    import torch
    dev = torch.device('cuda', 0)
    base = load_planes('1k').astype(np.float32)
    dbs = [np.ascontiguousarray(base[k * 200:(k + 1) * 200]) for k in range(4)]
    boxes, dims, orient, P_inv = synthetic.synth_detections(2, 30, base, seed=73)
    t = [torch.from_numpy(a).to(dev) for a in (boxes, dims, orient, P_inv.astype(np.float32))]
    ptrs = set()
    for db in dbs + dbs[::-1]:
        planes = torch.from_numpy(db).to(dev)                    # same size every time: the allocator recycles blocks
        ptrs.add(planes.data_ptr())
        out = gpp.fit_road_planes_torch(*t, planes, mode='exact', return_index=True)
        want = c_oracle.fit_road_planes_c(boxes, dims, orient, P_inv, db, return_index=True)
        _assert_identical([o.cpu().numpy() for o in out], want)
        del planes, out
    assert len(ptrs) < 8                                         # the scenario did occur: addresses were reused
    planes = torch.from_numpy(dbs[0]).to(dev)
    gpp.fit_road_planes_torch(*t, planes, mode='exact')
    planes.copy_(torch.from_numpy(dbs[1]).to(dev))               # in place: same object, new version
    out = gpp.fit_road_planes_torch(*t, planes, mode='exact', return_index=True)
    _assert_identical([o.cpu().numpy() for o in out], c_oracle.fit_road_planes_c(boxes, dims, orient, P_inv, dbs[1], return_index=True))
    side = torch.cuda.Stream(device=dev)
    planes2 = torch.from_numpy(dbs[2]).to(dev)
    torch.cuda.synchronize()
    with torch.cuda.stream(side):
        poller.set_planes_torch(planes2)                         # database updated on the side stream ...
    out = poller.fit_torch(*t, mode='exact', return_index=True)  # ... polled on the current one
    torch.cuda.synchronize()
    _assert_identical([o.cpu().numpy() for o in out], c_oracle.fit_road_planes_c(boxes, dims, orient, P_inv, dbs[2], return_index=True))
    # one database per call: an expand()ed view is one, a materialised tile is refused without touching the device
    tiled = planes2.unsqueeze(0).expand(2, -1, -1)
    out = gpp.fit_road_planes_torch(*t, tiled, mode='exact', return_index=True)
    _assert_identical([o.cpu().numpy() for o in out], c_oracle.fit_road_planes_c(boxes, dims, orient, P_inv, dbs[2], return_index=True))
    with pytest.raises(ValueError):
        gpp.fit_road_planes_torch(*t, tiled.contiguous(), mode='exact')


def test_database_fed_with_the_call_is_compared_while_the_gpu_polls(gpp, poller):
    """gpp_fit_planes_host: a small call polls against the resident database while the host compares the bytes it was
    handed; same bytes -> one launch; other bytes of the same shape -> upload and a second poll, results of the NEW
    database (also for the pose / KITTI outputs)."""
    base = load_planes('10k')
    db1, db2 = np.asfortranarray(base[:3000]), np.asfortranarray(base[3000:6000])           # as loadmat returns them
    boxes, dims, orient, P_inv = synthetic.synth_detections(1, 100, base, seed=97, n_valid=17)
    want1 = c_oracle.fit_road_planes_c(boxes, dims, orient, P_inv, db1, return_index=True)
    want2 = c_oracle.fit_road_planes_c(boxes, dims, orient, P_inv, db2, return_index=True)
    _assert_identical(gpp.fit_road_planes(boxes, dims, orient, P_inv, db1[None], return_index=True), want1)
    n0 = poller.launch_count()
    _assert_identical(gpp.fit_road_planes(boxes, dims, orient, P_inv, db1.copy(order='F')[None], return_index=True), want1)
    assert poller.launch_count() - n0 == 1                        # the resident database was the right one
    n0 = poller.launch_count()
    got = gpp.fit_road_planes(boxes, dims, orient, P_inv, db2[None], return_index=True, return_pose=True, return_kitti=True)
    _assert_identical(got[:4], want2)
    assert poller.launch_count() - n0 == 2 + 4                    # polled twice around the upload (4 launches)
    again = gpp.fit_road_planes(boxes, dims, orient, P_inv, db2[None], return_index=True, return_pose=True, return_kitti=True)
    for a, b in zip(got, again):
        assert np.array_equal(a, b, equal_nan=True)
    _assert_identical(gpp.fit_road_planes(boxes, dims, orient, P_inv, db1[None], mode='exact', return_index=True), want1)


def test_large_device_fed_database_gets_the_scan_order(gpp, poller):
    """A database of 32 rows or more that arrives as a CUDA tensor is read back once to derive the order it is scanned
    in (csrc/gpp_order.cu); the results are those of the host-fed database in every mode, on any stream."""
    import torch
    dev = torch.device('cuda', 0)
    planes = load_planes('10k')
    boxes, dims, orient, P_inv = synthetic.synth_detections(3, 100, planes, seed=91, n_valid=70)
    t = [torch.from_numpy(a).to(dev) for a in (boxes, dims, orient, P_inv.astype(np.float32))]
    want = c_oracle.fit_road_planes_c(boxes, dims, orient, P_inv, planes, return_index=True)
    host_fed = {m: gpp.fit_road_planes(boxes, dims, orient, P_inv, planes, mode=m, return_index=True) for m in ('fast', 'f64')}
    poller.set_planes(load_planes('100'))                          # something else resident in between
    tpl = torch.from_numpy(planes.astype(np.float32)).to(dev)
    side = torch.cuda.Stream(device=dev)
    side.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(side):
        got = {m: gpp.fit_road_planes_torch(*t, tpl, mode=m, return_index=True) for m in ('verified', 'exact', 'fast')}
    side.synchronize()
    for m in ('verified', 'exact'):
        _assert_identical([o.cpu().numpy() for o in got[m]], want)
    _assert_identical([o.cpu().numpy() for o in got['fast']], host_fed['fast'])


def test_tiled_numpy_databases_are_grouped_exactly(gpp):
    """(B, N, 4) databases as preprocessing/kitti.py:220 tiles them: one upload for a true tile, per-image groups when
    one image's copy differs in a single value"""
    from gpp_b200.layers.fit_road_planes import _plane_groups
    db = load_planes('100')
    tile = np.tile(db[None], (6, 1, 1))
    assert [(a, b) for a, b, _ in _plane_groups(tile, 6)] == [(0, 6)]
    assert [(a, b) for a, b, _ in _plane_groups(np.broadcast_to(db, (6,) + db.shape), 6)] == [(0, 6)]
    tile[3, 57, 2] = np.nextafter(tile[3, 57, 2], 1.0)
    assert [(a, b) for a, b, _ in _plane_groups(tile, 6)] == [(0, 3), (3, 4), (4, 6)]
    assert [(a, b) for a, b, _ in _plane_groups(np.asfortranarray(tile), 6)] == [(0, 3), (3, 4), (4, 6)]
    boxes, dims, orient, P_inv = synthetic.synth_detections(6, 20, db, seed=75)
    got = gpp.fit_road_planes(boxes, dims, orient, P_inv, tile, mode='exact', return_index=True)
    want = [c_oracle.fit_road_planes_c(boxes[b:b + 1], dims[b:b + 1], orient[b:b + 1], P_inv[b:b + 1], tile[b], return_index=True)
            for b in range(6)]
    _assert_identical(got, [np.concatenate([w[i] for w in want], axis=0) for i in range(4)])


def test_database_reupload_is_skipped_and_switching_works(gpp, poller):
    p1, p2 = load_planes('100'), load_planes('1k')
    boxes, dims, orient, P_inv = synthetic.synth_detections(1, 10, p1, seed=71)
    a1 = gpp.fit_road_planes(boxes, dims, orient, P_inv, p1, mode='exact', return_index=True)
    n0 = poller.launch_count()
    a1b = gpp.fit_road_planes(boxes, dims, orient, P_inv, p1.copy(), mode='exact', return_index=True)
    same_db = poller.launch_count() - n0               # the polling kernel alone
    n0 = poller.launch_count()
    a2 = gpp.fit_road_planes(boxes, dims, orient, P_inv, p2, mode='exact', return_index=True)
    assert poller.launch_count() - n0 == same_db + 4   # a new database adds 2 normalisations + scan index + interleave
    a1c = gpp.fit_road_planes(boxes, dims, orient, P_inv, p1, mode='exact', return_index=True)
    _assert_identical(a1b, a1)
    _assert_identical(a1c, a1)
    _assert_identical(a2, c_oracle.fit_road_planes_c(boxes, dims, orient, P_inv, p2, return_index=True))
    assert np.allclose(poller.normalised_planes()[:, :3].astype(np.float64).__pow__(2).sum(1), 1.0, atol=1e-6)


def test_large_batch_properties_config4_slice(gpp):
    """Full-size planes (21634) with a few thousand detections: size-independent properties instead of a
    row-by-row oracle check -- (a) chunked host entry == one device launch, (b) permuting the detections
    permutes the outputs, (c) the returned key-plane is the normalised database row of the returned index,
    (d) a random subset matches the oracle bit for bit."""
    import torch
    planes = load_planes('22k')
    B, D = 720, 100                                   # > one host chunk (65536 detections)
    boxes, dims, orient, P_inv = synthetic.synth_detections(B, D, planes, seed=81)
    host = gpp.fit_road_planes(boxes, dims, orient, P_inv, planes, mode='exact', return_index=True)
    dev = torch.device('cuda', 0)
    out = gpp.fit_road_planes_torch(torch.from_numpy(boxes).to(dev), torch.from_numpy(dims).to(dev),
                                    torch.from_numpy(orient).to(dev), torch.from_numpy(P_inv).to(dev),
                                    planes, mode='exact', return_index=True)
    torch.cuda.synchronize()
    _assert_identical([o.cpu().numpy() for o in out], host)
    perm = np.random.default_rng(0).permutation(B)
    permuted = gpp.fit_road_planes(boxes[perm], dims[perm], orient[perm], P_inv[perm], planes, mode='exact',
                                   return_index=True)
    _assert_identical(permuted, [h[perm] for h in host])
    npl = gpp.get_poller(0).normalised_planes()
    assert np.array_equal(host[1][:, :, 0, :], npl[host[3]])
    pick = np.random.default_rng(1).choice(B, size=6, replace=False)
    want = c_oracle.fit_road_planes_c(boxes[pick], dims[pick], orient[pick], P_inv[pick], planes, return_index=True)
    _assert_identical([h[pick] for h in host], want)


def test_multi_device_driver_equals_single_device(gpp):
    """fit_road_planes_multi over every visible GPU == fit_road_planes on one (bit for bit); with one GPU this
    exercises the single-shard path, under `gpurun --gpus N` the threaded one."""
    planes = load_planes('1k')
    boxes, dims, orient, P_inv = synthetic.synth_detections(37, 100, planes, seed=91, n_valid=40)
    n = gpp._lib.load().gpp_device_count()
    assert n >= 1
    want = gpp.fit_road_planes(boxes, dims, orient, P_inv, planes, return_index=True)
    got = gpp.fit_road_planes_multi(boxes, dims, orient, P_inv, planes, return_index=True)
    assert all(np.array_equal(a, b, equal_nan=True) for a, b in zip(got, want))
    got = gpp.fit_road_planes_multi(boxes, dims, orient, P_inv, planes[None], devices=list(range(n))[::-1], mode='f64')
    want = gpp.fit_road_planes(boxes, dims, orient, P_inv, planes, mode='f64')
    assert got[0].dtype == np.float64 and all(np.array_equal(a, b, equal_nan=True) for a, b in zip(got, want))


def test_chunked_host_path_pageable_and_pinned_memory(gpp):
    """Calls with more than 65536 detections are pipelined in chunks; pageable caller memory goes through pinned
    staging blocks, pinned caller memory is copied directly.  Both equal the oracle, with and without the index,
    with a ragged last chunk, and in the FP64 mode."""
    import torch
    planes = load_planes('100')
    boxes, dims, orient, P_inv = synthetic.synth_detections(1501, 100, planes, seed=5, n_valid=70)
    P32 = P_inv.astype(np.float32)
    want = c_oracle.fit_road_planes_c(boxes, dims, orient, P32, planes, return_index=True)
    got = gpp.fit_road_planes(boxes, dims, orient, P32, planes, mode='exact', return_index=True)        # pageable
    assert all(np.array_equal(a, b, equal_nan=True) for a, b in zip(got, want))
    got = gpp.fit_road_planes(boxes, dims, orient, P32, planes, mode='verified')
    assert all(np.array_equal(a, b, equal_nan=True) for a, b in zip(got, want[:3]))

    def pinned(a):
        t = torch.empty(a.shape, dtype=torch.from_numpy(a).dtype, pin_memory=True)
        t.numpy()[...] = a
        return t
    keep = [pinned(a) for a in (boxes, dims, orient, P32)]
    outs = [torch.empty(w.shape, dtype=torch.from_numpy(w).dtype, pin_memory=True) for w in want]
    res = gpp.fit_road_planes(*[t.numpy() for t in keep], planes, mode='verified', return_index=True,
                              out=[t.numpy() for t in outs])
    assert all(np.array_equal(a, b, equal_nan=True) for a, b in zip(res, want))
    w64 = c_oracle.fit_road_planes_c(boxes[:700], dims[:700], orient[:700], P32[:700], planes, dtype=np.float64)
    g64 = gpp.fit_road_planes(boxes[:700], dims[:700], orient[:700], P32[:700], planes, mode='f64')
    assert all(np.array_equal(a, b, equal_nan=True) for a, b in zip(g64, w64))
