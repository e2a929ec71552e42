"""GPU: the VERIFIED mode (FAST arithmetic as a filter, EXACT re-evaluation of everything that could win)
must return exactly what the EXACT mode returns -- index, key-points, key-planes, residuals, bit for bit.

Small cases are checked against the oracle; the benchmark-size batch (config 4 slice: 1024 images x 100
detections x 21634 planes = 2.2e9 hypotheses) is checked against the EXACT mode on the GPU, which itself is
pinned to the oracle per hypothesis (tests/test_scores_gpu.py).  The margin the filter relies on is checked
per hypothesis: |fast residual sum - exact residual sum| must stay below a quarter of the margin."""
import numpy as np
import pytest

from conftest import golden_cases, load_golden, load_planes
from gpp_b200.utils import synthetic
from oracle import c_oracle
from oracle import fit_road_planes_ref as R

pytestmark = pytest.mark.gpu


def _same(got, want):
    for g, w in zip(got, want):
        assert g.shape == w.shape and g.dtype == w.dtype
        assert np.array_equal(g, w, equal_nan=True)


@pytest.mark.parametrize('name', golden_cases())
def test_verified_equals_golden_vectors(gpp, name):
    g = load_golden(name)
    got = gpp.fit_road_planes(g['boxes'], g['dimensions'], g['orientations'], g['P_inv'], g['planes_raw'],
                              mode='verified')
    _same(got, [g['keypoints'], g['keyplanes'], g['residuals']])


@pytest.mark.parametrize('tag,B,D,seed,nv', [('10', 1, 20, 1, None), ('1k', 1, 100, 2, None), ('100', 3, 37, 5, 30),
                                            ('10k', 64, 100, 33, None), ('22k', 24, 100, 4, 80), ('22k', 1, 1, 6, None)])
def test_verified_equals_oracle(gpp, tag, B, D, seed, nv):
    planes = load_planes(tag)
    boxes, dims, orient, P_inv = synthetic.synth_detections(B, D, planes, seed=seed, n_valid=nv)
    got = gpp.fit_road_planes(boxes, dims, orient, P_inv, planes, mode='verified', return_index=True)
    want = c_oracle.fit_road_planes_c(boxes, dims, orient, P_inv, planes, return_index=True)
    _same(got, want)


def test_verified_on_hard_inputs(gpp):
    """max votes < 6 for whole detections, good planes only at the end of the database, swapped key-points
    (every plane fails the z-check), absurd dimensions, far-away objects (large margins)."""
    planes = load_planes('1k')
    boxes, dims, orient, P_inv = synthetic.synth_detections(3, 100, planes, seed=34)
    boxes, dims = boxes.copy(), dims.copy()
    dims[0, :50, 1] *= 1.6
    dims[1, :50, 0] *= 0.5
    boxes[2, :30, 4:6], boxes[2, :30, 8:10] = boxes[2, :30, 8:10].copy(), boxes[2, :30, 4:6].copy()
    dims[2, 30:40] = [40.0, 50.0, 60.0]
    for db in (planes, planes[::-1].copy(), load_planes('10k')[:3000]):
        want = c_oracle.fit_road_planes_c(boxes, dims, orient, P_inv, db, return_index=True)
        got = gpp.fit_road_planes(boxes, dims, orient, P_inv, db, mode='verified', return_index=True)
        _same(got, want)


def test_verified_equals_exact_on_a_config4_slice(gpp):
    import torch
    planes = load_planes('22k')
    B, D = 1024, 100
    boxes, dims, orient, P_inv = synthetic.synth_detections(B, D, planes, seed=101)
    dev = torch.device('cuda', 0)
    args = [torch.from_numpy(a).to(dev) for a in (boxes, dims, orient, P_inv.astype(np.float32))]
    ex = gpp.fit_road_planes_torch(*args, planes, mode='exact', return_index=True)
    ve = gpp.fit_road_planes_torch(*args, planes, mode='verified', return_index=True)
    torch.cuda.synchronize()
    for a, b in zip(ex, ve):
        assert torch.equal(a, b) or np.array_equal(a.cpu().numpy(), b.cpu().numpy(), equal_nan=True)
    # the same through the chunked host entry
    host = gpp.fit_road_planes(boxes[:700], dims[:700], orient[:700], P_inv[:700], planes, mode='verified',
                               return_index=True)
    _same(host, [t[:700].cpu().numpy() for t in ex])


def _oracle_bottom3(boxes, dims, orient, P_inv, planes, b, d):
    """EXACT-arithmetic sum of the three bottom-face residuals (|X_l-X_m|, |X_m-X_r|, |X_l-X_r| against their
    targets) of one detection against every plane, float32 like the oracle."""
    from oracle import fit_road_planes_ref as R
    f = np.float32
    bx, dm, pi, pl = R._feed(boxes, dims, P_inv, planes, f)
    npl = R.normalise_planes(pl[0] if pl.ndim == 3 else pl, f)
    rays = R.detection_rays(bx, pi, f)
    td = R.detection_dims(dm, orient, f)
    (Xl, Xm, Xr, Xt), votes, resid, zc = R.hypotheses(rays[b, d:d + 1], td[b, d:d + 1], npl, f)
    with np.errstate(all='ignore'):
        r1 = np.abs(R._dist(Xl, Xm) - td[b, d, 1])
        r2 = np.abs(R._dist(Xm, Xr) - td[b, d, 2])
        r3 = np.abs(R._dist(Xl, Xr) - td[b, d, 3])
        return ((r1 + r2) + r3)[0].astype(f)


def test_stage_one_margin_bounds_the_bottom_face_sum(poller):
    """Stage 1 of the all-six phase stops on S3 - (w ms + mc + 2^-20 S3) > best, S3 = fast sum of the three
    bottom-face residuals: that margin is at least 4x the observed |fast - exact| deviation of S3 wherever it could
    matter, and S3 (exact) never exceeds the exact residual sum."""
    planes = load_planes('22k')
    boxes, dims, orient, P_inv = synthetic.synth_detections(3, 16, planes, seed=212, kp_noise_px=3.0)
    poller.set_planes(planes)
    worst = 0.0
    for b in range(3):
        for d in range(16):
            out = poller.debug_scores(boxes[b, d], dims[b, d], orient[b, d], P_inv[b], which=3, with_margin=True)
            s3_fast, m1 = out[1], out[3]
            s3_exact = _oracle_bottom3(boxes, dims, orient, P_inv, planes, b, d)
            ev, er, ez = poller.debug_scores(boxes[b, d], dims[b, d], orient[b, d], P_inv[b], which=0)
            fin = np.isfinite(er) & np.isfinite(s3_exact)
            assert (s3_exact[fin] <= er[fin]).all()                # a lower bound of the residual sum
            rel = fin & (er < 6 * 0.7 + 1.0) & np.isfinite(s3_fast)
            assert np.isfinite(m1[rel]).all() and (m1[rel] > 0).all()
            ratio = np.abs(s3_fast[rel] - s3_exact[rel]) / m1[rel]
            worst = max(worst, float(ratio.max()) if ratio.size else 0.0)
    assert worst < 0.25, worst


def test_margin_bounds_the_fast_vs_exact_deviation(poller):
    """Per hypothesis, wherever it could matter (exact: six votes or close to it, finite): the filter's
    margin is at least 4x the observed |fast - exact| deviation of the residual sum, and the loosened vote /
    z-check tests are supersets of the exact ones."""
    planes = load_planes('22k')
    boxes, dims, orient, P_inv = synthetic.synth_detections(3, 24, planes, seed=202)
    poller.set_planes(planes)
    worst = 0.0
    for b in range(3):
        for d in range(24):
            for which in (1, 2):
                fv, fr, fz, fm, vhi, zok = poller.debug_scores(boxes[b, d], dims[b, d], orient[b, d], P_inv[b],
                                                               which=which, with_margin=True)
                ev, er, ez = poller.debug_scores(boxes[b, d], dims[b, d], orient[b, d], P_inv[b], which=0)
                rel = np.isfinite(er) & (er < 6 * 0.7 + 1.0)        # anything that can carry six votes
                assert np.isfinite(fm[rel]).all() and (fm[rel] > 0).all()
                ratio = np.abs(fr[rel] - er[rel]) / fm[rel]
                worst = max(worst, float(ratio.max()) if ratio.size else 0.0)
                # superset properties the filters rely on, on EVERY hypothesis
                assert (vhi >= ev).all()                            # votes possible within the margin >= exact votes
                assert zok[~ez].all()                               # exact z-check passes -> filter lets it pass
    assert worst < 0.25, worst


def test_padding_rows_are_polled_once_and_copied(gpp):
    """FilterDetections pads every image to D rows with -1: the repeats are computed once per image and
    copied, and still equal the oracle bit for bit in every mode."""
    planes = load_planes('22k')
    boxes, dims, orient, P_inv = synthetic.synth_detections(6, 100, planes, seed=303, n_valid=13)
    boxes[2] = -1.0; dims[2] = -1.0; orient[2] = -1            # an image without any detection
    boxes[3, 50] = boxes[3, 12]; dims[3, 50] = dims[3, 12]; orient[3, 50] = orient[3, 12]   # a valid row among padding
    want = c_oracle.fit_road_planes_c(boxes, dims, orient, P_inv, planes, return_index=True)
    for mode in ('verified', 'exact'):
        got = gpp.fit_road_planes(boxes, dims, orient, P_inv, planes, mode=mode, return_index=True)
        _same(got, want)
    got64 = gpp.fit_road_planes(boxes, dims, orient, P_inv, planes, mode='f64', return_index=True)
    want64 = c_oracle.fit_road_planes_c(boxes, dims, orient, P_inv, planes, return_index=True, dtype=np.float64)
    assert np.array_equal(got64[3], want64[3])
    assert np.allclose(got64[0], want64[0], rtol=1e-12, atol=0, equal_nan=True)


@pytest.mark.parametrize('ray_scale', [1.0, 1000.0, 1e-3])
def test_verified_is_robust_to_ray_scale_and_odd_geometry(gpp, poller, ray_scale):
    """The margin must not depend on how P_inv happens to be scaled (rays are only defined up to a factor), and
    the z-check bound must hold for any geometry: tiny / huge dimensions, far-away and very close objects."""
    planes = load_planes('10k')[:4000]
    boxes, dims, orient, P_inv = synthetic.synth_detections(2, 60, planes, seed=505)
    dims = dims.copy()
    dims[0, :10] *= 12.0                     # absurdly large boxes: key-point distances of tens of metres
    dims[0, 10:20] *= 0.05                   # absurdly small ones
    boxes = boxes.copy()
    boxes[1, :10, 4:] = boxes[1, :10, 4:] * 0.02 + 650.0      # key-points a few pixels apart: very distant geometry
    P_scaled = P_inv * ray_scale
    want = c_oracle.fit_road_planes_c(boxes, dims, orient, P_scaled, planes, return_index=True)
    got = gpp.fit_road_planes(boxes, dims, orient, P_scaled, planes, mode='verified', return_index=True)
    _same(got, want)
    poller.set_planes(planes)
    for b, d in ((0, 0), (0, 5), (0, 12), (1, 3), (1, 30), (0, 40)):
        for which in (1, 2):
            fv, fr, fz, fm, vhi, zok = poller.debug_scores(boxes[b, d], dims[b, d], orient[b, d], P_scaled[b],
                                                           which=which, with_margin=True)
            ev, er, ez = poller.debug_scores(boxes[b, d], dims[b, d], orient[b, d], P_scaled[b], which=0)
            assert (vhi >= ev).all()
            assert zok[~ez].all()
            fin = np.isfinite(er) & np.isfinite(fr)
            assert (np.abs(fr[fin] - er[fin]) <= fm[fin]).all()
        out = poller.debug_scores(boxes[b, d], dims[b, d], orient[b, d], P_scaled[b], which=3, with_margin=True)
        s3_exact = _oracle_bottom3(boxes, dims, orient, P_scaled, planes, b, d)
        fin = np.isfinite(s3_exact) & np.isfinite(out[1])
        assert (np.abs(out[1][fin] - s3_exact[fin]) <= out[3][fin]).all()         # stage-1 margin, odd geometry
