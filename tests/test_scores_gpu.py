"""GPU: per-hypothesis arithmetic of the search loops (gpp_debug_scores) against the oracle.

This is the test that pins the EXACT mode's arithmetic itself, not just its arg-min: every (detection, plane)
vote count, residual sum and z-check of the device function must equal the oracle's bit for bit -- an FMA
contraction anywhere (ptxas contracts packed mul.rn.f32x2 + add.rn.f32x2 even with explicit .rn, which is why
EXACT runs the scalar kernel) shows up here immediately.  For the FAST path the deviation is measured and
bounded."""
import numpy as np
import pytest

from conftest import load_golden, load_planes
from gpp_b200.utils import synthetic
from oracle import fit_road_planes_ref as R

pytestmark = pytest.mark.gpu


def _oracle_scores(boxes, dims, orient, P_inv, planes, b, d):
    f = np.float32
    bx, dm, pi, pl = R._feed(boxes, dims, P_inv, planes, f)
    npl = R.normalise_planes(pl[0], f)
    rays = R.detection_rays(bx, pi, f)
    td = R.detection_dims(dm, orient, f)
    X, votes, resid, zc = R.hypotheses(rays[b, d:d + 1], td[b, d:d + 1], npl, f)
    return votes[0].astype(np.int32), resid[0], (zc[0] < 0)


@pytest.mark.parametrize('tag,seed', [('1k', 2), ('10k', 3), ('22k', 4)])
def test_exact_scores_are_bit_identical_per_hypothesis(poller, tag, seed):
    planes = load_planes(tag)
    boxes, dims, orient, P_inv = synthetic.synth_detections(2, 12, planes, seed=seed, n_valid=10)
    poller.set_planes(planes)
    for b in range(2):
        for d in range(12):
            votes, resid, zneg = poller.debug_scores(boxes[b, d], dims[b, d], orient[b, d], P_inv[b], which=0)
            wv, wr, wz = _oracle_scores(boxes, dims, orient, P_inv, planes, b, d)
            assert np.array_equal(votes, wv)
            assert np.array_equal(zneg, wz)
            assert np.array_equal(resid, wr, equal_nan=True)


def test_exact_scores_on_the_edge_case(poller):
    g = load_golden('edge_2x12x24')
    for b in range(2):
        poller.set_planes(g['planes'][b])
        for d in range(12):
            votes, resid, zneg = poller.debug_scores(g['boxes'][b, d], g['dimensions'][b, d], g['orientations'][b, d],
                                                     g['P_inv'][b], which=0)
            wv, wr, wz = _oracle_scores(g['boxes'], g['dimensions'], g['orientations'], g['P_inv'],
                                        g['planes'][b], b, d)
            assert np.array_equal(votes, wv) and np.array_equal(zneg, wz)
            assert np.array_equal(resid, wr, equal_nan=True)


@pytest.mark.parametrize('which', [1, 2])
def test_fast_scores_deviation_is_small(poller, which):
    """FAST arithmetic (FMA, MUFU, algebraic shortcuts): residual sums within 2e-3 of the exact fp32 ones on
    every hypothesis that could matter (finite, residual < 50), votes/z-check differ on < 1e-3 of them."""
    planes = load_planes('22k')
    boxes, dims, orient, P_inv = synthetic.synth_detections(2, 16, planes, seed=7)
    poller.set_planes(planes)
    worst, flips, total = 0.0, 0, 0
    for b in range(2):
        for d in range(16):
            votes, resid, zneg = poller.debug_scores(boxes[b, d], dims[b, d], orient[b, d], P_inv[b], which=which)
            wv, wr, wz = _oracle_scores(boxes, dims, orient, P_inv, planes, b, d)
            ok = np.isfinite(wr) & (wr < 50) & np.isfinite(resid)
            worst = max(worst, float(np.abs(resid[ok] - wr[ok]).max()))
            flips += int((votes[ok] != wv[ok]).sum() + (zneg[ok] != wz[ok]).sum())
            total += int(ok.sum())
    assert worst < 2e-3, worst
    assert flips < 1e-3 * total, (flips, total)
