"""GPU: 'verified' must equal 'exact' on ANY float32 input, not only on the benchmark generator's geometry.  Inputs come
from gpp_b200.utils.adversarial: random / steep / un-normalised / degenerate plane databases against detections that are
far away, very close, collapsed to (almost) one pixel, on the vanishing line of a plane, seen through a random P_inv ...
  * 1.07 million detections (3.2e9 hypotheses): index, key-points, key-planes, residuals bit-identical;
  * per hypothesis: |fast residual sum - exact residual sum| stays below HALF the margin the filter relies on, the votes
    possible within the margin are never fewer than the exact votes, the z-check bound never excludes a passing plane
    (round 2 found and fixed a 9-fold violation on steep planes this way: DESIGN.md section 4.1.1)."""
import numpy as np
import pytest

from conftest import load_planes
from gpp_b200.utils import adversarial as adv

pytestmark = pytest.mark.gpu


def test_verified_equals_exact_on_a_million_adversarial_detections(gpp, poller):
    import torch
    dev = torch.device('cuda', 0)
    bases = {0: load_planes('10k'), 1: load_planes('22k')}
    n_det = n_bad = 0
    for seed in range(6):
        rng = np.random.default_rng(1000 + seed)
        base = bases[seed % 2]
        for pf in adv.PLANE_FLAVOURS:
            db = adv.planes(pf, 3000, rng, base=base)
            poller.set_planes(db)
            for df in adv.DET_FLAVOURS:
                boxes, dims, orient, P_inv = adv.detections(df, 64, 100, rng, base)
                t = [torch.from_numpy(a).to(dev) for a in (boxes, dims, orient, P_inv)]
                ve = poller.fit_torch(*t, mode='verified', return_index=True)
                ex = poller.fit_torch(*t, mode='exact', return_index=True)
                torch.cuda.synchronize()
                for a, b in zip(ve, ex):
                    a, b = a.cpu().numpy(), b.cpu().numpy()
                    bad = ~((a == b) | ((a != a) & (b != b)))
                    assert not bad.any(), 'verified != exact for planes %s / detections %s, seed %d: %d values' % (
                        pf, df, seed, int(bad.sum()))
                n_det += boxes.shape[0] * boxes.shape[1]
    assert n_det >= 1000000 and n_bad == 0


@pytest.mark.parametrize('plane_flavour', adv.PLANE_FLAVOURS)
def test_margin_holds_per_hypothesis_on_adversarial_inputs(gpp, poller, plane_flavour):
    rng = np.random.default_rng(77 + adv.PLANE_FLAVOURS.index(plane_flavour))
    base = load_planes('10k')
    db = adv.planes(plane_flavour, 4000, rng, base=base)
    poller.set_planes(db)
    worst, n_fin = 0.0, 0
    for df in adv.DET_FLAVOURS:
        boxes, dims, orient, P_inv = adv.detections(df, 2, 12, rng, base)
        for b in range(boxes.shape[0]):
            for d in range(boxes.shape[1]):
                ev, er, ez = poller.debug_scores(boxes[b, d], dims[b, d], orient[b, d], P_inv[b], which=0)
                for which in (1, 2):
                    fv, fr, fz, fm, vhi, zok = poller.debug_scores(boxes[b, d], dims[b, d], orient[b, d], P_inv[b],
                                                                   which=which, with_margin=True)
                    fin = np.isfinite(er) & np.isfinite(fr)
                    # a non-finite margin means "always re-evaluated exactly": nothing to bound there
                    bounded = fin & np.isfinite(fm)
                    if bounded.any():
                        ratio = np.abs(fr[bounded] - er[bounded]) / np.maximum(fm[bounded], np.float32(1e-30))
                        worst = max(worst, float(ratio.max()))
                        n_fin += int(bounded.sum())
                    if which == 1:
                        assert (vhi[fin] >= ev[fin]).all(), (plane_flavour, df)
                        assert zok[fin & ~ez].all(), (plane_flavour, df)
    assert n_fin > 500000
    assert worst <= 0.5, 'largest |fast - exact| / margin = %.3f for %s planes' % (worst, plane_flavour)
