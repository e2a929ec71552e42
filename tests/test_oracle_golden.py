"""CPU: the oracle (numpy and C restatements) against the golden vectors produced by executing the
reference's own fit_road_planes.py over numpy op stand-ins (tests/golden/make_golden.py)."""
import numpy as np
import pytest

from conftest import golden_cases, load_golden
from oracle import c_oracle
from oracle.fit_road_planes_ref import fit_road_planes_ref


def _same(a, b):
    return a.shape == b.shape and a.dtype == b.dtype and np.array_equal(a, b, equal_nan=True)


@pytest.mark.parametrize('name', golden_cases())
def test_numpy_oracle_matches_reference_graph(name):
    g = load_golden(name)
    kp, kpl, res = fit_road_planes_ref(g['boxes'], g['dimensions'], g['orientations'], g['P_inv'], g['planes_raw'])
    assert _same(kp, g['keypoints'])
    assert _same(kpl, g['keyplanes'])
    assert _same(res, g['residuals'])


@pytest.mark.parametrize('name', golden_cases())
def test_c_oracle_matches_reference_graph(name):
    g = load_golden(name)
    kp, kpl, res = c_oracle.fit_road_planes_c(g['boxes'], g['dimensions'], g['orientations'], g['P_inv'],
                                              g['planes_raw'])
    assert _same(kp, g['keypoints'])
    assert _same(kpl, g['keyplanes'])
    assert _same(res, g['residuals'])


@pytest.mark.parametrize('name', golden_cases())
@pytest.mark.parametrize('dtype', [np.float32, np.float64])
def test_c_oracle_equals_numpy_oracle(name, dtype):
    g = load_golden(name)
    a = fit_road_planes_ref(g['boxes'], g['dimensions'], g['orientations'], g['P_inv'], g['planes_raw'],
                            dtype=dtype, return_index=True)
    b = c_oracle.fit_road_planes_c(g['boxes'], g['dimensions'], g['orientations'], g['P_inv'], g['planes_raw'],
                                   dtype=dtype, return_index=True)
    for x, y in zip(a, b):
        assert _same(x, y)


def test_golden_set_covers_the_special_branches():
    g = load_golden('edge_2x12x24')
    sent = np.float32(100.0) / np.float32(6.0)
    assert (g['residuals'] == sent).sum() >= 3          # sentinel winners (all-masked / no-vote rows)
    assert (g['orientations'] == -1).any()              # padding rows
    assert g['planes'].shape == (2, 24, 4)              # per-image databases
    assert np.isnan(load_golden('allnan_1x3x2')['keypoints']).all()
