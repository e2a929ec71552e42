"""CPU: the oracle validated without the reference -- geometric self-consistency, dtype agreement and the
special-value semantics."""
import numpy as np

from conftest import load_planes
from gpp_b200.utils import synthetic
from oracle import c_oracle
from oracle.fit_road_planes_ref import (fit_road_planes_ref, normalise_planes, second_best_gap,
                                        tf_argmin_last_axis)


def test_zero_noise_scene_is_recovered_for_all_orientations():
    planes = load_planes('1k')
    boxes, dims, orient, P_inv, truth = synthetic.synth_detections(2, 100, planes, seed=1, kp_noise_px=0.0,
                                                                   dim_noise=0.0, return_truth=True)
    assert set(np.unique(orient)) == {0, 1, 2, 3}
    kp, kpl, res, idx = c_oracle.fit_road_planes_c(boxes, dims, orient, P_inv, planes, return_index=True)
    assert np.median(res) < 0.1 and res.max() < 1.0      # a wrong orientation table would cost metres
    # bottom-face key-points land on the true object: compare the mid key-point depth with the truth
    assert np.median(np.abs(kp[..., 1, 2] - truth['location'][..., 2]) / truth['location'][..., 2]) < 0.1
    # recovered height = |X_t - X_m| agrees with the input height
    h = np.linalg.norm(kp[..., 3, :] - kp[..., 1, :], axis=-1)
    assert np.median(np.abs(h - dims[..., 0])) < 0.1


def test_fp32_and_fp64_agree_on_noisy_detections():
    planes = load_planes('22k')
    boxes, dims, orient, P_inv = synthetic.synth_detections(3, 100, planes, seed=9)
    a = c_oracle.fit_road_planes_c(boxes, dims, orient, P_inv, planes, return_index=True)
    b = c_oracle.fit_road_planes_c(boxes, dims, orient, P_inv, planes, return_index=True, dtype=np.float64)
    assert np.mean(a[3] == b[3]) > 0.98
    same = a[3] == b[3]
    assert np.allclose(a[0][same], b[0][same], rtol=1e-4, atol=1e-4)


def test_keyplanes_are_normalised_and_flipped():
    planes = load_planes('100')
    boxes, dims, orient, P_inv = synthetic.synth_detections(1, 30, planes, seed=3)
    kp, kpl, res, idx = fit_road_planes_ref(boxes, dims, orient, P_inv, planes, return_index=True)
    npl = normalise_planes(planes)
    assert np.array_equal(kpl[0, :, 0, :], npl[idx[0]])
    assert (kpl[..., 1] <= 0).all()
    assert np.allclose(np.linalg.norm(kpl[..., :3], axis=-1), 1.0, atol=1e-6)


def test_duplicate_planes_resolve_to_lowest_index():
    planes = load_planes('100')
    dup = np.concatenate([planes[:40], planes[:40], planes[:40]], axis=0)
    boxes, dims, orient, P_inv = synthetic.synth_detections(1, 50, planes, seed=4)
    idx = c_oracle.fit_road_planes_c(boxes, dims, orient, P_inv, dup, return_index=True)[3]
    assert (idx < 40).all()


def test_padding_rows_are_deterministic_and_harmless():
    planes = load_planes('100')
    boxes, dims, orient, P_inv = synthetic.synth_detections(1, 20, planes, seed=5, n_valid=12)
    a = fit_road_planes_ref(boxes, dims, orient, P_inv, planes)
    b = c_oracle.fit_road_planes_c(boxes, dims, orient, P_inv, planes)
    for x, y in zip(a, b):
        assert np.array_equal(x, y, equal_nan=True)
    assert (orient[0, 12:] == -1).all()


def test_tf_argmin_semantics():
    f = np.float32
    hi = np.finfo(f).max
    assert tf_argmin_last_axis(np.array([[np.nan, 3.0, 2.0, 2.0]], f))[0] == 2      # NaN never wins, first min
    assert tf_argmin_last_axis(np.array([[np.nan, np.nan]], f))[0] == 0             # nothing selectable -> 0
    assert tf_argmin_last_axis(np.array([[np.inf, hi, np.inf]], f))[0] == 0         # not below `highest` -> 0
    assert tf_argmin_last_axis(np.array([[np.inf, 100.0, 100.0]], f))[0] == 1


def test_all_masked_detection_returns_first_masked_plane_and_sentinel():
    planes = load_planes('100')
    boxes, dims, orient, P_inv = synthetic.synth_detections(1, 8, planes, seed=6)
    boxes = boxes.copy()
    # swap l and r key-points: z_dir_check < 0 for every plane -> everything carries the sentinel
    boxes[0, :, 4:6], boxes[0, :, 8:10] = boxes[0, :, 8:10].copy(), boxes[0, :, 4:6].copy()
    kp, kpl, res, idx = fit_road_planes_ref(boxes, dims, orient, P_inv, planes, return_index=True)
    assert (idx == 0).all()
    assert np.array_equal(res, np.full_like(res, np.float32(100.0) / np.float32(6.0)))


def test_second_best_gap_helper():
    planes = load_planes('100')
    boxes, dims, orient, P_inv = synthetic.synth_detections(1, 10, planes, seed=8)
    best, v1, v2, getter = second_best_gap(boxes, dims, orient, P_inv, planes)
    assert (v2 >= v1).all()
    assert getter(0, 0, int(best[0, 0])) == v1[0, 0]


def test_partial_residual_sum_never_exceeds_the_full_sum():
    """The kernels stop a hypothesis early when (r1 + r2) + r3 already exceeds the best residual sum.  That is exact
    because rounded addition of non-negative terms is monotone: the reference's
    ((((r0 + r1) + r2) + r3) + r4) + r5 is never below (r1 + r2) + r3, in float32 and float64, for any magnitudes."""
    rng = np.random.default_rng(11)
    for dt in (np.float32, np.float64):
        n = 400000
        mag = rng.choice([1e-30, 1e-7, 1e-3, 1.0, 37.5, 1e4, 1e20], size=(n, 6))
        r = (rng.random((n, 6)) * mag).astype(dt)
        r[rng.random((n, 6)) < 0.05] = 0
        with np.errstate(over='ignore'):
            full = ((((r[:, 0] + r[:, 1]) + r[:, 2]) + r[:, 3]) + r[:, 4]) + r[:, 5]
            part = (r[:, 1] + r[:, 2]) + r[:, 3]
        assert full.dtype == dt and (part <= full).all()
        # the oracle's own residual sums obey it too (bottom-face residuals of real hypotheses)
    from gpp_b200.utils import synthetic
    from conftest import load_planes
    from oracle import fit_road_planes_ref as R
    planes = load_planes('1k')
    boxes, dims, orient, P_inv = synthetic.synth_detections(1, 8, planes, seed=3)
    f = np.float32
    bx, dm, pi, pl = R._feed(boxes, dims, P_inv, planes[None], f)
    npl = R.normalise_planes(pl[0], f)
    rays = R.detection_rays(bx, pi, f).reshape(-1, 4, 3)
    td = R.detection_dims(dm, orient, f).reshape(-1, 6)
    (Xl, Xm, Xr, Xt), votes, resid, zc = R.hypotheses(rays, td, npl, f)
    with np.errstate(all='ignore'):
        s3 = (np.abs(R._dist(Xl, Xm) - td[:, 1:2]) + np.abs(R._dist(Xm, Xr) - td[:, 2:3])) + np.abs(R._dist(Xl, Xr) - td[:, 3:4])
    ok = np.isfinite(resid)
    assert (s3[ok] <= resid[ok]).all()
