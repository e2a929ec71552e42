"""CPU: the C-ABI library builds for sm_100a, loads, exports every symbol include/gpp.h declares, and fails
loudly (no CPU fallback) when no GPU is present."""
import ctypes
import os
import re

import pytest

from conftest import ROOT


def _declared_symbols(header='gpp.h'):
    with open(os.path.join(ROOT, 'include', header)) as f:
        text = f.read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(gpp_[a-z0-9_]+)\s*\(', text)))


def test_header_declares_the_expected_surface():
    syms = _declared_symbols()
    for must in ('gpp_create', 'gpp_destroy', 'gpp_set_planes', 'gpp_fit_host', 'gpp_fit_device',
                 'gpp_fit_host_f64', 'gpp_fit_device_f64', 'gpp_pose_host', 'gpp_pose_device', 'gpp_last_error'):
        assert must in syms


def test_library_exports_every_declared_symbol(gpp):
    lib = gpp._lib.load()
    declared = _declared_symbols('gpp.h') + _declared_symbols('gpp_debug.h')
    for name in declared:
        assert hasattr(lib, name), 'libgpp.so does not export %s' % name
        assert name in gpp._lib.SIGNATURES, 'no ctypes signature for %s' % name
    assert sorted(gpp._lib.SIGNATURES) == sorted(set(declared)), 'ctypes signatures without a declaration'
    assert lib.gpp_version() == 100


def test_debug_hooks_are_not_part_of_the_drop_in_header():
    """Measurement / tuning / test hooks live in include/gpp_debug.h; the boundary header holds only what a caller
    of fit_road_planes and its neighbouring steps binds."""
    public = _declared_symbols('gpp.h')
    assert not [s for s in public if s.startswith(('gpp_debug_', 'gpp_microbench', 'gpp_audit_'))]
    dbg = _declared_symbols('gpp_debug.h')
    for must in ('gpp_microbench', 'gpp_debug_scores', 'gpp_debug_set_schedule',
                 'gpp_audit_set', 'gpp_audit_counts'):
        assert must in dbg


def test_library_is_in_tree_and_sm100a(gpp):
    path = gpp._lib.LIB_PATH
    assert os.path.dirname(path) == os.path.join(ROOT, 'ground-plane-polling_b200')
    assert os.path.exists(path)


def test_python_surface_mirrors_the_reference(gpp):
    import inspect
    sig = inspect.signature(gpp.fit_road_planes)
    assert list(sig.parameters)[:5] == ['boxes', 'dimensions', 'orientations', 'P_inv', 'planes']
    layer = gpp.FitRoadPlanes()
    shapes = layer.compute_output_shape([(2, 100, 12), (2, 100, 3), (2, 100), (2, 4, 3), (2, 10, 4)])
    assert shapes == [(2, 100, 4, 3), (2, 100, 1, 4), (2, 100)]
    assert layer.compute_mask([1, 2, 3, 4, 5]) == [None] * 5
    assert 'name' in layer.get_config()


def test_no_cpu_fallback_without_a_gpu(gpp):
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    lib = gpp._lib.load()
    h = ctypes.c_void_p()
    rc = lib.gpp_create(0, ctypes.byref(h))
    assert rc != 0 and not h.value
    assert b'no CPU fallback' in lib.gpp_last_error()
    import numpy as np
    with pytest.raises(RuntimeError):
        gpp.fit_road_planes(np.zeros((1, 1, 12), np.float32), np.ones((1, 1, 3), np.float32),
                            np.zeros((1, 1), np.int32), np.zeros((1, 4, 3), np.float32),
                            np.ones((3, 4), np.float32))


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, 'ground-plane-polling_b200')
    for base, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                with open(os.path.join(base, f)) as fh:
                    src = fh.read()
                assert not re.search(r'^\s*(from|import)\s+oracle\b', src, flags=re.M), (base, f)
                assert 'libgpp_oracle' not in src, (base, f)


def test_scan_order_is_a_permutation_that_groups_similar_planes(gpp):
    """csrc/gpp_order.cu (host code, no device needed): the order the packed scans visit a database in is a
    permutation; small databases keep the index order; rows of 64 consecutive positions hold similar planes, and the
    first 16 rows are a sample of the whole parameter range."""
    import numpy as np
    from gpp_b200.layers.fit_road_planes import scan_order
    small = np.load(os.path.join(ROOT, 'road_planes_database', 'road_planes_database_1k.npy'))
    assert np.array_equal(scan_order(small), np.arange(small.shape[0]))
    for tag in ('10k', '22k'):
        planes = np.load(os.path.join(ROOT, 'road_planes_database', 'road_planes_database_%s.npy' % tag))
        order = scan_order(planes)
        n = planes.shape[0]
        assert np.array_equal(np.sort(order), np.arange(n))
        assert np.array_equal(order, scan_order(planes.astype(np.float32)))          # deterministic
        p = planes * -np.sign(planes[:, 1:2])
        p = p / np.linalg.norm(p[:, :3], axis=1, keepdims=True)
        key = p[:, [0, 2, 3]] / p[:, [0, 2, 3]].std(0)
        def spread(idx):                                     # typical extent of a row of 64 in the scaled parameters
            rows = key[idx[:(len(idx) // 64) * 64]].reshape(-1, 64, 3)
            return np.median(rows.max(1) - rows.min(1))
        assert spread(order[1024:]) < 0.5 * spread(np.arange(n)[1024:])
        seeds = key[order[:1024]]
        pr = lambda a: np.percentile(a, 99, axis=0) - np.percentile(a, 1, axis=0)    # noqa: E731  (the 10k database has outliers)
        assert (pr(seeds) > 0.6 * pr(key)).all()
