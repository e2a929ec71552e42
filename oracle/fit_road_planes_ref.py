"""
ORACLE (test infrastructure, NOT product code) -- numpy restatement of the ground-plane polling graph.

Follows /root/reference/keras_retinanet_3D/layers/fit_road_planes.py:18-139 op by op (``poll`` :18-32,
``calc_X_t`` :34-47, ``fit_road_planes`` :49-139).  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import this module; the product path
(``ground-plane-polling_b200``) never does.

Pinning status
--------------
The reference ships no test, golden vector or known-answer fixture for this path and TensorFlow/Keras are
not installable here, so the reference graph cannot be executed on its own runtime.  The oracle is instead
pinned against the reference's *own source file* executed unmodified over a numpy stand-in for the handful
of TF/Keras ops it calls (``tests/golden/make_golden.py`` + ``tests/golden/tf_numpy_shim.py``); the
resulting vectors are committed under ``tests/golden/`` and checked bit-for-bit in
``tests/test_oracle_golden.py``.  What remains unpinned is the ulp-level rounding of the TF kernels
themselves (summation order / FMA contraction inside ``tf.matmul``, which TF does not define across
devices): "parity unpinned at the ulp level, pinned at the graph level".

Canonical arithmetic (the contract the CUDA ``exact`` mode reproduces bit for bit)
--------------------------------------------------------------------------------
* everything in ``dtype`` (float32 by default, float64 for the verify mode), IEEE round-to-nearest,
  no FMA contraction, denormals kept;
* 3-term sums are evaluated left to right: ``(x0*y0 + x1*y1) + x2*y2``;
* ``tf.cross(a, b) = (a1*b2 - a2*b1, a2*b0 - a0*b2, a0*b1 - a1*b0)``;
* division and square root are the correctly rounded IEEE operations;
* ``tf.argmin`` semantics (Eigen ArgMinTupleReducer): scan with strict ``<`` starting from
  ``(index 0, value = highest finite)``, so NaN / +inf / FLT_MAX entries are never selected and an
  all-unselectable row returns 0.  (Assumption: TF is absent here, so this follows Eigen's published
  reducer; numpy's own argmin would return the first NaN instead.)
"""
import numpy as np

VOTE_THRESHOLD = 0.7   # metres, fit_road_planes.py:94
SENTINEL = 100.0       # fit_road_planes.py:117-118
N_POLLS = 6.0          # fit_road_planes.py:131


def tf_argmin_last_axis(r):
    """tf.argmin over the last axis with Eigen's reducer semantics (see module docstring)."""
    hi = np.finfo(r.dtype).max
    masked = np.where(r < hi, r, np.inf)          # NaN, +inf and `highest` can never win
    return np.argmin(masked, axis=-1).astype(np.int64)


def normalise_planes(planes, dtype=np.float32):
    """fit_road_planes.py:75-77 -- flip so the normal's y is <= 0, then divide all 4 coeffs by |n|."""
    p = np.ascontiguousarray(planes, dtype=dtype)
    direction = -np.sign(p[..., 1:2])
    p = p * direction
    a, b, c = p[..., 0], p[..., 1], p[..., 2]
    rho = np.sqrt((a * a + b * b) + c * c)[..., None]
    with np.errstate(invalid='ignore', divide='ignore'):
        return (p / rho).astype(dtype)


def detection_rays(boxes, P_inv, dtype=np.float32):
    """fit_road_planes.py:80-83 -- 4 homogeneous pixels -> P_inv(4x3) @ x -> rows 0..2 -> times sign(z).

    boxes (B, D, 12), P_inv (B, 4, 3)  ->  rays (B, D, 4 keypoints [l, m, r, t], 3 xyz)
    """
    boxes = np.asarray(boxes, dtype=dtype)
    P_inv = np.asarray(P_inv, dtype=dtype)
    B, D = boxes.shape[:2]
    uv = boxes[:, :, 4:].reshape(B, D, 4, 2)
    u, v = uv[..., 0], uv[..., 1]                                 # (B, D, 4)
    one = dtype(1.0)
    rows = []
    for r in range(3):                                            # 4th row of P_inv @ x is discarded (:83)
        p0 = P_inv[:, r, 0][:, None, None]
        p1 = P_inv[:, r, 1][:, None, None]
        p2 = P_inv[:, r, 2][:, None, None]
        rows.append((p0 * u + p1 * v) + p2 * one)
    g = np.stack(rows, axis=-1)                                   # (B, D, 4, 3)
    return (g * np.sign(g[..., 2:3])).astype(dtype)


def detection_dims(dimensions, orientations, dtype=np.float32):
    """fit_road_planes.py:66-72,97-108 -- the six target lengths per detection.

    Returns (B, D, 6): [h, e1, e2, d_wl, f1, f2] matching polls 0..5.
    """
    dims = np.asarray(dimensions, dtype=dtype)
    o = np.asarray(orientations)
    h, w, l = dims[..., 0], dims[..., 1], dims[..., 2]
    d_hw = np.sqrt(h * h + w * w)
    d_wl = np.sqrt(w * w + l * l)
    d_hl = np.sqrt(h * h + l * l)
    one_hot = (o[..., None] == np.arange(4)).astype(dtype)        # class -1 -> all-zero row (:72)

    def pick(c0, c1, c2, c3):
        s = one_hot[..., 0] * c0
        s = s + one_hot[..., 1] * c1
        s = s + one_hot[..., 2] * c2
        s = s + one_hot[..., 3] * c3
        return s

    e1 = pick(l, w, w, l)
    e2 = pick(w, l, l, w)
    f1 = pick(d_hl, d_hw, d_hw, d_hl)
    f2 = pick(d_hw, d_hl, d_hl, d_hw)
    return np.stack([h, e1, e2, d_wl, f1, f2], axis=-1).astype(dtype)


def _cross(a, b):
    return (a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0])


def _dot3(a, b):
    return (a[0] * b[0] + a[1] * b[1]) + a[2] * b[2]


def _dist(a, b):
    dx, dy, dz = a[0] - b[0], a[1] - b[1], a[2] - b[2]
    return np.sqrt((dx * dx + dy * dy) + dz * dz)


def hypotheses(rays, tdims, nplanes, dtype=np.float32):
    """All (detection, plane) hypotheses for a flat list of detections against ONE normalised DB.

    rays (M, 4, 3), tdims (M, 6), nplanes (N, 4).
    Returns X (4 keypoints x 3 xyz, each (M, N)), votes (M, N), residual sum (M, N), z_dir_check (M, N).
    fit_road_planes.py:84-113.
    """
    n = tuple(nplanes[None, :, i] for i in range(3))              # (1, N)
    d4 = nplanes[None, :, 3]
    X = []
    with np.errstate(all='ignore'):
        for k in range(3):                                        # l, m, r  (:86-87)
            dk = tuple(rays[:, k, i][:, None] for i in range(3))  # (M, 1)
            t = _dot3(n, dk)
            s = np.abs((-d4) / t)
            X.append(tuple(dk[i] * s for i in range(3)))
        Xl, Xm, Xr = X
        a = tuple(Xl[i] - Xm[i] for i in range(3))
        b = tuple(Xr[i] - Xm[i] for i in range(3))
        zc = a[2] * b[0] - a[0] * b[2]                            # y of cross (:88-89)
        dt = tuple(rays[:, 3, i][:, None] for i in range(3))
        c = _cross(n, dt)                                         # calc_X_t (:43-46)
        perp = _cross(dt, c)
        num = _dot3(perp, Xm)
        den = _dot3(perp, n)
        q = num / den
        Xt = tuple(Xm[i] - q * n[i] for i in range(3))
        pairs = ((Xm, Xt), (Xl, Xm), (Xm, Xr), (Xl, Xr), (Xl, Xt), (Xr, Xt))   # (:95-109)
        thr = dtype(VOTE_THRESHOLD)
        votes = None
        resid = None
        for k, (pa, pb) in enumerate(pairs):
            r = np.abs(_dist(pa, pb) - tdims[:, k][:, None])
            v = np.where(r > thr, dtype(0.0), dtype(1.0))         # NaN > thr is False -> a vote (:31)
            votes = v if votes is None else votes + v
            resid = r if resid is None else resid + r
    return (Xl, Xm, Xr, Xt), votes, resid, zc


def select(votes, resid, zc, dtype=np.float32):
    """fit_road_planes.py:116-119 -- max-vote mask, z-check mask, tf.argmin."""
    with np.errstate(all='ignore'):
        v = votes - votes.max(axis=1, keepdims=True)
        r = np.where(v < 0, dtype(SENTINEL), resid)
        r = np.where(zc < 0, dtype(SENTINEL), r)                  # NaN < 0 is False -> passes
    return tf_argmin_last_axis(r), r


def _feed(boxes, dimensions, P_inv, planes, dtype):
    """Keras casts every fed array to floatx = float32 (run_network.py:105 feeds float64 P_inv / planes);
    the float64 verify mode starts from those SAME float32 values, promoted exactly."""
    f32 = np.float32
    boxes = np.asarray(boxes, dtype=f32).astype(dtype)
    dimensions = np.asarray(dimensions, dtype=f32).astype(dtype)
    P_inv = np.asarray(P_inv, dtype=f32).astype(dtype)
    planes = np.asarray(planes, dtype=f32).astype(dtype)
    if planes.ndim == 2:
        planes = planes[None]
    return boxes, dimensions, P_inv, planes


def fit_road_planes_ref(boxes, dimensions, orientations, P_inv, planes, dtype=np.float32,
                        return_index=False, chunk=16):
    """numpy restatement of ``fit_road_planes`` (fit_road_planes.py:49-139).

    Args (same order/shapes as the reference, :52-56)
        boxes (B, D, 12), dimensions (B, D, 3), orientations (B, D) int, P_inv (B, 4, 3),
        planes (B, N, 4) raw road planes (also accepts (N, 4) / (1, N, 4): one DB for every image).
    Returns [keypoints (B, D, 4, 3), keyplanes (B, D, 1, 4), residuals (B, D)] (+ best index (B, D) int64).
    """
    dtype = np.dtype(dtype).type
    boxes, dimensions, P_inv, planes = _feed(boxes, dimensions, P_inv, planes, dtype)
    B, D = boxes.shape[:2]
    rays = detection_rays(boxes, P_inv, dtype)
    tdims = detection_dims(dimensions, orientations, dtype)
    keypoints = np.empty((B, D, 4, 3), dtype)
    keyplanes = np.empty((B, D, 1, 4), dtype)
    residuals = np.empty((B, D), dtype)
    best = np.empty((B, D), np.int64)
    shared = planes.shape[0] == 1
    npl_shared = normalise_planes(planes[0], dtype) if shared else None
    for bi in range(B):
        npl = npl_shared if shared else normalise_planes(planes[bi], dtype)
        for d0 in range(0, D, chunk):
            sl = slice(d0, min(D, d0 + chunk))
            X, votes, resid, zc = hypotheses(rays[bi, sl], tdims[bi, sl], npl, dtype)
            idx, r = select(votes, resid, zc, dtype)
            rows = np.arange(idx.shape[0])
            best[bi, sl] = idx
            keyplanes[bi, sl, 0, :] = npl[idx]                    # normalised, flipped plane (:122)
            with np.errstate(all='ignore'):
                residuals[bi, sl] = r[rows, idx] / dtype(N_POLLS)   # (:130-131)
            for k in range(4):
                for i in range(3):
                    keypoints[bi, sl, k, i] = X[k][i][rows, idx]  # (:133-137)
    out = [keypoints, keyplanes, residuals]
    if return_index:
        out.append(best)
    return out


def second_best_gap(boxes, dimensions, orientations, P_inv, planes, dtype=np.float64, chunk=16):
    """Test helper for the near-tie rule: per detection, the masked residual row's best and second-best
    DISTINCT values (in ``dtype``), so tests can tell a legitimate near-tie from a real mismatch.

    Returns (best_idx (B, D), best_val (B, D), second_val (B, D), masked residual getter) where
    ``getter(b, d, j)`` returns R''[b, d, j].
    """
    dtype = np.dtype(dtype).type
    boxes, dimensions, P_inv, planes = _feed(boxes, dimensions, P_inv, planes, dtype)
    B, D = boxes.shape[:2]
    rays = detection_rays(boxes, P_inv, dtype)
    tdims = detection_dims(dimensions, orientations, dtype)
    best = np.zeros((B, D), np.int64)
    v1 = np.zeros((B, D), dtype)
    v2 = np.zeros((B, D), dtype)
    rows_all = {}
    for bi in range(B):
        npl = normalise_planes(planes[0 if planes.shape[0] == 1 else bi], dtype)
        for d0 in range(0, D, chunk):
            sl = slice(d0, min(D, d0 + chunk))
            _, votes, resid, zc = hypotheses(rays[bi, sl], tdims[bi, sl], npl, dtype)
            idx, r = select(votes, resid, zc, dtype)
            best[bi, sl] = idx
            for i in range(r.shape[0]):
                row = r[i]
                rows_all[(bi, d0 + i)] = row
                fin = np.sort(row[np.isfinite(row)])
                v1[bi, d0 + i] = fin[0] if fin.size else np.nan
                bigger = fin[fin > fin[0]] if fin.size else fin
                v2[bi, d0 + i] = bigger[0] if bigger.size else np.inf
    return best, v1, v2, (lambda b, d, j: rows_all[(b, d)][j])
