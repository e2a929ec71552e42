"""
numpy stand-ins for the TensorFlow / Keras ops that
/root/reference/keras_retinanet_3D/layers/fit_road_planes.py calls, so that the reference file can be
executed UNMODIFIED in a container without TensorFlow (golden-vector generation only; see make_golden.py).

Each stand-in implements the documented semantics of the TF op of the same name on numpy arrays, in
float32, with the canonical arithmetic stated in oracle/fit_road_planes_ref.py (left-to-right 3-term sums,
no FMA; Eigen's argmin reducer).  The op list is exactly what the reference file uses:
  keras.backend: shape abs greater less zeros_like ones_like permute_dimensions reshape concatenate ones
                 tile expand_dims sign sum max argmin cast stack floatx
  keras.layers.Layer (base class of FitRoadPlanes)
  keras_retinanet_3D.backend (backend/tensorflow_backend.py:20-157): norm where cross matmul multiply
                 divide gather gather_nd one_hot map_fn range
"""
import importlib.util
import sys
import types

import numpy as np

FLOATX = 'float32'


# ------------------------------------------------------------------ keras.backend
class _T(np.ndarray):
    """ndarray with the two Tensor methods the reference calls on results (set_shape, dtype.name checks)."""

    def set_shape(self, shape):
        assert tuple(int(s) for s in shape) == self.shape, (shape, self.shape)


def _t(x):
    return np.asarray(x).view(_T)


def _nms(boxes, scores, max_output_size, iou_threshold=0.5):
    """tf.image.non_max_suppression: greedy over descending scores (ties: lower index first -- an assumption,
    TF's order among exactly equal scores is implementation-defined), a box is suppressed iff its IoU with an
    already selected box is > iou_threshold; IoU as in TF's kernel (float32, area <= 0 never suppresses)."""
    boxes = np.asarray(boxes, dtype=np.float32)
    scores = np.asarray(scores, dtype=np.float32)
    order = np.lexsort((np.arange(scores.shape[0]), -scores))
    f = np.float32
    sel = []
    for i in order:
        if len(sel) >= int(max_output_size):
            break
        ok = True
        y1i, x1i = min(boxes[i, 0], boxes[i, 2]), min(boxes[i, 1], boxes[i, 3])
        y2i, x2i = max(boxes[i, 0], boxes[i, 2]), max(boxes[i, 1], boxes[i, 3])
        ai = f(f(y2i - y1i) * f(x2i - x1i))
        for j in sel:
            y1j, x1j = min(boxes[j, 0], boxes[j, 2]), min(boxes[j, 1], boxes[j, 3])
            y2j, x2j = max(boxes[j, 0], boxes[j, 2]), max(boxes[j, 1], boxes[j, 3])
            aj = f(f(y2j - y1j) * f(x2j - x1j))
            if ai <= 0 or aj <= 0:
                continue
            ih = max(f(min(y2i, y2j) - max(y1i, y1j)), f(0))
            iw = max(f(min(x2i, x2j) - max(x1i, x1j)), f(0))
            inter = f(ih * iw)
            iou = f(inter / f(f(ai + aj) - inter))
            if iou > f(iou_threshold):
                ok = False
                break
        if ok:
            sel.append(int(i))
    return np.asarray(sel, dtype=np.int32)


def _top_k(x, k):
    """tf.nn.top_k on a vector: values descending, ties by lower index."""
    x = np.asarray(x)
    order = np.lexsort((np.arange(x.shape[0]), -x))[:int(k)]
    return x[order], order.astype(np.int32)


def _kb():
    kb = types.ModuleType('keras.backend')
    kb.floatx = lambda: FLOATX
    kb.image_data_format = lambda: 'channels_last'
    kb.shape = lambda x: np.asarray(x).shape
    kb.abs = np.abs
    kb.greater = lambda a, b: np.greater(a, b)
    kb.less = lambda a, b: np.less(a, b)
    kb.zeros_like = np.zeros_like
    kb.ones_like = np.ones_like
    kb.permute_dimensions = lambda x, pattern: np.transpose(x, pattern)
    kb.reshape = lambda x, shape: np.reshape(x, tuple(int(s) for s in shape))
    kb.concatenate = lambda xs, axis=-1: np.concatenate(xs, axis=axis)
    kb.ones = lambda shape, dtype=None: np.ones(tuple(int(s) for s in shape), dtype=dtype or FLOATX)
    kb.tile = lambda x, n: np.tile(x, tuple(int(s) for s in n))
    kb.expand_dims = lambda x, axis=-1: np.expand_dims(x, axis)
    kb.sign = np.sign
    kb.sum = lambda x, axis=None, keepdims=False: _seq_sum(x, axis, keepdims)
    kb.max = lambda x, axis=None, keepdims=False: np.max(x, axis=axis, keepdims=keepdims)
    kb.argmin = _tf_argmin
    kb.cast = lambda x, dtype: _t(np.asarray(x).astype(dtype)) if np.ndim(x) else np.dtype(dtype).type(x)
    kb.stack = lambda xs, axis=0: np.stack([np.asarray(x) for x in xs], axis=axis)
    kb.argmax = lambda x, axis=-1: np.argmax(x, axis=axis).astype(np.int64)
    kb.minimum = lambda a, b: np.minimum(a, b)
    kb.maximum = lambda a, b: np.maximum(a, b)
    kb.int_shape = lambda x: tuple(np.asarray(x).shape)
    kb.transpose = np.transpose
    kb.constant = lambda v, dtype=None: np.asarray(v, dtype=dtype or FLOATX)
    kb.arange = lambda a, b=None, dtype='int32': np.arange(a, b, dtype=dtype)
    kb.ones = lambda shape, dtype=None: np.ones(tuple(int(s) for s in np.atleast_1d(shape)), dtype=dtype or FLOATX)
    return kb


def _seq_sum(x, axis, keepdims):
    """reduce_sum along one axis, accumulated left to right (canonical order)."""
    x = np.asarray(x)
    xs = np.moveaxis(x, axis, 0)
    acc = xs[0].copy()
    for i in range(1, xs.shape[0]):
        acc = acc + xs[i]
    return np.expand_dims(acc, axis) if keepdims else acc


def _tf_argmin(x, axis=-1):
    """tf.argmin == Eigen ArgMinTupleReducer: strict '<' scan from (index 0, highest finite)."""
    x = np.asarray(x)
    hi = np.finfo(x.dtype).max
    with np.errstate(invalid='ignore'):
        masked = np.where(x < hi, x, np.inf)
    return np.argmin(masked, axis=axis).astype(np.int64)


# ------------------------------------------------------------------ keras_retinanet_3D.backend
def _norm(x, ord='euclidean', axis=None, keep_dims=False):
    x = np.asarray(x)
    with np.errstate(all='ignore'):
        return np.sqrt(_seq_sum(x * x, axis, keep_dims))


def _cross(a, b):
    a, b = np.broadcast_arrays(np.asarray(a), np.asarray(b))
    with np.errstate(all='ignore'):
        return np.stack([a[..., 1] * b[..., 2] - a[..., 2] * b[..., 1],
                         a[..., 2] * b[..., 0] - a[..., 0] * b[..., 2],
                         a[..., 0] * b[..., 1] - a[..., 1] * b[..., 0]], axis=-1)


def _matmul(a, b):
    """Batched matmul, inner dimension accumulated left to right without FMA."""
    a, b = np.asarray(a), np.asarray(b)
    K = a.shape[-1]
    assert b.shape[-2] == K
    with np.errstate(all='ignore'):
        acc = a[..., :, 0:1] * b[..., 0:1, :]
        for k in range(1, K):
            acc = acc + a[..., :, k:k + 1] * b[..., k:k + 1, :]
    return acc


def _divide(a, b):
    with np.errstate(all='ignore'):
        return np.divide(a, b)


def _multiply(a, b):
    with np.errstate(all='ignore'):
        return np.multiply(a, b)


def _gather(params, indices, axis=0):
    return np.take(np.asarray(params), np.asarray(indices), axis=axis)


def _gather_nd(params, indices):
    params, indices = np.asarray(params), np.asarray(indices)
    return params[tuple(indices[..., i] for i in range(indices.shape[-1]))]


def _one_hot(indices, depth, dtype=FLOATX):
    return (np.asarray(indices)[..., None] == np.arange(depth)).astype(dtype)


def _map_fn(fn, elems, dtype=None, parallel_iterations=None):
    n = np.asarray(elems[0]).shape[0]
    pick = lambda e, i: [o[i] for o in e] if isinstance(e, (list, tuple)) else e[i]  # noqa: E731
    outs = [fn([pick(e, i) for e in elems]) for i in range(n)]
    if isinstance(outs[0], (list, tuple)):
        return [np.stack([np.asarray(o[k]) for o in outs], axis=0) for k in range(len(outs[0]))]
    return np.stack([np.asarray(o) for o in outs], axis=0)


def _backend():
    be = types.ModuleType('keras_retinanet_3D.backend')
    be.norm = _norm
    be.where = lambda c, a=None, b=None: (np.argwhere(c).astype(np.int64) if a is None else np.where(c, a, b))
    be.non_max_suppression = _nms
    be.top_k = _top_k
    be.pad = lambda x, paddings, constant_values=0: _t(np.pad(np.asarray(x), [tuple(int(v) for v in p) for p in paddings],
                                                                mode='constant', constant_values=constant_values))
    be.clip_by_value = np.clip
    be.cross = _cross
    be.matmul = _matmul
    be.multiply = _multiply
    be.divide = _divide
    be.gather = _gather
    be.gather_nd = _gather_nd
    be.one_hot = _one_hot
    be.map_fn = _map_fn
    be.range = lambda n: np.arange(int(n), dtype=np.int64)
    return be


class _Layer(object):
    def __init__(self, **kwargs):
        self.name = kwargs.get('name', self.__class__.__name__.lower())

    def get_config(self):
        return {'name': self.name, 'trainable': True}

    def __call__(self, inputs, **kwargs):
        return self.call(inputs, **kwargs)


def load_reference_module(path='/root/reference/keras_retinanet_3D/layers/fit_road_planes.py',
                          name='keras_retinanet_3D.layers.fit_road_planes', with_common=False):
    """Execute one of the reference's files unmodified with the numpy stand-ins installed.  ``with_common``
    additionally executes backend/common.py (dim_transform_inv, bbox_transform_inv) into the stand-in backend."""
    saved = {k: sys.modules.get(k) for k in
             ('keras', 'keras.backend', 'keras.layers', 'keras_retinanet_3D', 'keras_retinanet_3D.backend',
              'keras_retinanet_3D.layers')}
    keras = types.ModuleType('keras')
    keras.backend = _kb()
    keras.layers = types.ModuleType('keras.layers')
    keras.layers.Layer = _Layer
    pkg = types.ModuleType('keras_retinanet_3D')
    pkg.__path__ = []
    pkg.backend = _backend()
    lay = types.ModuleType('keras_retinanet_3D.layers')
    lay.__path__ = []
    sys.modules.update({'keras': keras, 'keras.backend': keras.backend, 'keras.layers': keras.layers,
                        'keras_retinanet_3D': pkg, 'keras_retinanet_3D.backend': pkg.backend,
                        'keras_retinanet_3D.layers': lay})
    utils = types.ModuleType('keras_retinanet_3D.utils')
    utils.__path__ = []
    utils.anchors = types.ModuleType('keras_retinanet_3D.utils.anchors')
    pkg.utils = utils
    saved.update({k: sys.modules.get(k) for k in ('keras_retinanet_3D.utils', 'keras_retinanet_3D.utils.anchors',
                                                  'keras_retinanet_3D.backend.dynamic')})
    sys.modules.update({'keras_retinanet_3D.utils': utils, 'keras_retinanet_3D.utils.anchors': utils.anchors})
    try:
        if with_common:
            pkg.backend.__path__ = []
            dyn = types.ModuleType('keras_retinanet_3D.backend.dynamic')
            dyn.meshgrid = np.meshgrid
            sys.modules['keras_retinanet_3D.backend.dynamic'] = dyn
            cspec = importlib.util.spec_from_file_location(
                'keras_retinanet_3D.backend.common', '/root/reference/keras_retinanet_3D/backend/common.py')
            cmod = importlib.util.module_from_spec(cspec)
            cspec.loader.exec_module(cmod)
            for fn in ('dim_transform_inv', 'bbox_transform_inv', 'shift'):
                setattr(pkg.backend, fn, getattr(cmod, fn))
        spec = importlib.util.spec_from_file_location(name, path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    return mod
