"""
Anchor boxes of the detector, as the reference lays them out
(/root/reference/keras_retinanet_3D/utils/anchors.py:140-265): pyramid levels 3..7, stride 2^p, base size 2^(p+2),
3 ratios x 4 scales = 12 anchors per cell, cells in row-major order, the 12 anchors of a cell contiguous, levels
concatenated.  137,256 anchors for a KITTI image resized to 402 x 1333.  Host-side numpy (computed once per image
shape and cached); float64 like the reference.
"""
import functools

import numpy as np

__all__ = ['generate_anchors', 'shift', 'guess_shapes', 'anchors_for_shape', 'cached_anchors']

DEFAULT_RATIOS = (0.5, 1.0, 2.0)
DEFAULT_SCALES = (2 ** (-2.0 / 3.0), 2 ** 0, 2 ** (1.0 / 3.0), 2 ** (2.0 / 3.0))     # anchors.py:186


def generate_anchors(base_size=16, ratios=None, scales=None):
    """The len(ratios) * len(scales) anchors of one cell, centred on the origin, (x1, y1, x2, y2); ratio-major
    order (anchors.py:234-265).  Width = sqrt(area / ratio), height = width * ratio."""
    ratios = np.asarray(DEFAULT_RATIOS if ratios is None else ratios, dtype=np.float64)
    scales = np.asarray(DEFAULT_SCALES if scales is None else scales, dtype=np.float64)
    side = base_size * np.tile(scales, len(ratios))                # scale index runs fastest
    ratio = np.repeat(ratios, len(scales))
    w = np.sqrt((side * side) / ratio)
    h = w * ratio
    out = np.zeros((w.shape[0], 4))
    out[:, 0] = 0.0 - w * 0.5
    out[:, 1] = 0.0 - h * 0.5
    out[:, 2] = w - w * 0.5
    out[:, 3] = h - h * 0.5
    return out


def shift(shape, stride, anchors):
    """Copies of the cell anchors at every cell centre (x + 0.5, y + 0.5) * stride of a (rows, cols) map
    (anchors.py:203-231)."""
    cx = (np.arange(0, shape[1]) + 0.5) * stride
    cy = (np.arange(0, shape[0]) + 0.5) * stride
    gx, gy = np.meshgrid(cx, cy)
    centres = np.stack([gx.ravel(), gy.ravel(), gx.ravel(), gy.ravel()], axis=1)      # (K, 4)
    return (centres[:, None, :] + np.asarray(anchors)[None, :, :]).reshape(-1, 4)


def guess_shapes(image_shape, pyramid_levels):
    """Feature-map shape per level: ceil(side / 2^p) (anchors.py:140-152)."""
    side = np.array(image_shape[:2])
    return [(side + 2 ** p - 1) // (2 ** p) for p in pyramid_levels]


def anchors_for_shape(image_shape, pyramid_levels=None, ratios=None, scales=None, strides=None, sizes=None,
                      shapes_callback=None):
    """All anchors of an image of ``image_shape`` (rows, cols[, channels]): (A, 4) float64 (anchors.py:155-200)."""
    levels = [3, 4, 5, 6, 7] if pyramid_levels is None else list(pyramid_levels)
    strides = [2 ** p for p in levels] if strides is None else strides
    sizes = [2 ** (p + 2) for p in levels] if sizes is None else sizes
    shapes = (guess_shapes if shapes_callback is None else shapes_callback)(image_shape, levels)
    per_level = [shift(shapes[i], strides[i], generate_anchors(sizes[i], ratios, scales)) for i in range(len(levels))]
    return np.concatenate([np.zeros((0, 4))] + per_level, axis=0)


@functools.lru_cache(maxsize=8)
def cached_anchors(rows, cols):
    """float32 anchors of a (rows, cols) image with the default configuration, cached per shape (read-only)."""
    a = anchors_for_shape((rows, cols)).astype(np.float32)
    a.setflags(write=False)
    return a
