"""
Image loading / preprocessing of the reference's driver (/root/reference/keras_retinanet_3D/utils/image.py:25-60
read + ImageNet-mean subtraction, :174-200 resize).  Only what ``bin/run_network.py`` needs to feed a detector and to
obtain the ``scale`` that the calibration and the boxes are corrected with; host-side (cv2 / numpy).
"""
import numpy as np

__all__ = ['read_image_bgr', 'preprocess_image', 'resize_image']

BGR_MEAN = (103.939, 116.779, 123.68)         # image.py:56-58


def read_image_bgr(path):
    """uint8 (rows, cols, 3) image in BGR order (image.py:25-32)."""
    import cv2
    img = cv2.imread(str(path), cv2.IMREAD_COLOR)
    if img is None:
        raise ValueError('cannot read image %s' % path)
    return img


def preprocess_image(x):
    """float32 copy with the ImageNet BGR mean subtracted, channels last (image.py:35-60, Keras floatx float32)."""
    x = np.asarray(x).astype(np.float32)
    x[..., 0] -= 103.939
    x[..., 1] -= 116.779
    x[..., 2] -= 123.68
    return x


def resize_image(img, min_side=800, max_side=1333):
    """Scale so that the smaller side is ``min_side`` unless the larger one would exceed ``max_side``
    (image.py:174-200).  Returns (image, scale); a KITTI frame (375 x 1242) becomes 402 x 1333."""
    import cv2
    rows, cols = img.shape[:2]
    scale = min_side / min(rows, cols)
    if max(rows, cols) * scale > max_side:
        scale = max_side / max(rows, cols)
    return cv2.resize(img, None, fx=scale, fy=scale), scale
