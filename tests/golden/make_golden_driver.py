"""
Generates tests/golden/driver_utils.npz by running the reference's OWN utilities (run in the build container, where
/root/reference exists): utils/anchors.py is imported as is (numpy only); utils/image.py is executed over the Keras
stand-in of tf_numpy_shim (it only asks Keras for floatx / image_data_format).

  anchors_small      anchors_for_shape((96, 160, 3))                       full array
  anchors_kitti_*    anchors_for_shape((402, 1333, 3)): row count, sha256 of the float64 bytes, every 997th row
  image_*            a seeded 45 x 150 uint8 image -> preprocess_image -> resize_image (+ scale); KITTI scale
"""
import hashlib
import importlib.util
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import tf_numpy_shim  # noqa: E402

REF = '/root/reference/keras_retinanet_3D'


def main():
    spec = importlib.util.spec_from_file_location('ref_anchors', os.path.join(REF, 'utils', 'anchors.py'))
    anchors = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(anchors)
    import types
    stub = types.ModuleType('keras_retinanet_3D.utils.transform')      # image.py imports it for the augmentation
    stub.change_transform_origin = None                                 # helpers, which are not exercised here
    sys.modules['keras_retinanet_3D.utils.transform'] = stub
    image = tf_numpy_shim.load_reference_module(os.path.join(REF, 'utils', 'image.py'),
                                                'keras_retinanet_3D.utils.image', with_common=False)

    out = {}
    out['anchors_small'] = anchors.anchors_for_shape((96, 160, 3))
    big = anchors.anchors_for_shape((402, 1333, 3))
    out['anchors_kitti_rows'] = np.int64(big.shape[0])
    out['anchors_kitti_sha256'] = np.frombuffer(hashlib.sha256(np.ascontiguousarray(big).tobytes()).digest(), dtype=np.uint8)
    out['anchors_kitti_sample'] = big[::997]

    rng = np.random.default_rng(7)
    raw = rng.integers(0, 256, size=(45, 150, 3), dtype=np.uint8)
    pre = image.preprocess_image(raw.copy())
    resized, scale = image.resize_image(pre)                       # default sides: 400 x 1333, stored as digest + sample
    small, small_scale = image.resize_image(pre, min_side=24, max_side=64)
    out['image_raw'] = raw
    out['image_preprocessed'] = pre
    out['image_resized_shape'] = np.array(resized.shape)
    out['image_resized_sha256'] = np.frombuffer(hashlib.sha256(np.ascontiguousarray(resized).tobytes()).digest(), dtype=np.uint8)
    out['image_resized_sample'] = resized.reshape(-1)[::4099]
    out['image_scale'] = np.float64(scale)
    out['image_small'] = small
    out['image_small_scale'] = np.float64(small_scale)
    _, kitti_scale = image.resize_image(np.zeros((375, 1242, 3), np.float32))
    out['kitti_scale'] = np.float64(kitti_scale)
    np.savez_compressed(os.path.join(HERE, 'driver_utils.npz'), **out)
    print({k: getattr(v, 'shape', v) for k, v in out.items()})


if __name__ == '__main__':
    main()
