"""
gpp_b200 (directory ``ground-plane-polling_b200``): B200-native ground-plane polling.

The one hot path of arangesh/Ground-Plane-Polling -- the per-detection search over the road-plane database
(keras_retinanet_3D/layers/fit_road_planes.py) plus the pose recovery that consumes it
(keras_retinanet_3D/bin/run_network.py:137-247) -- as hand-written CUDA for sm_100a behind the reference's
own call surface.  Import as ``import gpp_b200`` (root-level alias) or
``importlib.import_module('ground-plane-polling_b200')``.
"""
from . import _lib  # noqa: F401
from .layers import fit_road_planes as _frp_module  # noqa: F401
from .layers.fit_road_planes import (FitRoadPlanes, PlanePoller, fit_road_planes, fit_road_planes_dlpack,  # noqa: F401
                                     fit_road_planes_torch, get_poller)
from .layers._misc import RegressBoxes, RegressDims, decode, decode_torch  # noqa: F401
from .layers.filter_detections import (FilterDetections, filter_detections, filter_detections_batch,  # noqa: F401
                                       filter_detections_torch)
from . import pipeline, sharding  # noqa: F401
from .sharding import fit_road_planes_multi, fit_road_planes_sharded, shard_bounds  # noqa: F401
from .pipeline import detections_from_heads  # noqa: F401
from .utils import synthetic  # noqa: F401
from .utils import pose  # noqa: F401
from .utils.pose import recover_pose, recover_pose_torch  # noqa: F401
from .utils import calibration, kitti  # noqa: F401
from .utils.calibration import load_calibration, load_road_planes  # noqa: F401
from .utils.kitti import image_detections, kitti_records, postprocess_image, select_detections  # noqa: F401

__version__ = '0.1.0'
