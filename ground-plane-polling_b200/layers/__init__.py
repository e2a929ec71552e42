from ._misc import RegressBoxes, RegressDims  # noqa: F401
from .filter_detections import FilterDetections  # noqa: F401
from .fit_road_planes import FitRoadPlanes, fit_road_planes  # noqa: F401
