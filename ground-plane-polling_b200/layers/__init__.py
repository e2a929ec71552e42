from .fit_road_planes import FitRoadPlanes, fit_road_planes  # noqa: F401
