from .yolov8 import YOLOv8, build_yolov8, level_shapes  # noqa: F401
from .mnv3 import build_trunk, calibrate_bn, fold_bn  # noqa: F401
