"""reference path: upsnet/operators/modules/mask_roi.py:24-146"""
from upsnet_b200.detection import MaskROIModule as MaskROI  # noqa: F401
