"""reference path: upsnet/operators/modules/fpn_roi_align.py"""
from upsnet_b200.operators import FPNRoIAlign  # noqa: F401
