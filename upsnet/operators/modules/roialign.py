"""reference path: upsnet/operators/modules/roialign.py"""
from upsnet_b200.operators import RoIAlign, ROIAlign  # noqa: F401
