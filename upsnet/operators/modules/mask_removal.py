"""reference path: upsnet/operators/modules/mask_removal.py"""
from upsnet_b200.operators import MaskRemoval  # noqa: F401
