"""reference path: upsnet/operators/modules/deform_conv.py"""
from upsnet_b200.operators import DeformConv, DeformConvWithOffset  # noqa: F401
