"""reference path: upsnet/operators/modules/mod_deform_conv.py"""
from upsnet_b200.operators import ModDeformConv, ModDeformConvWithOffsetMask, ModulatedDeformConv  # noqa: F401
