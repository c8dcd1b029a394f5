"""reference path: upsnet/operators/functions/deform_conv.py (forward only)"""
from upsnet_b200.operators import deform_conv  # noqa: F401


class DeformConvFunction:
    @staticmethod
    def apply(data, offset, weight, bias, in_channels, out_channels, kernel_size, stride, padding, dilation, groups,
              deformable_groups):
        return deform_conv(data, offset, weight, bias, stride, padding, dilation, deformable_groups)
