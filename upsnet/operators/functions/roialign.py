"""reference path: upsnet/operators/functions/roialign.py"""
from upsnet_b200.operators import RoIAlignFunction  # noqa: F401
