"""reference path: upsnet/operators/modules/mask_matching.py:27-62 (MaskMatching: panoptic ground-truth assembly)"""
from upsnet_b200.operators import MaskMatching  # noqa: F401
