"""reference path: upsnet/operators/modules/unary_logits.py (SegTerm; MaskTerm is the training twin, config #4)"""
from upsnet_b200.operators import SegTerm  # noqa: F401
