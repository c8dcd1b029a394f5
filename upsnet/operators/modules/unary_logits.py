"""reference path: upsnet/operators/modules/unary_logits.py (SegTerm :69-105; MaskTerm :24-66, the training twin)"""
from upsnet_b200.operators import MaskTerm, SegTerm  # noqa: F401
