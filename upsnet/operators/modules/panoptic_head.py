"""new module path for the packaged panoptic head (reference: inline in models/resnet_upsnet.py:217-247)"""
from upsnet_b200.operators import PanopticHead  # noqa: F401
