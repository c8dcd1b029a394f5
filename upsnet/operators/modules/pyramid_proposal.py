"""reference path: upsnet/operators/modules/pyramid_proposal.py:24-67"""
from upsnet_b200.detection import PyramidProposal  # noqa: F401
