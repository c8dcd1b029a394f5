"""reference path: upsnet/models/resnet_upsnet.py"""
from upsnet_b200.model import resnet_50_upsnet, resnet_101_upsnet, resnet_upsnet  # noqa: F401
