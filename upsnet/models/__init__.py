"""reference path: upsnet/models/__init__.py (`from .resnet_upsnet import resnet_50_upsnet, resnet_101_upsnet`)."""
from .resnet_upsnet import resnet_50_upsnet, resnet_101_upsnet, resnet_upsnet  # noqa: F401
