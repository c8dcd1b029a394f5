"""reference path: upsnet/models/resnet_upsnet.py:250-257 -- the two factories `upsnet_end2end_test.py:162` evaluates
(`eval(config.symbol)()`), zero-argument like the reference's: the architecture is read from `upsnet.config.config.config`
(the reference's own module in an overlay, this repository's subset otherwise)."""
from upsnet_b200.model import UPSNetConfig, resnet_upsnet  # noqa: F401
from upsnet_b200 import model as _m


def _cfg():
    from upsnet.config.config import config
    return UPSNetConfig.from_reference_config(config)


def resnet_50_upsnet():
    return _m.resnet_50_upsnet(_cfg())


def resnet_101_upsnet():
    return _m.resnet_101_upsnet(_cfg())
