"""reference path: upsnet/config/config.py -- the hot-path knobs only (see upsnet_b200.model.UPSNetConfig)."""
from upsnet_b200.model import UPSNetConfig  # noqa: F401

config = UPSNetConfig()
