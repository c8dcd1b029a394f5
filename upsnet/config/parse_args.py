"""reference path: upsnet/config/parse_args.py:18-32 (same flags; parses, then merges the experiment yaml)."""
import argparse

from .config import config, update_config  # noqa: F401


def parse_args(description=""):
    parser = argparse.ArgumentParser(description=description)
    parser.add_argument("--cfg", help="experiment configure file name", required=True, type=str)
    parser.add_argument("--eval_only", help="if only eval existing results", action="store_true")
    parser.add_argument("--weight_path", help="manually specify model weights", type=str, default="")
    args, _rest = parser.parse_known_args()
    update_config(args.cfg)
    return parser.parse_args()
