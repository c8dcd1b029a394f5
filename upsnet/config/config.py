"""reference path: upsnet/config/config.py.

If a reference checkout is importable further down the `upsnet` namespace path, ITS config module is executed and
re-exported verbatim (all ~150 knobs the reference's dataset / training code reads).  Otherwise this file provides the
hot-path subset with the reference's default values (config.py:19-174) and `update_config` (config.py:177-198)."""
import importlib.util
import os
import sys

import numpy as np


def _reference_config_module():
    import upsnet
    here = os.path.realpath(__file__)
    for p in list(getattr(upsnet, "__path__", [])):
        cand = os.path.join(p, "config", "config.py")
        if os.path.exists(cand) and os.path.realpath(cand) != here:
            try:
                spec = importlib.util.spec_from_file_location("upsnet.config._reference_config", cand)
                mod = importlib.util.module_from_spec(spec)
                spec.loader.exec_module(mod)
                return mod
            except Exception:      # e.g. easydict missing: fall back to the subset below
                return None
    return None


class AttrDict(dict):
    """easydict.EasyDict semantics for the calls the reference makes (attribute access, nested dicts, item access)."""

    def __init__(self, d=None, **kw):
        super().__init__()
        for k, v in dict(d or {}, **kw).items():
            self[k] = v

    def __setitem__(self, k, v):
        super().__setitem__(k, AttrDict(v) if isinstance(v, dict) and not isinstance(v, AttrDict) else v)

    __setattr__ = __setitem__

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)


_ref = _reference_config_module()
if _ref is not None:
    config, update_config = _ref.config, _ref.update_config
else:
    config = AttrDict()
    config.debug_mode = False
    config.symbol = "resnet_50_upsnet"
    config.gpus = "0"
    config.network = AttrDict(
        backbone_fix_bn=True, backbone_with_dilation=False, backbone_with_dpyramid=False, backbone_with_dconv=100,
        backbone_freeze_at=2, use_caffe_model=True, use_syncbn=False, has_rcnn=True, has_mask_head=True, has_fcn_head=False,
        has_panoptic_head=False, pixel_means=np.array((102.9801, 115.9465, 122.7717,)), cls_agnostic_bbox_reg=False,
        rcnn_feat_stride=32, bbox_reg_weights=(10., 10., 5., 5.,), rpn_feat_stride=(4, 8, 16, 32, 64,),
        anchor_ratios=(0.5, 1, 2), anchor_scales=(8,), num_anchors=3, rpn_with_norm="none", has_fpn=True,
        fpn_feature_dim=256, fpn_with_gap=False, fpn_upsample_method="nearest", fpn_with_norm="none",
        rcnn_with_norm="none", mask_size=28, binary_thresh=0.5, has_mask_rcnn=True, fcn_with_norm="none", fcn_num_layers=3)
    config.dataset = AttrDict()
    config.train = AttrDict(use_horovod=False, panoptic_box_keep_fraction=0.7, fcn_with_roi_loss=False, batch_size=1,
                            rpn_individual_proposals=True)
    config.test = AttrDict(vis_mask=False, rpn_nms_thresh=0.7, rpn_pre_nms_top_n=1000, rpn_post_nms_top_n=1000,
                           rpn_min_size=0, nms_thresh=0.5, max_det=100, score_thresh=0.05, panoptic_score_thresh=0.6,
                           panoptic_stuff_area_limit=4096)

    def update_config(config_file):
        """config.py:177-198: merge an experiment yaml into `config` (nested sections key by key)."""
        import yaml
        with open(config_file) as f:
            exp_config = AttrDict(yaml.safe_load(f))
        for k, v in exp_config.items():
            if k in config and isinstance(v, dict):
                if k == "train" and "bbox_weights" in v:
                    v["bbox_weights"] = np.array(v["bbox_weights"])
                elif k == "network" and "pixel_means" in v:
                    v["pixel_means"] = np.array(v["pixel_means"])
                for vk, vv in v.items():
                    config[k][vk] = vv
            else:
                config[k] = v
        if config.debug_mode:
            config.train.use_horovod = False
            config.gpus = "0"
