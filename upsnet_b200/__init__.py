"""upsnet_b200 -- B200-native (sm_100a) implementation of the UPSNet per-image inference hot path
behind the reference's own operator API.  All compute lives in libupsnet_b200.so (C ABI:
include/upsnet_b200.h); this package is the thin PyTorch-facing host layer."""
from .operators import (DeformConv, DeformConvWithOffset, ModDeformConv, ModDeformConvWithOffsetMask,  # noqa: F401
                        ModulatedDeformConv, RoIAlign, ROIAlign, RoIAlignFunction, FPNRoIAlign, PanopticHead,
                        MaskRemoval, SegTerm, MaskTerm, MaskMatching,
                        conv2d, linear, deform_conv, roi_align, fpn_roi_align, nms, nms_segmented, gpu_nms,
                        gpu_nms_wrapper, panoptic_fuse, set_precision, unified_pan_result, prep_image, im_post, im_post_rle)

from .pipeline import PipelinedEngine  # noqa: F401,E402

__all__ = ["PipelinedEngine", "DeformConv", "DeformConvWithOffset", "ModDeformConv", "ModDeformConvWithOffsetMask",
           "ModulatedDeformConv", "RoIAlign", "ROIAlign", "RoIAlignFunction", "FPNRoIAlign", "PanopticHead",
           "MaskRemoval", "SegTerm", "MaskTerm", "MaskMatching",
           "conv2d", "linear", "deform_conv", "roi_align", "fpn_roi_align", "nms", "nms_segmented", "gpu_nms",
           "gpu_nms_wrapper", "panoptic_fuse", "set_precision", "unified_pan_result", "prep_image", "im_post", "im_post_rle"]
