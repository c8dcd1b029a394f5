"""CPU execution of the whole resnet_upsnet forward = the "reference CPU path" of BASELINE.md
section 3 (iv): torch.nn fp32 on the host cores with the restated custom ops plugged in.

TEST INFRASTRUCTURE ONLY (tests/, smoke(), bench.py cpu_baseline / --impl reference).  It works by
temporarily substituting the CUDA entry points of upsnet_b200.operators with CPU implementations
(torch CPU convolutions = the reference's own dense ops; C oracle for ROIAlign / DCN / NMS /
panoptic head), so the *host logic* of upsnet_b200.model / detection is exercised unchanged.
The product never imports this module.
"""
import contextlib

import numpy as np
import torch
import torch.nn.functional as F

from . import oracle as O


def _conv2d(x, weight, bias=None, stride=1, padding=0, dilation=1, residual=None, relu=False, precision=None):
    y = F.conv2d(x.float(), weight.float(), None if bias is None else bias.float(), stride, padding, dilation)
    if residual is not None:
        y = y + residual
    return F.relu(y) if relu else y


def _linear(x, weight, bias=None, relu=False, precision=None):
    y = F.linear(x, weight, bias)
    return F.relu(y) if relu else y


def _pair(v):
    return v if isinstance(v, (tuple, list)) else (v, v)


def _deform_conv(data, offset, weight, bias=None, stride=1, padding=0, dilation=1, deformable_groups=1, mask=None,
                 relu=False, precision=None):
    s, p, d = _pair(stride)[0], _pair(padding)[0], _pair(dilation)[0]
    y = O.deform_conv(data.detach().numpy(), offset.detach().numpy(), weight.detach().numpy(),
                      None if bias is None else bias.detach().numpy(),
                      None if mask is None else mask.detach().numpy(), s, p, d, deformable_groups)
    y = torch.from_numpy(y)
    return F.relu(y) if relu else y


def _roi_align(features, rois, ph, pw, scale, sampling_ratio=2, layout="nchw"):
    assert layout == "nchw"
    return torch.from_numpy(O.roi_align(features.detach().numpy(), rois.detach().numpy(), ph, pw, scale, sampling_ratio))


def _fpn_roi_align(feats, rois, ph, pw, scales, sampling_ratio=2, layout="nchw", return_levels=False):
    assert layout == "nchw"
    out = torch.from_numpy(O.fpn_roi_align([f.detach().numpy() for f in feats], rois.detach().numpy(), ph, pw, scales))
    if return_levels:
        return out, torch.from_numpy(O.fpn_level_numpy(rois.detach().numpy()))
    return out


def _nms_segmented(boxes_sorted, seg_offsets, max_seg_len, thresh):
    b = boxes_sorted.detach().numpy()
    offs = seg_offsets.numpy()
    S = len(offs) - 1
    keep = torch.zeros((S, max_seg_len), dtype=torch.int32)
    cnt = torch.zeros((S,), dtype=torch.int32)
    for s in range(S):
        seg = b[offs[s]:offs[s + 1]]
        n = seg.shape[0]
        if n == 0:
            continue
        d = np.concatenate([seg, (np.arange(n, 0, -1, dtype=np.float32) / n)[:, None]], 1)  # already sorted
        k = O.nms(d, thresh)
        keep[s, :len(k)] = torch.tensor(k, dtype=torch.int32)
        cnt[s] = len(k)
    return keep, cnt


def _panoptic_fuse(fcn_output, mask_rois, cls_prob, mask_logit, cls_idx, num_stuff, fraction_threshold=0.3,
                   want_sem=False):
    r = O.panoptic_head(fcn_output[0].detach().numpy(), mask_rois.detach().numpy(), cls_prob.detach().numpy(),
                        mask_logit.detach().numpy().reshape(-1, 28, 28), cls_idx.numpy(), num_stuff,
                        fraction_threshold, want_sem=want_sem)
    out = (torch.from_numpy(r[0]), torch.from_numpy(r[1])[None])
    if want_sem:
        out = out + (torch.from_numpy(r[2])[None],)
    return out


@contextlib.contextmanager
def cpu_ops():
    """Run upsnet_b200.model on CPU tensors (checker / baseline only)."""
    import upsnet_b200.detection as det
    import upsnet_b200.operators as ops
    names = ["conv2d", "linear", "deform_conv", "roi_align", "fpn_roi_align", "nms_segmented", "panoptic_fuse"]
    impl = [_conv2d, _linear, _deform_conv, _roi_align, _fpn_roi_align, _nms_segmented, _panoptic_fuse]
    saved = {n: getattr(ops, n) for n in names}
    saved_det = det.nms_segmented
    try:
        for n, f in zip(names, impl):
            setattr(ops, n, f)
        det.nms_segmented = _nms_segmented
        yield
    finally:
        for n, f in saved.items():
            setattr(ops, n, f)
        det.nms_segmented = saved_det


def synthetic_model(cfg=None, depth=(3, 4, 6, 3), seed=0, device="cpu"):
    """Random-init resnet_upsnet per SURVEY.md section 8d config 2: reference initialisers, except
    (a) offset convs get N(0, 0.5)-scaled weights so DCN offsets are non-zero (the reference zero-inits
    them, modules/deform_conv.py:72-73 -- zero offsets would hide DCN bugs), (b) BN statistics are
    randomised so folding is exercised, (c) the class / mask heads are biased so that several dozen
    detections survive score > 0.6 and reach the panoptic head."""
    from upsnet_b200.model import resnet_upsnet, Bottleneck
    from upsnet_b200.operators import DeformConvWithOffset
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        m = resnet_upsnet(list(depth), cfg)
        for mod in m.modules():
            if isinstance(mod, torch.nn.BatchNorm2d):
                mod.weight.copy_(torch.empty_like(mod.weight).uniform_(0.5, 1.5, generator=g))
                mod.bias.copy_(torch.empty_like(mod.bias).normal_(0, 0.1, generator=g))
                mod.running_mean.copy_(torch.empty_like(mod.running_mean).normal_(0, 0.1, generator=g))
                mod.running_var.copy_(torch.empty_like(mod.running_var).uniform_(0.5, 1.5, generator=g))
            elif isinstance(mod, DeformConvWithOffset):
                w = mod.conv_offset.weight
                w.copy_(torch.empty(w.shape).normal_(0, 0.5 / (w.shape[1] * 9) ** 0.5, generator=g).to(w.device))
                cw = mod.conv.weight
                cw.copy_(torch.empty(cw.shape).normal_(0, (2.0 / (cw.shape[1] * 9)) ** 0.5, generator=g).to(cw.device))
                mod.conv.bias.zero_()
            elif isinstance(mod, Bottleneck):
                for c in (mod.conv1, mod.conv2, mod.conv3):
                    w = c.weight
                    fan = w.shape[1] * w.shape[2] * w.shape[3]
                    w.copy_(torch.empty(w.shape).normal_(0, (2.0 / fan) ** 0.5, generator=g).to(w.device))
                if mod.deformable:
                    w = mod.conv2_offset.weight
                    w.copy_(torch.empty(w.shape).normal_(0, 0.5 / (w.shape[1] * 9) ** 0.5, generator=g))
                if mod.downsample is not None:
                    w = mod.downsample[0].weight
                    w.copy_(torch.empty(w.shape).normal_(0, (1.0 / w.shape[1]) ** 0.5, generator=g))
                mod.bn3.weight.mul_(0.3)  # keep the residual stream bounded at random init
        sw = m.resnet_backbone.conv1.conv1.weight
        sw.copy_(torch.empty(sw.shape).normal_(0, (2.0 / 147) ** 0.5 / 50, generator=g))
        # heads: make detections plentiful and confident enough for the panoptic branch
        m.rpn.cls_score.weight.copy_(torch.empty_like(m.rpn.cls_score.weight).normal_(0, 0.05, generator=g))
        m.rpn.bbox_pred.weight.copy_(torch.empty_like(m.rpn.bbox_pred.weight).normal_(0, 0.01, generator=g))
        m.rcnn.cls_score.weight.copy_(torch.empty_like(m.rcnn.cls_score.weight).normal_(0, 0.25, generator=g))
        m.rcnn.cls_score.bias[0] = -1.0
        m.rcnn.bbox_pred.weight.copy_(torch.empty_like(m.rcnn.bbox_pred.weight).normal_(0, 0.02, generator=g))
        m.mask_branch.mask_score.bias.fill_(0.2)
        m.fcn_head.score.weight.copy_(torch.empty_like(m.fcn_head.score.weight).normal_(0, 0.1, generator=g))
    m = m.to(device)
    m.prepare()
    return m


def synthetic_input(H=1024, W=2048, seed=0, device="cpu"):
    g = torch.Generator().manual_seed(seed)
    data = (torch.randn(1, 3, H, W, generator=g) * 50).to(device)   # mean-subtracted BGR scale
    return {"data": data, "im_info": np.array([[H, W, 1.0]], np.float32)}
