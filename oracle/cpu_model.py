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


def _conv2d(x, weight, bias=None, stride=1, padding=0, dilation=1, residual=None, relu=False, precision=None,
            out_format=None, out_dtype=None, residual_up2=False, pair_group=0, sigmoid_from=None):
    y = F.conv2d(x.float(), weight.float(), None if bias is None else bias.float(), stride, padding, dilation)
    if residual is not None:
        y = y + (F.interpolate(residual, scale_factor=2, mode="nearest") if residual_up2 else residual)
    y = F.relu(y) if relu else y
    if sigmoid_from is not None:
        y = torch.cat([y[:, :sigmoid_from], torch.sigmoid(y[:, sigmoid_from:])], 1)
    return y


def _linear(x, weight, bias=None, relu=False, precision=None, out_dtype=None):
    y = F.linear(x, weight, bias)
    return F.relu(y) if relu else y


def _pair(v):
    return v if isinstance(v, (tuple, list)) else (v, v)


def _deform_conv(data, offset, weight, bias=None, stride=1, padding=0, dilation=1, deformable_groups=1, mask=None,
                 relu=False, precision=None, out_format=None, out_dtype=None):
    """Reference structure (functions/deform_conv.py:44-57): per image, deformable im2col (C oracle,
    OpenMP) into a column buffer, then torch.mm on the host cores, then bias."""
    s, p, d = _pair(stride)[0], _pair(padding)[0], _pair(dilation)[0]
    x = np.ascontiguousarray(data.detach().numpy(), np.float32)
    off = np.ascontiguousarray(offset.detach().numpy(), np.float32)
    m = None if mask is None else np.ascontiguousarray(mask.detach().numpy(), np.float32)
    N, Cin, H, W = x.shape
    Cout, _, kh, kw = weight.shape
    Ho, Wo = O.conv_out(H, p, d, kh, s), O.conv_out(W, p, d, kw, s)
    col = np.empty((Cin * kh * kw, Ho * Wo), np.float32)
    w2 = weight.detach().reshape(Cout, -1).float()
    ys = []
    for i in range(N):
        O.lib().oracle_deform_im2col(x[i], off[i], None if m is None else m[i].ctypes.data_as(O.C.c_void_p),
                                     Cin, H, W, kh, kw, p, p, s, s, d, d, deformable_groups, col)
        ys.append(torch.mm(w2, torch.from_numpy(col)).view(1, Cout, Ho, Wo))
    y = torch.cat(ys)
    if bias is not None:
        y = y + bias.detach().view(1, -1, 1, 1)
    return F.relu(y) if relu else y


def _roi_align(features, rois, ph, pw, scale, sampling_ratio=2, layout="nchw"):
    assert layout == "nchw"
    return torch.from_numpy(O.roi_align(features.detach().numpy(), rois.detach().numpy(), ph, pw, scale, sampling_ratio))


def _fpn_roi_align(feats, rois, ph, pw, scales, sampling_ratio=2, layout="nchw", return_levels=False):
    assert layout in ("nchw", "auto")
    out = torch.from_numpy(O.fpn_roi_align([f.detach().contiguous().numpy() for f in feats], rois.detach().numpy(), ph, pw, scales))
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
                   want_sem=False, n_dev=None, up4=False):
    assert not up4, "the CPU path materialises fcn_output (model._semantic fuses the up-sampling on CUDA only)"
    n = mask_rois.shape[0] if n_dev is None else max(min(int(n_dev.item()), mask_rois.shape[0]), 1)
    r = O.panoptic_head(fcn_output[0].detach().contiguous().numpy(), mask_rois[:n].detach().contiguous().numpy(),
                        cls_prob[:n].detach().numpy(), mask_logit[:n].detach().contiguous().numpy().reshape(-1, 28, 28),
                        cls_idx[:n].numpy(), num_stuff, fraction_threshold, want_sem=want_sem)
    keep = torch.from_numpy(r[0])
    out = [keep, torch.from_numpy(r[1])[None]]
    if want_sem:
        out.append(torch.from_numpy(r[2])[None])
    if n_dev is not None:      # static-shape contract: padded keep + device-style count
        pad = torch.zeros(mask_rois.shape[0], dtype=torch.int64)
        pad[:keep.numel()] = keep
        out[0] = pad
        out.append(torch.tensor([keep.numel()], dtype=torch.int32))
    return tuple(out)


@contextlib.contextmanager
def cpu_ops():
    """Run upsnet_b200.model on CPU tensors (checker / baseline only)."""
    import upsnet_b200.detection as det
    import upsnet_b200.operators as ops
    names = ["conv2d", "linear", "deform_conv", "roi_align", "fpn_roi_align", "nms_segmented", "panoptic_fuse",
             "max_pool2d", "upsample_bilinear"]
    impl = [_conv2d, _linear, _deform_conv, _roi_align, _fpn_roi_align, _nms_segmented, _panoptic_fuse,
            lambda x, k, s, p: torch.nn.functional.max_pool2d(x, k, s, p),
            lambda x, f: torch.nn.functional.interpolate(x, None, f, mode="bilinear", align_corners=False)]
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


from upsnet_b200.synthetic import synthetic_input, synthetic_model  # noqa: E402,F401  (re-export for tests)
