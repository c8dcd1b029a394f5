"""Generates tests/golden/*.npz by importing the REFERENCE's own python (build container only:
/root/reference is not present on the GPU box).  Run: python tests/golden/make_golden.py

What can be imported from the reference as-is: the pure-numpy NMS
(upsnet/nms/py_cpu_nms.py), bbox_transform / clip_boxes (upsnet/bbox/bbox_transform.py, with the
un-buildable Cython module it imports stubbed out) and generate_anchors (np.float alias needed).
The custom CUDA ops have no importable CPU path; their goldens come from the C oracle after it
has been cross-checked against torchvision here (recorded in the npz as *_tv_maxdiff).
"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

# --- shims so that the reference modules import on numpy 2.x without its Cython extensions ---
if not hasattr(np, "float"):
    np.float = float
if not hasattr(np, "int"):
    np.int = int
_stub = types.ModuleType("upsnet.bbox.bbox")
_stub.bbox_overlaps = lambda *a, **k: None
sys.modules["upsnet.bbox.bbox"] = _stub
class _EasyDict(dict):  # easydict is not installed; config.py only needs attribute access
    def __init__(self, d=None, **kw):
        super().__init__()
        for k, v in dict(d or {}, **kw).items():
            self[k] = v
    def __setattr__(self, k, v):
        self[k] = v
    def __setitem__(self, k, v):
        super().__setitem__(k, _EasyDict(v) if isinstance(v, dict) and not isinstance(v, _EasyDict) else v)
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)
_ed = types.ModuleType("easydict")
_ed.EasyDict = _EasyDict
sys.modules["easydict"] = _ed
from upsnet.nms.py_cpu_nms import py_cpu_nms  # noqa: E402
from upsnet.bbox.bbox_transform import bbox_transform, clip_boxes  # noqa: E402
from upsnet.rpn.generate_anchors import generate_anchors  # noqa: E402


def rand_dets(rng, n, extent, smin, smax):
    c = rng.uniform(0, extent, (n, 2))
    s = np.exp(rng.uniform(np.log(smin), np.log(smax), (n, 2)))
    scores = (rng.permutation(n).astype(np.float64) + 1) / (n + 1)  # distinct: no tie hazard
    return np.concatenate([c - s / 2, c + s / 2, scores[:, None]], 1).astype(np.float32)


def main():
    rng = np.random.default_rng(20260923)
    out = {}
    # ---- NMS: reference py_cpu_nms (IoU > thresh suppresses, +1 areas) ----
    cases = [(1, 100, 10, 50, 0.5), (63, 200, 16, 128, 0.5), (64, 200, 16, 128, 0.7), (65, 150, 16, 128, 0.3),
             (300, 300, 16, 200, 0.7), (1000, 600, 16, 300, 0.7), (2500, 700, 16, 300, 0.5)]
    for idx, (n, extent, smin, smax, thr) in enumerate(cases):
        d = rand_dets(rng, n, extent, smin, smax)
        keep = np.array([int(i) for i in py_cpu_nms(d, thr)], np.int64)
        out["nms%d_dets" % idx] = d
        out["nms%d_thresh" % idx] = np.float32(thr)
        out["nms%d_keep" % idx] = keep
    out["nms_cases"] = np.int64(len(cases))
    # ---- bbox_transform / clip_boxes ----
    boxes = rand_dets(rng, 200, 1000, 8, 400)[:, :4]
    deltas = (rng.standard_normal((200, 36)) * 0.5).astype(np.float32)
    deltas[0, 2] = 50.0  # exercises the log(1000/16) clamp
    pred = bbox_transform(boxes, deltas, (10., 10., 5., 5.))
    out["bt_boxes"], out["bt_deltas"], out["bt_pred"] = boxes, deltas, pred
    out["bt_clipped"] = clip_boxes(pred.copy(), (600, 900))
    pred1 = bbox_transform(boxes, deltas[:, :4])
    out["bt_pred_w1"] = pred1
    # ---- anchors (rpn/generate_anchors.py:50-76), float64 as the reference builds them ----
    for s in (4, 8, 16, 32, 64):
        out["anchors_%d" % s] = generate_anchors(stride=s, sizes=np.array((8,)) * s, aspect_ratios=(0.5, 1, 2))
    np.savez_compressed(os.path.join(HERE, "reference_numpy.npz"), **out)
    print("wrote reference_numpy.npz with", len(out), "arrays")

    # ---- C-oracle goldens for the CUDA-only ops, cross-checked against torchvision here ----
    import torch
    import torchvision
    from oracle import oracle as O
    g = {}
    feat = rng.standard_normal((2, 8, 40, 56)).astype(np.float32)
    c = rng.uniform(0, 224, (24, 2)); s = np.exp(rng.uniform(np.log(4), np.log(300), (24, 2)))
    rois = np.concatenate([rng.integers(0, 2, (24, 1)), c - s / 2, c + s / 2], 1).astype(np.float32)
    rois[0, 1:] = [-30, -20, 10, 12]       # partly outside
    rois[1, 1:] = [50, 60, 50, 60]         # degenerate (forced 1x1)
    rois[2, 1:] = [200, 140, 400, 300]     # beyond the map
    ra = O.roi_align(feat, rois, 7, 7, 0.25)
    tv = torchvision.ops.roi_align(torch.from_numpy(feat), torch.from_numpy(rois), (7, 7), 0.25, 2, False).numpy()
    g["ra_feat"], g["ra_rois"], g["ra_out"], g["ra_tv_maxdiff"] = feat, rois, ra, np.float32(np.abs(ra - tv).max())
    x = rng.standard_normal((2, 8, 14, 18)).astype(np.float32)
    w = (rng.standard_normal((12, 8, 3, 3)) / np.sqrt(72)).astype(np.float32)
    b = rng.standard_normal(12).astype(np.float32)
    off = (rng.standard_normal((2, 2 * 18, 14, 18)) * 2).astype(np.float32)
    y = O.deform_conv(x, off, w, b, pad=1, dg=2)
    tv = torchvision.ops.deform_conv2d(torch.from_numpy(x), torch.from_numpy(off), torch.from_numpy(w),
                                       torch.from_numpy(b), padding=1).numpy()
    g["dcn_x"], g["dcn_w"], g["dcn_b"], g["dcn_off"], g["dcn_y"] = x, w, b, off, y
    g["dcn_tv_maxdiff"] = np.float32(np.abs(y - tv).max())
    om = rng.standard_normal((2, 27, 14, 18)).astype(np.float32)
    y2 = O.mod_deform_conv(x, om, w, b, pad=1, dg=1)
    o1, o2, m = np.split(om, 3, 1)
    tv = torchvision.ops.deform_conv2d(torch.from_numpy(x), torch.from_numpy(np.concatenate([o1, o2], 1)),
                                       torch.from_numpy(w), torch.from_numpy(b), padding=1,
                                       mask=2 * torch.sigmoid(torch.from_numpy(m))).numpy()
    g["dcn2_om"], g["dcn2_y"], g["dcn2_tv_maxdiff"] = om, y2, np.float32(np.abs(y2 - tv).max())
    # panoptic: fused C oracle pinned by the literal numpy restatement AND by real cv2.resize
    H, W, n = 96, 160, 24
    fcn = (rng.standard_normal((19, H, W)) * 3).astype(np.float32)
    cc = np.stack([rng.uniform(0, W, n), rng.uniform(0, H, n)], 1)
    ss = np.exp(rng.uniform(np.log(6), np.log(70), (n, 2)))
    bx = np.concatenate([cc - ss / 2, cc + ss / 2], 1).astype(np.float32)
    bx[:, 0::2] = np.clip(bx[:, 0::2], 0, W - 1); bx[:, 1::2] = np.clip(bx[:, 1::2], 0, H - 1)
    prob = (0.6 + 0.4 * (rng.permutation(n) + 1) / (n + 1)).astype(np.float32)
    ml = (rng.standard_normal((n, 28, 28)) * 2 + 0.5).astype(np.float32)
    cls = rng.integers(1, 9, n).astype(np.int64)
    k1, l1 = O.panoptic_head(fcn, bx, prob, ml, cls, 11)
    k2, l2 = O.panoptic_head_literal(fcn, bx, prob, ml, cls, 11)
    k3, l3 = O.panoptic_head_literal(fcn, bx, prob, ml, cls, 11, resize="cv2")
    assert np.array_equal(k1, k2) and np.array_equal(l1, l2)
    g.update(pan_fcn=fcn, pan_boxes=bx, pan_prob=prob, pan_ml=ml, pan_cls=cls, pan_keep=k1, pan_labels=l1,
             pan_cv2_label_diff=np.int64((l1 != l3).sum()), pan_cv2_keep_equal=np.bool_(np.array_equal(k1, k3)))
    np.savez_compressed(os.path.join(HERE, "oracle_ops.npz"), **g)
    print("wrote oracle_ops.npz; tv maxdiffs:", g["ra_tv_maxdiff"], g["dcn_tv_maxdiff"], g["dcn2_tv_maxdiff"],
          "cv2 label diff:", g["pan_cv2_label_diff"])


if __name__ == "__main__":
    main()
