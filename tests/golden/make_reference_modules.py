"""Generates tests/golden/reference_modules.npz by EXECUTING the reference's own python-level operators
(build container only: /root/reference is not present on the GPU box).  Run: python tests/golden/make_reference_modules.py

Executed, unmodified, from /root/reference/upsnet:
  operators/modules/mask_removal.py:23-93      MaskRemoval.forward           (real cv2.resize)
  operators/modules/unary_logits.py:69-105     SegTerm.forward
  operators/modules/mask_roi.py:24-146         MaskROI.forward
  operators/functions/pyramid_proposal.py:23-222 + modules/pyramid_proposal.py:36-67   PyramidProposal
The panoptic glue between MaskRemoval and SegTerm is `models/resnet_upsnet.py:223-240`; that file cannot be imported
(it pulls the un-buildable CUDA extensions), so `reference_panoptic_glue` below replays those eighteen lines on the
reference modules' outputs and cites each line.

Import shims (nothing inside the executed functions is changed):
  * easydict / the Cython modules (bbox, cpu_nms, gpu_nms) are stubbed like in make_golden.py;
  * `gpu_nms_wrapper` -> the reference's own pure-python NMS (`nms/nms.py:47-86 py_nms`; same IoU rule `> thresh`
    with +1 areas as nms_kernel.cu:30-38) -- the only CPU NMS the reference ships that matches the GPU kernel's rule;
  * `Tensor.cuda()` / `.pin_memory()` / `.get_device()` / `.to(cuda)` are mapped to the CPU (there is no GPU here);
  * torch >= 1.3 refuses legacy `Function` instances being called: `PyramidProposalFunction.__call__` is pointed at its
    own `forward` (what the legacy call did for a non-differentiable forward).
"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
# NOT the repo root: its `upsnet/` import shim is a regular package and would shadow the reference's namespace package
sys.path.insert(0, "/root/reference")

if not hasattr(np, "float"):
    np.float = float
if not hasattr(np, "int"):
    np.int = int


def _install_shims():
    import torch

    class _EasyDict(dict):
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

    ed = types.ModuleType("easydict"); ed.EasyDict = _EasyDict
    sys.modules["easydict"] = ed
    for name, attrs in (("upsnet.bbox.bbox", ("bbox_overlaps",)), ("upsnet.nms.cpu_nms", ("cpu_nms", "cpu_soft_nms")),
                        ("upsnet.nms.gpu_nms", ("gpu_nms",))):
        m = types.ModuleType(name)
        for a in attrs:
            setattr(m, a, lambda *a_, **k_: (_ for _ in ()).throw(RuntimeError("Cython extension not built")))
        sys.modules[name] = m
    # no GPU in the build container: device moves are identities
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.Tensor.pin_memory = lambda self, *a, **k: self
    torch.Tensor.get_device = lambda self: 0
    _to = torch.Tensor.to

    def to(self, *a, **k):
        a = tuple(torch.device("cpu") if (isinstance(x, torch.device) and x.type == "cuda") or (isinstance(x, int) and not isinstance(x, bool))
                  else x for x in a)
        if isinstance(k.get("device"), (torch.device, int)):
            k["device"] = torch.device("cpu")
        return _to(self, *a, **k)
    torch.Tensor.to = to


def _reference():
    _install_shims()
    from upsnet.config.config import config
    config.dataset.num_classes = 9          # experiments/upsnet_resnet50_cityscapes_16gpu.yaml
    config.dataset.num_seg_classes = 19
    import upsnet.nms.nms as ref_nms
    import upsnet.operators.modules.mask_roi as ref_mask_roi
    import upsnet.operators.functions.pyramid_proposal as ref_ppf
    import upsnet.operators.modules.pyramid_proposal as ref_ppm
    from upsnet.operators.modules.mask_removal import MaskRemoval
    from upsnet.operators.modules.unary_logits import SegTerm
    ref_mask_roi.gpu_nms_wrapper = lambda thresh, device_id=None: ref_nms.py_nms_wrapper(thresh)
    ref_ppf.gpu_nms_wrapper = lambda thresh, device_id=None: ref_nms.py_nms_wrapper(thresh)
    ref_ppf.PyramidProposalFunction.__call__ = lambda self, *a: self.forward(*a)
    return config, ref_mask_roi.MaskROI, ref_ppm.PyramidProposal, MaskRemoval, SegTerm


def _reference_training_modules():
    """MaskTerm (unary_logits.py:24-66) and MaskMatching (mask_matching.py:27-62); the latter's module imports matplotlib
    (not installed, unused by MaskMatching) -- stubbed."""
    for name in ("matplotlib", "matplotlib.pyplot"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    from upsnet.operators.modules.unary_logits import MaskTerm
    from upsnet.operators.modules.mask_matching import MaskMatching
    return MaskTerm, MaskMatching


def reference_panoptic_glue(mask_removal, seg_term, fcn_output, mask_rois, cls_prob, mask_score, cls_idx, num_seg_classes,
                            num_classes):
    """models/resnet_upsnet.py:223-240 (enable_void branch, the one every shipped yaml takes) on the reference modules."""
    import torch
    keep_inds, mask_logits = mask_removal(mask_rois[:, 1:], cls_prob, mask_score, cls_idx, fcn_output.shape[2:])   # :223
    mask_rois = mask_rois[keep_inds]                                                                             # :224
    cls_idx = cls_idx[keep_inds]                                                                                 # :225
    cls_prob = cls_prob[keep_inds]                                                                               # :226
    seg_logits, seg_inst_logits = seg_term(cls_idx, fcn_output, mask_rois * 4.0)                                 # :227
    void_logits = torch.max(fcn_output[:, (num_seg_classes - num_classes + 1):, ...], dim=1, keepdim=True)[0] - \
        torch.max(seg_inst_logits, dim=1, keepdim=True)[0]                                                       # :235
    inst_logits = (seg_inst_logits + mask_logits)                                                                # :236
    panoptic_logits = torch.cat([seg_logits, inst_logits, void_logits], dim=1)                                   # :237
    void_id = panoptic_logits.shape[1] - 1                                                                       # :238
    panoptic_output = torch.max(panoptic_logits, dim=1)[1]                                                       # :239
    panoptic_output[panoptic_output == void_id] = 255                                                            # :240
    return keep_inds, mask_logits, seg_inst_logits, cls_idx, cls_prob, panoptic_output


def rand_boxes(rng, n, H, W, smin, smax):
    c = np.stack([rng.uniform(0, W, n), rng.uniform(0, H, n)], 1)
    s = np.exp(rng.uniform(np.log(smin), np.log(smax), (n, 2)))
    b = np.concatenate([c - s / 2, c + s / 2], 1).astype(np.float32)
    b[:, 0::2] = np.clip(b[:, 0::2], 0, W - 1); b[:, 1::2] = np.clip(b[:, 1::2], 0, H - 1)
    return b


def main():
    import torch
    config, MaskROI, PyramidProposal, MaskRemoval, SegTerm = _reference()
    rng = np.random.default_rng(20260924)
    out = {}

    # ------------------------------------------------------------------ panoptic head (a14, a15, a16)
    pan_cases = [dict(H=96, W=160, n=24, smin=6, smax=70), dict(H=64, W=96, n=7, smin=8, smax=60),
                 dict(H=128, W=256, n=60, smin=5, smax=120), dict(H=48, W=64, n=1, smin=10, smax=30, dummy=True),
                 dict(H=80, W=112, n=12, smin=20, smax=80, same_class=True)]
    for ci, c in enumerate(pan_cases):
        H, W, n = c["H"], c["W"], c["n"]
        fcn = (rng.standard_normal((1, 19, H, W)) * 3).astype(np.float32)
        bx = rand_boxes(rng, n, H, W, c["smin"], c["smax"])
        rois = np.concatenate([np.zeros((n, 1), np.float32), bx], 1)
        prob = (0.6 + 0.4 * (rng.permutation(n) + 1) / (n + 1)).astype(np.float32)           # distinct scores
        ml = (rng.standard_normal((n, 1, 28, 28)) * 2 + 0.5).astype(np.float32)
        cls = rng.integers(1, 9, n).astype(np.int64)
        if c.get("same_class"):
            cls[:] = 3                                                                       # heavy overlap pruning
        if c.get("dummy"):                                                                   # mask_roi.py:132-139 dummy detection
            rois[:] = 0; prob[:] = 1; cls[:] = 0
        mr, st = MaskRemoval(0.3), SegTerm(19)
        keep, mlog, sinst, kcls, kprob, pan = reference_panoptic_glue(
            mr, st, torch.from_numpy(fcn), torch.from_numpy(rois), torch.from_numpy(prob), torch.from_numpy(ml),
            torch.from_numpy(cls), 19, 9)
        p = "pan%d_" % ci
        out.update({p + "fcn": fcn, p + "rois": rois, p + "prob": prob, p + "mask_score": ml, p + "cls": cls,
                    p + "keep": keep.numpy().astype(np.int64), p + "panoptic": pan.numpy().astype(np.int64),
                    p + "kept_cls": kcls.numpy().astype(np.int64),
                    # energy planes are large: keep a checksum pair per plane (sum, sum of squares in float64)
                    p + "mask_energy_sum": mlog.double().sum(dim=(0, 2, 3)).numpy(),
                    p + "mask_energy_cnt": (mlog != 0).sum(dim=(0, 2, 3)).numpy().astype(np.int64),
                    p + "seg_inst_sum": sinst.double().sum(dim=(0, 2, 3)).numpy(),
                    p + "seg_inst_cnt": (sinst != 0).sum(dim=(0, 2, 3)).numpy().astype(np.int64)})
    out["pan_cases"] = np.int64(len(pan_cases))

    # ------------------------------------------------------------------ MaskROI (a10)
    im_info = np.array([[512, 768, 1.0]], np.float32)
    mr_cases = [dict(R=300, agnostic=False, score=0.05, name="det"), dict(R=300, agnostic=True, score=0.6, name="pan"),
                dict(R=400, agnostic=False, score=0.05, name="tie", tie=True), dict(R=50, agnostic=False, score=0.05, name="none", none=True),
                dict(R=50, agnostic=True, score=0.6, name="none_pan", none=True), dict(R=40, agnostic=False, score=0.05, name="few")]
    for ci, c in enumerate(mr_cases):
        R, C = c["R"], 9
        rois = np.concatenate([np.zeros((R, 1), np.float32), rand_boxes(rng, R, 512, 768, 16, 300)], 1)
        delta = (rng.standard_normal((R, 4 * C)) * 0.5).astype(np.float32)
        logits = rng.standard_normal((R, C)) * (3.0 if not (c.get("few") or c.get("tie")) else 1.0)
        prob = torch.softmax(torch.from_numpy(logits).float(), 1).numpy()
        if c.get("none"):
            prob = np.full((R, C), 0.01, np.float32); prob[:, 0] = 0.92
        if c.get("tie"):
            # many candidates share the score that ends up being the max_det-th largest (mask_roi.py:110-113 keeps them all)
            hot = rng.choice(R, 160, replace=False)[:90]
            prob[hot, 1 + (hot % 8)] = np.float32(0.5)
            rois[hot, 1:] = rand_boxes(rng, 90, 512, 768, 10, 20)      # small far-apart boxes: survive NMS
            delta[hot] = 0
        m = MaskROI(clip_boxes=True, bbox_class_agnostic=False, top_n=config.test.max_det, num_classes=C,
                    nms_thresh=0.5, class_agnostic=c["agnostic"], score_thresh=c["score"])
        sc, bxs, ci_ = m(torch.from_numpy(rois), torch.from_numpy(delta), torch.from_numpy(prob), im_info)
        p = "mroi%d_" % ci
        out.update({p + "rois": rois, p + "delta": delta, p + "prob": prob.astype(np.float32), p + "agnostic": np.bool_(c["agnostic"]),
                    p + "score_thresh": np.float32(c["score"]), p + "out_scores": sc.numpy().astype(np.float32),
                    p + "out_boxes": bxs.numpy().astype(np.float32), p + "out_cls": ci_.numpy().astype(np.int64)})
    out["mroi_cases"] = np.int64(len(mr_cases)); out["mroi_im_info"] = im_info

    # ------------------------------------------------------------------ PyramidProposal (a6)
    pp_cases = [dict(H=128, W=192, pre=1000, post=1000), dict(H=256, W=384, pre=1000, post=1000), dict(H=192, W=320, pre=200, post=300)]
    strides = (4, 8, 16, 32, 64)
    for ci, c in enumerate(pp_cases):
        H, W = c["H"], c["W"]
        info = np.array([[H, W, 1.0]], np.float32)
        probs, deltas = [], []
        for s in strides:
            h, w = -(-H // s), -(-W // s)
            # distinct scores per level (np.argpartition + argsort order is unspecified on ties)
            sc = rng.permutation(3 * h * w).astype(np.float64).reshape(1, 3, h, w)
            probs.append(((sc + 1) / (3 * h * w + 1)).astype(np.float32))
            deltas.append((rng.standard_normal((1, 12, h, w)) * 0.4).astype(np.float32))
        m = PyramidProposal(feat_stride=np.array(strides), scales=np.array((8,)), ratios=np.array((0.5, 1, 2)),
                            rpn_pre_nms_top_n=c["pre"], rpn_post_nms_top_n=c["post"], threshold=0.7, rpn_min_size=0,
                            individual_proposals=True)
        rois, sc = m([torch.from_numpy(p_) for p_ in probs], [torch.from_numpy(d_) for d_ in deltas], info)
        p = "pp%d_" % ci
        for l in range(5):
            out[p + "prob%d" % l] = probs[l]; out[p + "delta%d" % l] = deltas[l]
        out.update({p + "im_info": info, p + "pre": np.int64(c["pre"]), p + "post": np.int64(c["post"]),
                    p + "rois": rois.numpy().astype(np.float32), p + "scores": sc.numpy().astype(np.float32)})
    out["pp_cases"] = np.int64(len(pp_cases))

    # ------------------------------------------------------------------ MaskTerm / MaskMatching (a17: training twins)
    MaskTerm, MaskMatching = _reference_training_modules()
    H4, W4, n = 48, 80, 9
    masks = (rng.standard_normal((n, 1, 28, 28)) * 2).astype(np.float32)
    bx = rand_boxes(rng, n, 4 * H4, 4 * W4, 20, 200)
    bx[0] = [-13.0, 5.0, 60.0, 90.0]            # partly outside (negative corner truncates toward zero in .long())
    rois = np.concatenate([np.zeros((n, 1), np.float32), bx], 1)
    cls = rng.integers(1, 9, n).astype(np.int64)
    seg = (rng.standard_normal((1, 19, H4, W4))).astype(np.float32)
    mt = MaskTerm(19, box_scale=1 / 4.0)
    energy = mt(torch.from_numpy(masks), torch.from_numpy(rois), torch.from_numpy(cls), torch.from_numpy(seg))
    out.update(mterm_masks=masks, mterm_rois=rois, mterm_cls=cls, mterm_seg_shape=np.array(seg.shape), mterm_energy=energy.numpy())
    gt_segs = rng.integers(0, 19, (1, H4, W4)).astype(np.int64)
    gt_segs[0, :4] = 255
    gt_masks = np.zeros((5, H4, W4), np.int64)
    for i in range(5):
        y, x = rng.integers(0, H4 - 12), rng.integers(0, W4 - 16)
        gt_masks[i, y:y + 12, x:x + 16] = 1
    gt_masks[2, 0:3, 0:3] = 255
    mm = MaskMatching(19, enable_void=True)
    m_all = mm(torch.from_numpy(gt_segs), torch.from_numpy(gt_masks))
    keep = np.array([3, 0, 4], np.int64)
    m_keep = mm(torch.from_numpy(gt_segs), torch.from_numpy(gt_masks), torch.from_numpy(keep))
    out.update(mmatch_gt_segs=gt_segs, mmatch_gt_masks=gt_masks, mmatch_keep=keep, mmatch_all=m_all.numpy(), mmatch_kept=m_keep.numpy())

    # ------------------------------------------------------------------ get_unified_pan_result (f2)
    # dataset/base_dataset.py pulls pycocotools / Cython at import: the METHOD's own source lines (332-371) are compiled
    # from the file, unmodified, with `np` and the reference `config` as its globals.
    import ast
    src = open("/root/reference/upsnet/dataset/base_dataset.py").read()
    fn = [n for n in ast.walk(ast.parse(src)) if isinstance(n, ast.FunctionDef) and n.name == "get_unified_pan_result"][0]
    ns = {"np": np, "config": config}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), "base_dataset.py:332-371", "exec"), ns)
    get_unified = ns["get_unified_pan_result"]
    uni_cases = [dict(H=96, W=160, k=14, limit=400), dict(H=128, W=192, k=30, limit=4 * 64 * 64), dict(H=64, W=64, k=0, limit=100)]
    for ci, c in enumerate(uni_cases):
        H, W, k = c["H"], c["W"], c["k"]
        seg = rng.integers(0, 19, (H, W)).astype(np.int64)
        # blocky semantic map so that majorities exist
        seg = np.kron(rng.integers(0, 19, (H // 16, W // 16)), np.ones((16, 16), np.int64)).astype(np.int64)
        pan = np.kron(rng.integers(0, 11, (H // 8, W // 8)), np.ones((8, 8), np.int64)).astype(np.int64)
        cls = rng.integers(1, 9, max(k, 1)).astype(np.int64)
        for j in range(k):
            y, x = rng.integers(0, H - 24), rng.integers(0, W - 24)
            hh, ww = rng.integers(6, 24), rng.integers(6, 24)
            pan[y:y + hh, x:x + ww] = 11 + j
            if j % 3 == 0:      # make the instance's own class the semantic majority under it
                seg[y:y + hh, x:x + ww] = cls[j] + 10
            elif j % 3 == 1:    # a stuff class holds the majority: the segment is re-labelled as stuff
                seg[y:y + hh, x:x + ww] = rng.integers(0, 11)
        pan[:3, :7] = 255
        res = get_unified(None, [seg], [pan], [cls], stuff_area_limit=c["limit"])[0]
        p = "uni%d_" % ci
        out.update({p + "seg": seg, p + "pan": pan, p + "cls": cls[:k] if k else np.zeros((0,), np.int64), p + "limit": np.int64(c["limit"]),
                    p + "out": res})
    out["uni_cases"] = np.int64(len(uni_cases))

    # ------------------------------------------------------------------ input pipeline (f3): prep_im_for_blob + im_list_to_blob
    fnames = ("prep_im_for_blob", "im_list_to_blob")
    fns = [n for n in ast.walk(ast.parse(src)) if isinstance(n, ast.FunctionDef) and n.name in fnames]
    import cv2
    ns2 = {"np": np, "config": config, "cv2": cv2}
    exec(compile(ast.Module(body=fns, type_ignores=[]), "base_dataset.py:143-174,898-923", "exec"), ns2)
    config.network.has_fpn = True
    for ci, (h, w, target, max_size) in enumerate(((120, 200, 120, 400), (100, 150, 160, 1333), (97, 131, 80, 100))):
        im = rng.integers(0, 256, (h, w, 3)).astype(np.uint8)
        ims, scales = ns2["prep_im_for_blob"](None, im.copy(), config.network.pixel_means, [target], max_size)
        blob = ns2["im_list_to_blob"](None, [ims[0].transpose(2, 0, 1)])
        p = "prep%d_" % ci
        out.update({p + "im": im, p + "scale": np.float64(scales[0]), p + "resized_hw": np.array(ims[0].shape[:2]), p + "blob": blob})
    out["prep_cases"] = np.int64(3)
    out["prep_pixel_means"] = np.asarray(config.network.pixel_means, np.float64)

    np.savez_compressed(os.path.join(HERE, "reference_modules.npz"), **out)
    print("wrote reference_modules.npz with", len(out), "arrays;",
          "panoptic kept:", [int(out["pan%d_keep" % i].shape[0]) for i in range(len(pan_cases))],
          "maskroi out:", [int(out["mroi%d_out_scores" % i].shape[0]) for i in range(len(mr_cases))],
          "proposals:", [int(out["pp%d_rois" % i].shape[0]) for i in range(len(pp_cases))])


if __name__ == "__main__":
    main()
