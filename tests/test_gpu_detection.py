"""Fused detection-glue kernels (csrc/detection.cu) against the torch restatement of the reference's numpy glue
(upsnet_b200/detection.py with FUSED off -- the same code the CPU reference arm runs), on the GPU."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available()
    return torch.device("cuda", 0)


def _both(fn):
    from upsnet_b200 import detection as D
    outs = []
    for fused in (True, False):
        D.FUSED["on"] = fused
        try:
            outs.append(fn())
        finally:
            D.FUSED["on"] = True
    torch.cuda.synchronize()
    return outs


@pytest.mark.parametrize("seed", [0, 1])
def test_rpn_decode_fused_matches_torch_glue(dev, seed):
    from upsnet_b200.detection import StaticProposalGenerator
    g = torch.Generator(device="cpu").manual_seed(seed)
    gen = StaticProposalGenerator(pre_nms_top_n=300, post_nms_top_n=200)
    A = 3
    shapes = [(48, 80), (24, 40), (12, 20), (6, 10), (3, 5)]
    # distinct scores: the order of equal scores is unspecified in torch.topk (and in the reference's argsort)
    probs = [((torch.randperm(A * h * w, generator=g).float() + 0.5) / (A * h * w)).reshape(1, A, h, w).to(dev)
             for h, w in shapes]
    # large deltas exercise the exp clamp and the image clip
    deltas = [(torch.randn(1, 4 * A, h, w, generator=g) * (0.5 + 2.0 * (i == 1))).to(dev) for i, (h, w) in enumerate(shapes)]
    im_info = np.array([190.0, 317.0, 1.0], np.float32)
    (r1, s1, v1), (r0, s0, v0) = _both(lambda: gen(probs, deltas, im_info))
    assert torch.equal(v1, v0) and torch.equal(s1, s0)
    assert torch.equal(r1, r0)


@pytest.mark.parametrize("cfg", [
    dict(R=1000, C=9, agnostic=False, thresh=0.05, top_n=100, sharp=3.0),
    dict(R=1000, C=9, agnostic=True, thresh=0.6, top_n=100, sharp=6.0),
    dict(R=1000, C=9, agnostic=False, thresh=0.05, top_n=100, sharp=0.3),     # many candidates per roi
    dict(R=300, C=9, agnostic=True, thresh=0.2, top_n=40, sharp=4.0),         # agnostic with thresh < 0.5: M = R*(C-1)
    dict(R=1000, C=9, agnostic=False, thresh=0.999, top_n=100, sharp=1.0),    # nothing survives -> dummy detection
    dict(R=64, C=5, agnostic=False, thresh=0.05, top_n=0, sharp=2.0),         # top_n = 0: keep all
])
def test_maskroi_fused_matches_torch_glue(dev, cfg):
    from upsnet_b200.detection import StaticMaskROI
    g = torch.Generator(device="cpu").manual_seed(5)
    R, C = cfg["R"], cfg["C"]
    x1 = torch.rand(R, generator=g) * 1800
    y1 = torch.rand(R, generator=g) * 900
    w = torch.rand(R, generator=g) * 300 + 4
    h = torch.rand(R, generator=g) * 200 + 4
    # clusters of overlapping boxes so that NMS has work; some exact duplicates (score ties keep roi order)
    x1[R // 2:] = x1[:R - R // 2] + torch.randn(R - R // 2, generator=g) * 6
    y1[R // 2:] = y1[:R - R // 2] + torch.randn(R - R // 2, generator=g) * 6
    rois = torch.stack([torch.zeros(R), x1, y1, x1 + w, y1 + h], 1).to(dev)
    logits = torch.randn(R, C, generator=g) * cfg["sharp"]
    logits[R // 2:] = logits[:R - R // 2]
    cls_prob = torch.softmax(logits, 1).to(dev)
    deltas = (torch.randn(R, 4 * C, generator=g) * 0.8).to(dev)
    valid = (torch.rand(R, generator=g) > 0.1).to(dev)
    mr = StaticMaskROI(cfg["top_n"], C, 0.5, cfg["agnostic"], cfg["thresh"], (10., 10., 5., 5.))
    im_info = np.array([1024.0, 2048.0, 1.0], np.float32)
    (s1, b1, c1, n1), (s0, b0, c0, n0) = _both(lambda: mr(rois, valid, deltas, cls_prob, im_info))
    n = int(n0)
    assert int(n1) == n and n >= 1
    assert c1.dtype == c0.dtype and torch.equal(c1[:n], c0[:n])
    assert torch.equal(s1[:n], s0[:n])
    # boxes: the kernel divides the deltas by the regression weights like the reference's numpy (true division);
    # torch's CUDA `tensor / python_scalar` multiplies by the rounded reciprocal -> last-bit differences
    assert (b1[:n] - b0[:n]).abs().max().item() < 1e-3
    assert float(s1[n:].abs().sum()) == 0.0 and float(b1[n:].abs().sum()) == 0.0


@pytest.mark.parametrize("quant", [0, 64, 3])
def test_rpn_topk_matches_stable_sort(dev, quant):
    """Radix-select top-k of every level == the first k of a stable descending sort of the (y,x,a)-ordered scores,
    also with massive ties (quantised / saturated scores)."""
    from upsnet_b200 import operators as OPS
    g = torch.Generator(device="cpu").manual_seed(11 + quant)
    A = 3
    shapes = [(96, 160), (48, 80), (24, 40), (12, 20), (6, 10)]
    probs = []
    for h, w in shapes:
        pr = torch.sigmoid(torch.randn(A, h, w, generator=g) * 4)
        if quant:
            pr = torch.round(pr * quant) / quant
        probs.append(pr.to(dev))
    for pre in (1000, 300, 2048):
        sc, idx, ks = OPS.rpn_topk(probs, A, pre)
        torch.cuda.synchronize()
        o = 0
        for pr, k in zip(probs, ks):
            flat = pr.permute(1, 2, 0).reshape(-1)
            assert k == min(pre, flat.numel())
            want_s, want_i = torch.sort(flat, descending=True, stable=True)
            assert torch.equal(idx[o:o + k], want_i[:k])
            assert torch.equal(sc[o:o + k], want_s[:k])
            o += k


def test_maskroi_finish_reports_truncation(dev):
    """the fixed-size MaskROI buffers must not drop detections silently.  n_out[1] of upsnet_maskroi_finish:
    bit 1 = more boxes tied at / above the top-n threshold than output slots, bit 0 = more NMS survivors than the 4096
    candidate slots (the reference keeps them all, mask_roi.py:96-121)."""
    from upsnet_b200 import operators as ops
    def run(nseg, M, scores, top_n, cap):
        n = nseg * M
        keep = torch.arange(M, dtype=torch.int32, device=dev).repeat(nseg, 1).contiguous()
        cnt = torch.full((nseg,), M, dtype=torch.int32, device=dev)
        offs = (torch.arange(nseg + 1, dtype=torch.int32, device=dev) * M).contiguous()
        sc = scores.to(dev).float().contiguous()
        cls = torch.ones(n, dtype=torch.int32, device=dev)
        bx = torch.rand(n, 4, device=dev)
        o_sc, o_bx, o_cls, n_out, flags = ops.maskroi_finish(keep, cnt, offs, sc, cls, bx, top_n, cap)
        return int(n_out.item()), int(flags.item()), o_sc.cpu()
    # 200 boxes with the same score: the tie at the threshold (all 200) does not fit 128 slots
    n, f, sc = run(1, 200, torch.full((200,), 0.9), 100, 128)
    assert n == 128 and f == 2 and bool((sc == 0.9).all())
    # distinct scores: exactly top_n survive, nothing dropped
    n, f, sc = run(1, 200, torch.linspace(0.1, 0.9, 200), 100, 128)
    assert n == 100 and f == 0
    # 5 x 1000 survivors > 4096 candidate slots
    n, f, sc = run(5, 1000, torch.rand(5000), 100, 128)
    assert (f & 1) == 1
