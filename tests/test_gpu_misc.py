"""Small fused kernels that replaced torch eager ops inside the captured step."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available()
    return torch.device("cuda", 0)


def test_fcn_score_fuse_vs_torch(dev):
    import upsnet_b200 as U
    from upsnet_b200 import operators as ops
    g = torch.Generator(device="cpu").manual_seed(0)
    for (H, W) in ((64, 96), (256, 512), (200, 336)):
        s = [torch.randn(1, 19, H >> l, W >> l, generator=g).to(dev) for l in range(4)]
        want = s[0]
        for l in range(1, 4):
            want = want + F.interpolate(s[l], None, 2 ** l, mode="bilinear", align_corners=False)
        got = ops.fcn_score_fuse(*s)
        assert (got - want).abs().max().item() < 1e-5


@pytest.mark.parametrize("prec", ["bf16x3", "bf16", "fp32"])
def test_rpn_head_sigmoid_epilogue(dev, prec):
    """[cls | bbox | sigmoid(cls)] from one launch == separate convs + torch.sigmoid (models/rpn.py:52-56)."""
    import upsnet_b200 as U
    from upsnet_b200.model import RPN
    torch.manual_seed(1)
    rpn = RPN(3, 256).to(dev)
    rpn.cls_score.weight.data.normal_(0, 0.05); rpn.bbox_pred.weight.data.normal_(0, 0.02)
    rpn.cls_score.bias.data.normal_(0, 0.5)
    rpn.prepare()
    x = torch.randn(1, 256, 40, 56, device=dev)
    t_ = F.relu(F.conv2d(x, rpn.conv_proposal[0].weight, rpn.conv_proposal[0].bias, padding=1))
    want_cls = F.conv2d(t_, rpn.cls_score.weight, rpn.cls_score.bias)
    want_box = F.conv2d(t_, rpn.bbox_pred.weight, rpn.bbox_pred.bias)
    U.set_precision(prec)
    try:
        with torch.no_grad():
            cls, box, prob = rpn(x)
    finally:
        U.set_precision("fp32")
    tol = {"bf16x3": 1e-3, "bf16": 5e-2, "fp32": 1e-3}[prec]
    assert cls.shape == want_cls.shape and box.shape == want_box.shape and prob.shape == want_cls.shape
    assert (cls - want_cls).abs().max().item() < tol and (box - want_box).abs().max().item() < tol
    assert (prob - torch.sigmoid(cls)).abs().max().item() < 2e-7       # the epilogue's sigmoid of ITS logits
    assert box.is_contiguous() or box.shape[0] == 1
