"""Backward kernels of the custom operators (training configuration, rows g1 / K2 K3 K5 K6 K8) against torchvision's CPU
autograd of the same operators (independent implementation, same MSRA / Caffe2 lineage as the reference's kernels)."""
import numpy as np
import pytest
import torch
import torchvision

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available()
    return torch.device("cuda", 0)


@pytest.mark.parametrize("modulated", [False, True])
@pytest.mark.parametrize("cfg", [dict(N=2, Cin=8, Cout=12, H=14, W=18, stride=1, pad=1, dil=1),
                                 dict(N=1, Cin=64, Cout=32, H=20, W=24, stride=1, pad=1, dil=1),
                                 dict(N=1, Cin=16, Cout=16, H=17, W=19, stride=2, pad=2, dil=2)])
def test_deform_conv_backward_vs_torchvision(dev, cfg, modulated):
    import upsnet_b200 as U
    g = torch.Generator().manual_seed(3)
    N, Cin, Cout, H, W = cfg["N"], cfg["Cin"], cfg["Cout"], cfg["H"], cfg["W"]
    Ho = (H + 2 * cfg["pad"] - (cfg["dil"] * 2 + 1)) // cfg["stride"] + 1
    Wo = (W + 2 * cfg["pad"] - (cfg["dil"] * 2 + 1)) // cfg["stride"] + 1
    x = torch.randn(N, Cin, H, W, generator=g)
    om_ch = 27 if modulated else 18
    om = torch.randn(N, om_ch, Ho, Wo, generator=g) * 1.5
    mod = (U.ModDeformConv if modulated else U.DeformConv)(Cin, Cout, 3, stride=cfg["stride"], padding=cfg["pad"], dilation=cfg["dil"]).to(dev)
    gy = torch.randn(N, Cout, Ho, Wo, generator=g)
    xd, omd = x.to(dev).requires_grad_(True), om.to(dev).requires_grad_(True)
    y = mod(xd, omd)
    y.backward(gy.to(dev))
    # reference: torchvision CPU
    xc, omc = x.clone().requires_grad_(True), om.clone().requires_grad_(True)
    w = mod.weight.detach().cpu().clone().requires_grad_(True)
    b = mod.bias.detach().cpu().clone().requires_grad_(True)
    if modulated:
        o1, o2, m = torch.chunk(omc, 3, dim=1)
        yr = torchvision.ops.deform_conv2d(xc, torch.cat((o1, o2), 1), w, b, stride=cfg["stride"], padding=cfg["pad"],
                                           dilation=cfg["dil"], mask=torch.sigmoid(m) * 2)
    else:
        yr = torchvision.ops.deform_conv2d(xc, omc, w, b, stride=cfg["stride"], padding=cfg["pad"], dilation=cfg["dil"])
    yr.backward(gy)

    def close(a, b_, name):
        a, b_ = a.detach().float().cpu(), b_.detach().float()
        d = (a - b_).abs().max().item()
        assert d <= 2e-4 * max(1.0, b_.abs().max().item()), (name, d)
    close(y, yr, "y"); close(xd.grad, xc.grad, "dx"); close(omd.grad, omc.grad, "doffset(+dmask)")
    close(mod.weight.grad, w.grad, "dweight"); close(mod.bias.grad, b.grad, "dbias")


def test_roi_align_backward_vs_torchvision(dev):
    import upsnet_b200 as U
    g = torch.Generator().manual_seed(4)
    feat = torch.randn(2, 16, 30, 44, generator=g)
    rng = np.random.default_rng(2)
    n = 25
    c = rng.uniform(0, 1, (n, 2)) * np.array([170, 115]); s = np.exp(rng.uniform(np.log(4), np.log(150), (n, 2)))
    rois = np.concatenate([rng.integers(0, 2, (n, 1)), c - s / 2, c + s / 2], 1).astype(np.float32)
    rois[0, 1:] = [-20, -10, 30, 25]
    rois_t = torch.from_numpy(rois)
    gy = torch.randn(n, 16, 7, 7, generator=g)
    fd = feat.to(dev).requires_grad_(True)
    y = U.RoIAlign(7, 7, 0.25)(fd, rois_t.to(dev))
    y.backward(gy.to(dev))
    fc = feat.clone().requires_grad_(True)
    yr = torchvision.ops.roi_align(fc, rois_t, (7, 7), 0.25, 2, False)
    yr.backward(gy)
    assert (y.detach().cpu() - yr.detach()).abs().max().item() < 1e-5
    assert (fd.grad.cpu() - fc.grad).abs().max().item() < 1e-4


def test_mask_term_is_differentiable(dev):
    """MaskTerm (training twin of the mask paste) back-propagates to the mask logits through the device tensor ops."""
    import upsnet_b200 as U
    masks = torch.randn(3, 1, 28, 28, device=dev, requires_grad=True)
    boxes = torch.tensor([[0, 8, 8, 100, 90], [0, 40, 20, 160, 120], [0, 0, 0, 30, 30]], dtype=torch.float32, device=dev)
    seg = torch.zeros(1, 19, 48, 80, device=dev)
    e = U.MaskTerm(19, box_scale=0.25)(masks, boxes, torch.tensor([1, 2, 3], device=dev), seg)
    e.sum().backward()
    assert masks.grad is not None and float(masks.grad.abs().sum()) > 0
