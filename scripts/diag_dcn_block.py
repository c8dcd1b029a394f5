"""Diagnostic: per-op comparison of DCN bottlenecks (pair stream) against torch / torchvision on the same inputs."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torchvision
import torch.nn.functional as F
import upsnet_b200 as U
from upsnet_b200 import operators as ops
from upsnet_b200.model import UPSNetConfig, _fold_bn
from upsnet_b200.synthetic import synthetic_input, synthetic_model

dev = torch.device("cuda", 0)
m = synthetic_model(UPSNetConfig.coco_r101_dcn(), depth=(3, 4, 23, 3), seed=1, device=dev)
U.set_precision("bf16x3")
inp = synthetic_input(800, 1344, seed=5, device=dev)
def rel(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return ((a - b).abs().max() / max(1.0, float(b.abs().max()))).item()
with torch.no_grad():
    bb = m.resnet_backbone
    x = bb.conv1(inp["data"]); x = bb.res2(x); x = bb.res3(x)
    for name, blk in (("res4", bb.res4), ("res5", bb.res5)):
        for bi, b in enumerate(blk.layers):
            f = b._f
            xin = x
            xf = xin.float().cpu()
            out1 = ops.conv2d(xin, f["w1"], f["b1"], stride=b.stride, relu=True)
            w1 = F.relu(F.conv2d(xf, f["w1"].cpu(), f["b1"].cpu(), b.stride))
            e1 = rel(out1, w1)
            off = ops.conv2d(out1, b.conv2_offset.weight, b.conv2_offset.bias, 1, 1, 1, out_format="nchw")
            o1f = out1.float().cpu()
            woff = F.conv2d(o1f, b.conv2_offset.weight.cpu(), b.conv2_offset.bias.cpu(), 1, 1, 1)
            e2 = rel(off, woff)
            out2 = ops.deform_conv(out1, off, f["w2"], f["b2"], 1, b.dilation, b.dilation, relu=True)
            w2 = F.relu(torchvision.ops.deform_conv2d(o1f, off.float().cpu(), f["w2"].cpu(), f["b2"].cpu(), padding=b.dilation, dilation=b.dilation))
            e3 = rel(out2, w2)
            res = xin if b.downsample is None else ops.conv2d(xin, f["wd"], f["bd"], stride=b.stride)
            rf = xf if b.downsample is None else F.conv2d(xf, f["wd"].cpu(), f["bd"].cpu(), b.stride)
            e4 = rel(res, rf)
            out3 = ops.conv2d(out2, f["w3"], f["b3"], residual=res, relu=True)
            w3 = F.relu(F.conv2d(out2.float().cpu(), f["w3"].cpu(), f["b3"].cpu()) + res.float().cpu())
            e5 = rel(out3, w3)
            print(name, bi, tuple(xin.shape), "conv1 %.2e offset %.2e dcn %.2e down %.2e conv3 %.2e |x| %.2f" % (e1, e2, e3, e4, e5, float(out3.float().abs().max())), flush=True)
            x = out3
            if bi >= 3 and name == "res4":
                break
U.set_precision("fp32")
