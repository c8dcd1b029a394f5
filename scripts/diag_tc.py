"""GPU diagnostic (not a test): localise tcgen05-vs-fp32 differences."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import upsnet_b200 as U
from upsnet_b200 import operators as ops
dev = torch.device("cuda", 0)
rng = np.random.default_rng(6)

def dcn_case(H, W, pad, dil, Cin=64, Cout=64, zero_off=False, off_scale=2.5, stride=1):
    x = torch.from_numpy(rng.standard_normal((1, Cin, H, W)).astype(np.float32)).to(dev)
    w = torch.from_numpy((rng.standard_normal((Cout, Cin, 3, 3)) / np.sqrt(Cin * 9)).astype(np.float32)).to(dev)
    Ho = (H + 2 * pad - (dil * 2 + 1)) // stride + 1; Wo = (W + 2 * pad - (dil * 2 + 1)) // stride + 1
    off = torch.from_numpy((rng.standard_normal((1, 18, Ho, Wo)) * (0 if zero_off else off_scale)).astype(np.float32)).to(dev)
    a = U.deform_conv(x, off, w, None, stride, pad, dil, 1, precision=0)
    b = U.deform_conv(x, off, w, None, stride, pad, dil, 1, precision=1).contiguous()
    d = (a - b).abs()
    bad = (d > 1e-3)
    print("dcn H=%d W=%d pad=%d dil=%d zero_off=%s scale=%.1f: max %.4g, bad %d/%d" % (H, W, pad, dil, zero_off, off_scale, d.max().item(), int(bad.sum()), d.numel()))
    if bad.any():
        idx = bad.nonzero()
        print("   bad rows(ho) uniq:", sorted(set(idx[:, 2].tolist()))[:30], " cols(wo) uniq:", sorted(set(idx[:, 3].tolist()))[:30], " chans:", len(set(idx[:, 1].tolist())))

for args in [(16, 16, 1, 1), (20, 20, 1, 1), (20, 20, 2, 2), (16, 16, 2, 2), (20, 20, 2, 1), (20, 20, 1, 2)]:
    dcn_case(*args)
dcn_case(20, 20, 2, 2, zero_off=True)
dcn_case(20, 20, 2, 2, off_scale=0.3)
dcn_case(32, 32, 2, 2)
dcn_case(20, 20, 3, 3)

# ---- engine: stage-by-stage fp32 vs bf16x3 ----
from upsnet_b200.model import UPSNetConfig
from upsnet_b200.synthetic import synthetic_input, synthetic_model
m = synthetic_model(UPSNetConfig.cityscapes_r50(), depth=(1, 1, 1, 1), seed=3, device=dev)
inp = synthetic_input(256, 384, seed=4, device=dev)
def stages(prec):
    U.set_precision(prec)
    out = {}
    with torch.no_grad():
        bb = m.resnet_backbone
        c1 = bb.conv1(inp["data"]); out["c1"] = c1
        r2 = bb.res2(c1); r3 = bb.res3(r2); r4 = bb.res4(r3); r5 = bb.res5(r4)
        out.update(r2=r2, r3=r3, r4=r4, r5=r5)
        p = m.fpn(r2, r3, r4, r5)
        for i, t in enumerate(p): out["p%d" % (i + 2)] = t
        sub = m.fcn_head.fcn_subnet
        x = p[0]
        for li in range(sub.num_layers):
            l = sub.conv[li][0]
            off = ops.conv2d(x, l.conv_offset.weight, l.conv_offset.bias, 1, 1, 1, out_format="nchw")
            out["fcn_p2_off%d" % li] = off
            x = ops.deform_conv(x, off, l.conv.weight, l.conv.bias, 1, 1, 1, 1, relu=True)
            out["fcn_p2_l%d" % li] = x
        out["fcn_output"] = m.fcn_head(*p[:4])["fcn_output"]
    U.set_precision("fp32")
    return {k: v.float().contiguous() for k, v in out.items()}
a = stages("fp32"); b = stages("bf16x3"); c = stages("bf16")
for k in a:
    sc = a[k].abs().max().item()
    print("%-12s scale %9.3f  x3 relerr %.3g   bf16 relerr %.3g" % (k, sc, (a[k] - b[k]).abs().max().item() / sc, (a[k] - c[k]).abs().max().item() / sc))
