"""A/B of the 18-channel offset convs of the semantic head on the pair stream (UPSNET_TMA_HALO=3 forces halo / resident-weight
mode for pairs).  Usage: python scripts/exp_offset_conv.py   (run once per environment setting)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import upsnet_b200 as U
from upsnet_b200.operators import Pair
dev = torch.device("cuda", 0)
torch.manual_seed(0)
U.set_precision("bf16x3")
def gpu_ms(fn, iters=20):
    fn(); fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters
for cin, h, w in [(256, 256, 512), (128, 256, 512), (256, 128, 256), (128, 128, 256), (256, 64, 128)]:
    x = Pair.from_float(torch.randn(1, cin, h, w, device=dev))
    wt = torch.randn(18, cin, 3, 3, device=dev) / (cin * 9) ** 0.5
    b = torch.randn(18, device=dev)
    from upsnet_b200 import operators as ops
    ops.DENSE_WINDOW["on"] = True
    ms = gpu_ms(lambda: U.conv2d(x, wt, b, 1, 1, 1, out_format="nchw"))
    ops.DENSE_WINDOW["on"] = False
    ms0 = gpu_ms(lambda: U.conv2d(x, wt, b, 1, 1, 1, out_format="nchw"))
    print("offset conv %d->18 @%dx%d: window %.4f ms   per-tap TMA %.4f ms" % (cin, h, w, ms, ms0), flush=True)
from upsnet_b200 import operators as ops
ops.DENSE_WINDOW.update(max_cout=256, min_pixels=0)
for cin, cout, h, w in [(64, 64, 256, 512), (128, 128, 128, 256), (256, 256, 64, 128), (256, 256, 128, 256)]:
    x = Pair.from_float(torch.randn(1, cin, h, w, device=dev)); wt = torch.randn(cout, cin, 3, 3, device=dev) / (3 * cin ** 0.5); b = torch.randn(cout, device=dev)
    ops.DENSE_WINDOW["on"] = True
    ms = gpu_ms(lambda: U.conv2d(x, wt, b, 1, 1, 1, relu=True))
    ops.DENSE_WINDOW["on"] = False
    ms0 = gpu_ms(lambda: U.conv2d(x, wt, b, 1, 1, 1, relu=True))
    print("conv3x3 %d->%d @%dx%d pair->pair: window %.4f ms   TMA kernels %.4f ms" % (cin, cout, h, w, ms, ms0), flush=True)
