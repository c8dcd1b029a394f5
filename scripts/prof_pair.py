"""ncu driver for the pair-stream (bf16x3) kernels: one launch each of representative layers on hi/lo bf16 pairs.
Usage: ncu --set full --clock-control none --import-source on -k regex:igemm -s <warm> -c <n> python scripts/prof_pair.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import upsnet_b200 as U
from upsnet_b200.operators import Pair
dev = torch.device("cuda", 0)
torch.manual_seed(0)
def act(n, c, h, w): return Pair.from_float(torch.randn(n, c, h, w, device=dev))
def wgt(co, ci, k): return torch.randn(co, ci, k, k, device=dev) / (ci * k * k) ** 0.5
U.set_precision("bf16x3")
layers = [
    ("fpn/rpn 3x3 256->256 @256x512", act(1, 256, 256, 512), wgt(256, 256, 3), 1, None),
    ("mask head 3x3 256->256 N256 14x14", act(256, 256, 14, 14), wgt(256, 256, 3), 1, None),
    ("res4 conv2 3x3 256->256 @64x128", act(1, 256, 64, 128), wgt(256, 256, 3), 1, None),
    ("res2 conv3 1x1 64->256 +res @256x512", act(1, 64, 256, 512), wgt(256, 64, 1), 0, act(1, 256, 256, 512)),
    ("res4 conv3 1x1 256->1024 +res @64x128", act(1, 256, 64, 128), wgt(1024, 256, 1), 0, act(1, 1024, 64, 128)),
    ("res5 conv2 3x3 512->512 @32x64", act(1, 512, 32, 64), wgt(512, 512, 3), 1, None),
]
off = torch.randn(1, 18, 256, 512, device=dev) * 1.5
xd, wd = act(1, 256, 256, 512), wgt(128, 256, 3)
def run():
    for name, x, w, pad, res in layers:
        U.conv2d(x, w, None, 1, pad, 1, residual=res, relu=True)
    U.deform_conv(xd, off, wd, None, 1, 1, 1, relu=True)          # semantic head L0 @ P2 (igemm_tc_kernel<1,2>)
run(); torch.cuda.synchronize()
run(); torch.cuda.synchronize()
print("\n".join(l[0] for l in layers) + "\ndcn pair 256->128 @256x512")
