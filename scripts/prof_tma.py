"""ncu driver for the TMA-fed conv kernel: one launch each of representative backbone / head layers
(bf16 NHWC activation stream).  Usage: ncu --set full -k regex:igemm_tma -c 8 python scripts/prof_tma.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import upsnet_b200 as U
dev = torch.device("cuda", 0)
torch.manual_seed(0)
def cl(t): return t.contiguous(memory_format=torch.channels_last)
def act(n, c, h, w): return cl(torch.randn(n, c, h, w, device=dev).bfloat16())
def wgt(co, ci, k): return torch.randn(co, ci, k, k, device=dev) / (ci * k * k) ** 0.5
U.set_precision("bf16")
layers = [
    ("res2 conv2 3x3 64->64 @256x512", act(1, 64, 256, 512), wgt(64, 64, 3), 1, None),
    ("res2 conv3 1x1 64->256 +res @256x512", act(1, 64, 256, 512), wgt(256, 64, 1), 0, act(1, 256, 256, 512)),
    ("res4 conv2 3x3 256->256 @64x128", act(1, 256, 64, 128), wgt(256, 256, 3), 1, None),
    ("res4 conv3 1x1 256->1024 +res @64x128", act(1, 256, 64, 128), wgt(1024, 256, 1), 0, act(1, 1024, 64, 128)),
    ("fpn 3x3 256->256 @256x512", act(1, 256, 256, 512), wgt(256, 256, 3), 1, None),
    ("mask head 3x3 256->256 N128 14x14", act(128, 256, 14, 14), wgt(256, 256, 3), 1, None),
    ("res5 conv2 3x3 512->512 @32x64", act(1, 512, 32, 64), wgt(512, 512, 3), 1, None),
    ("res3 conv1 1x1 512->128 @128x256", act(1, 512, 128, 256), wgt(128, 512, 1), 0, None),
]
for name, x, w, pad, res in layers:      # warm-up: packed weights, attribute set-up
    U.conv2d(x, w, None, 1, pad, 1, residual=res, relu=True)
torch.cuda.synchronize()
for name, x, w, pad, res in layers:
    U.conv2d(x, w, None, 1, pad, 1, residual=res, relu=True)
torch.cuda.synchronize()
print("\n".join(l[0] for l in layers))
