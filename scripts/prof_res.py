"""ncu driver: epilogue-heavy layers (1x1 conv + residual, bf16 activation stream)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import upsnet_b200 as U
dev = torch.device("cuda", 0)
torch.manual_seed(0)
def cl(t): return t.contiguous(memory_format=torch.channels_last)
x64 = cl(torch.randn(1, 64, 256, 512, device=dev).bfloat16())
res = cl(torch.randn(1, 256, 256, 512, device=dev).bfloat16())
w1 = torch.randn(256, 64, 1, 1, device=dev) / 8
x256 = cl(torch.randn(1, 256, 64, 128, device=dev).bfloat16())
res4 = cl(torch.randn(1, 1024, 64, 128, device=dev).bfloat16())
w4 = torch.randn(1024, 256, 1, 1, device=dev) / 16
U.set_precision("bf16")
for _ in range(3):
    U.conv2d(x64, w1, None, residual=res, relu=True)        # res2 conv3 (a2)
    U.conv2d(x64, w1, None, relu=True)                      # same without residual
    U.conv2d(x256, w4, None, residual=res4, relu=True)      # res4 conv3
torch.cuda.synchronize()
