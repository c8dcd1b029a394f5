"""Small driver for ncu: the dominant kernel family on representative layer shapes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import upsnet_b200 as U
dev = torch.device("cuda", 0)
torch.manual_seed(0)
prec = {"bf16": 2, "bf16x3": 1}[os.environ.get("PREC", "bf16")]
x = torch.randn(1, 256, 256, 512, device=dev).permute(0, 2, 3, 1).contiguous().permute(0, 3, 1, 2)   # channels_last
w3 = torch.randn(256, 256, 3, 3, device=dev) / 48
off = torch.randn(1, 18, 256, 512, device=dev) * 2
wd = torch.randn(128, 256, 3, 3, device=dev) / 48
x64 = torch.randn(1, 64, 256, 512, device=dev).permute(0, 2, 3, 1).contiguous().permute(0, 3, 1, 2)
w1 = torch.randn(256, 64, 1, 1, device=dev) / 8
for _ in range(3):
    U.conv2d(x, w3, None, 1, 1, 1, precision=prec)          # FPN output conv shape (a4)
    U.deform_conv(x, off, wd, None, 1, 1, 1, precision=prec)  # semantic head L1@P2 (a12)
    U.conv2d(x64, w1, None, precision=prec)                   # res2 1x1 64->256 (a2, HBM-bound)
torch.cuda.synchronize()
