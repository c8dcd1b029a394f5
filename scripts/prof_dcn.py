"""ncu driver for the fused deformable conv (igemm_tc_kernel<1, true>): the first semantic-head layer at P2 and a
P3 layer, bf16 NHWC activations.  Usage: ncu --set full -k regex:igemm_tc_kernel -s 2 -c 2 python scripts/prof_dcn.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import upsnet_b200 as U
dev = torch.device("cuda", 0)
torch.manual_seed(0)
def cl(t): return t.contiguous(memory_format=torch.channels_last)
U.set_precision("bf16")
cases = []
for (c, h, w) in ((256, 256, 512), (256, 128, 256)):
    x = cl(torch.randn(1, c, h, w, device=dev).bfloat16())
    off = torch.randn(1, 18, h, w, device=dev) * 1.5
    wt = torch.randn(128, c, 3, 3, device=dev) / (c * 9) ** 0.5
    cases.append((x, off, wt))
for _ in range(2):
    for x, off, wt in cases:
        U.deform_conv(x, off, wt, None, 1, 1, 1, 1, relu=True)
    torch.cuda.synchronize()
