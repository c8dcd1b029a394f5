"""A/B of the 2-CTA (cta_group::2) variant on the pair stream: per-layer µs with UPSNET_TMA_2CTA=0 / 1."""
import os, sys, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) == 1:
    for v in ("0", "1"):
        print("== UPSNET_TMA_2CTA=%s" % v, flush=True)
        subprocess.run([sys.executable, __file__, "run"], env=dict(os.environ, UPSNET_TMA_2CTA=v))
    sys.exit(0)
import torch
import upsnet_b200 as U
from upsnet_b200.operators import Pair
dev = torch.device("cuda", 0)
torch.manual_seed(0)
def act(n, c, h, w): return Pair.from_float(torch.randn(n, c, h, w, device=dev))
def wgt(co, ci, k): return torch.randn(co, ci, k, k, device=dev) / (ci * k * k) ** 0.5
U.set_precision("bf16x3")
layers = [("fpn/rpn 3x3 256->256 @256x512", act(1, 256, 256, 512), wgt(256, 256, 3), 1, None),
          ("fpn 3x3 256->256 @128x256", act(1, 256, 128, 256), wgt(256, 256, 3), 1, None),
          ("res4 conv2 3x3 256->256 @64x128", act(1, 256, 64, 128), wgt(256, 256, 3), 1, None),
          ("res5 conv2 3x3 512->512 @32x64", act(1, 512, 32, 64), wgt(512, 512, 3), 1, None),
          ("res3 conv2 3x3 128->128 @128x256", act(1, 128, 128, 256), wgt(128, 128, 3), 1, None),
          ("res2 conv2 3x3 64->64 @256x512", act(1, 64, 256, 512), wgt(64, 64, 3), 1, None),
          ("res4 conv1 1x1 1024->256 @64x128", act(1, 1024, 64, 128), wgt(256, 1024, 1), 0, None),
          ("res3 conv1 1x1 512->128 @128x256", act(1, 512, 128, 256), wgt(128, 512, 1), 0, None),
          ("mask head 3x3 N256 14x14", act(256, 256, 14, 14), wgt(256, 256, 3), 1, None),
          ("fc6 12544->1024 N1000", act(1000, 12544, 1, 1), wgt(1024, 12544, 1), 0, None),
          ("offset conv 3x3 256->18 @256x512 (fp32 nchw)", act(1, 256, 256, 512), wgt(18, 256, 3), 1, "nchw")]
def run(x, w, pad, fmt, reps=20):
    f = lambda: U.conv2d(x, w, None, 1, pad, 1, relu=fmt is None, out_format=fmt)
    for _ in range(3): f()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        with torch.cuda.graph(g):
            for _ in range(reps): f()
    torch.cuda.current_stream().wait_stream(s)
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / reps
for name, x, w, pad, fmt in layers:
    print("%-50s %.1f us" % (name, run(x, w, pad, fmt)), flush=True)
