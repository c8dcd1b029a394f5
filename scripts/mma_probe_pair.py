"""Timing experiment (pair stream, precision bf16x3): what paces a k-block of the TMA conv kernel?  Debug variants of the
MMA loop (results are garbage in modes 1-3; only time matters): 1 = alternate the two TMEM accumulators per k-slice,
2 = first k-slice only (3 MMAs per k-block instead of 12), 3 = no MMAs (operand delivery only)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["UPSNET_TMA_HALO"] = os.environ.get("UPSNET_TMA_HALO", "0")
import torch
import upsnet_b200 as U
from upsnet_b200.operators import Pair
dev = torch.device("cuda", 0)
torch.manual_seed(0)
def act(n, c, h, w): return Pair.from_float(torch.randn(n, c, h, w, device=dev))
def wgt(co, ci, k): return torch.randn(co, ci, k, k, device=dev) / (ci * k * k) ** 0.5
U.set_precision("bf16x3")
layers = [("res4 conv2 3x3 256->256 @64x128 (BN128)", act(1, 256, 64, 128), wgt(256, 256, 3), 1, None),
          ("res5 conv2 3x3 512->512 @32x64 (BN64)", act(1, 512, 32, 64), wgt(512, 512, 3), 1, None),
          ("offset conv 3x3 256->18 @256x512 (BN32, fp32 NCHW out)", act(1, 256, 256, 512), wgt(18, 256, 3), 1, "nchw"),
          ("fpn 3x3 256->256 @128x256 (BN128)", act(1, 256, 128, 256), wgt(256, 256, 3), 1, None),
          ("res4 conv1 1x1 1024->256 @64x128", act(1, 1024, 64, 128), wgt(256, 1024, 1), 0, None)]
def run(x, w, pad, fmt, reps=20):
    f = lambda: U.conv2d(x, w, None, 1, pad, 1, relu=fmt is None, out_format=fmt)
    for _ in range(3):
        f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        with torch.cuda.graph(g):
            for _ in range(reps):
                f()
    torch.cuda.current_stream().wait_stream(s)
    g.replay(); torch.cuda.synchronize()
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / reps
print("halo env =", os.environ["UPSNET_TMA_HALO"])
print("| layer | normal | alt accumulators | 1 k-slice / k-block | no MMAs |\n|---|---:|---:|---:|---:|")
for name, x, w, pad, fmt in layers:
    t = []
    for mode in (0, 1, 2, 3):
        os.environ["UPSNET_TMA_DEBUG"] = str(mode)
        t.append(run(x, w, pad, fmt))
    os.environ["UPSNET_TMA_DEBUG"] = "0"
    print("| %s | %.1f | %.1f | %.1f | %.1f |" % (name, *t), flush=True)
