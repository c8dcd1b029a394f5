"""Timing experiment: what paces a k-block of the TMA conv kernel for small N tiles?  Runs three 3x3 layers with the
halo mode off and the debug variants of the MMA loop (results are garbage in modes 1-3; only time matters)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["UPSNET_TMA_HALO"] = "0"
import torch
import upsnet_b200 as U
dev = torch.device("cuda", 0)
torch.manual_seed(0)
def cl(t): return t.contiguous(memory_format=torch.channels_last)
def act(n, c, h, w): return cl(torch.randn(n, c, h, w, device=dev).bfloat16())
def wgt(co, ci, k): return torch.randn(co, ci, k, k, device=dev) / (ci * k * k) ** 0.5
U.set_precision("bf16")
layers = [("res2 conv2 3x3 64->64 @256x512", act(1, 64, 256, 512), wgt(64, 64, 3), 1),
          ("res4 conv2 3x3 256->256 @64x128", act(1, 256, 64, 128), wgt(256, 256, 3), 1),
          ("res5 conv2 3x3 512->512 @32x64", act(1, 512, 32, 64), wgt(512, 512, 3), 1),
          ("res3 conv1 1x1 512->128 @128x256", act(1, 512, 128, 256), wgt(128, 512, 1), 0)]
def run(x, w, pad, reps=20):
    for _ in range(3):
        U.conv2d(x, w, None, 1, pad, 1, relu=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        with torch.cuda.graph(g):
            for _ in range(reps):
                U.conv2d(x, w, None, 1, pad, 1, relu=True)
    torch.cuda.current_stream().wait_stream(s)
    g.replay(); torch.cuda.synchronize()
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / reps
print("| layer | normal | alt accumulators | 1 MMA / k-block | no MMAs |\n|---|---:|---:|---:|---:|")
for name, x, w, pad in layers:
    t = []
    for mode in (0, 1, 2, 3):
        os.environ["UPSNET_TMA_DEBUG"] = str(mode)
        t.append(run(x, w, pad))
    os.environ["UPSNET_TMA_DEBUG"] = "0"
    print("| %s | %.1f | %.1f | %.1f | %.1f |" % (name, *t), flush=True)
