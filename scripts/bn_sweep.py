"""Times representative TMA-conv layers with the N tile forced to 64 / 128 / 256 (UPSNET_TMA_FORCE_BN), CUDA events,
L2-warm (as inside the engine, where the producer layer has just written the input)."""
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
    ("res2 conv1 1x1 256->64 @256x512", act(1, 256, 256, 512), wgt(64, 256, 1), 0, None),
    ("res2 conv3 1x1 64->256 +res @256x512", act(1, 64, 256, 512), wgt(256, 64, 1), 0, act(1, 256, 256, 512)),
    ("res3 conv2 3x3 128->128 @128x256", act(1, 128, 128, 256), wgt(128, 128, 3), 1, None),
    ("res3 conv3 1x1 128->512 +res @128x256", act(1, 128, 128, 256), wgt(512, 128, 1), 0, act(1, 512, 128, 256)),
    ("res4 conv1 1x1 1024->256 @64x128", act(1, 1024, 64, 128), wgt(256, 1024, 1), 0, None),
    ("res4 conv2 3x3 256->256 @64x128", act(1, 256, 64, 128), wgt(256, 256, 3), 1, None),
    ("res4 conv3 1x1 256->1024 +res @64x128", act(1, 256, 64, 128), wgt(1024, 256, 1), 0, act(1, 1024, 64, 128)),
    ("res5 conv1 1x1 2048->512 @32x64", act(1, 2048, 32, 64), wgt(512, 2048, 1), 0, None),
    ("res5 conv2 3x3 512->512 @32x64", act(1, 512, 32, 64), wgt(512, 512, 3), 1, None),
    ("res5 conv3 1x1 512->2048 +res @32x64", act(1, 512, 32, 64), wgt(2048, 512, 1), 0, act(1, 2048, 32, 64)),
    ("fpn 3x3 256->256 @128x256", act(1, 256, 128, 256), wgt(256, 256, 3), 1, None),
    ("fpn 3x3 256->256 @64x128", act(1, 256, 64, 128), wgt(256, 256, 3), 1, None),
    ("fpn 3x3 256->256 @32x64", act(1, 256, 32, 64), wgt(256, 256, 3), 1, None),
    ("mask head 3x3 256->256 N256 14x14", act(256, 256, 14, 14), wgt(256, 256, 3), 1, None),
    ("mask deconv 1x1 256->1024 N256 14x14", act(256, 256, 14, 14), wgt(1024, 256, 1), 0, None),
    ("fc6 12544->1024 N1000", act(1000, 12544, 1, 1), wgt(1024, 12544, 1), 0, None),
    ("fc7 1024->1024 N1000", act(1000, 1024, 1, 1), wgt(1024, 1024, 1), 0, None),
]
def run(x, w, pad, res, reps=20):
    for _ in range(3):
        U.conv2d(x, w, None, 1, pad, 1, residual=res, relu=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        with torch.cuda.graph(g):
            for _ in range(reps):
                U.conv2d(x, w, None, 1, pad, 1, residual=res, relu=True)
    torch.cuda.current_stream().wait_stream(s)
    g.replay(); torch.cuda.synchronize()
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / reps
print("| layer | default | BN=64 | BN=128 | BN=256 |\n|---|---:|---:|---:|---:|")
for name, x, w, pad, res in layers:
    t = []
    for bn in (None, 64, 128, 256):
        if bn is None: os.environ.pop("UPSNET_TMA_FORCE_BN", None)
        else: os.environ["UPSNET_TMA_FORCE_BN"] = str(bn)
        t.append(run(x, w, pad, res))
    os.environ.pop("UPSNET_TMA_FORCE_BN", None)
    print("| %s | %.1f | %.1f | %.1f | %.1f |" % (name, *t), flush=True)
