"""DCN pixel-block experiment (pair stream): per-layer CUDA-event timing of the semantic-head deformable convs."""
import os, sys, subprocess, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) == 1:
    for tile in ("0", "88", "84"):
        env = dict(os.environ, UPSNET_DCN_TILE=tile)
        print("== UPSNET_DCN_TILE=%s" % tile, flush=True)
        subprocess.run([sys.executable, __file__, "run"], env=env)
    sys.exit(0)
import torch
import upsnet_b200 as U
from upsnet_b200.operators import Pair
dev = torch.device("cuda", 0)
U.set_precision("bf16x3")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for (H, W, Cin, Cout) in ((256, 512, 256, 128), (256, 512, 128, 128), (128, 256, 256, 128), (64, 128, 256, 128)):
    x = Pair.from_float(torch.randn(1, Cin, H, W, device=dev))
    w = torch.randn(Cout, Cin, 3, 3, device=dev) / 48
    off = torch.randn(1, 18, H, W, device=dev) * 1.5
    f = lambda: U.deform_conv(x, off, w, None, 1, 1, 1, relu=True, precision=1)
    for _ in range(3): f()
    tot = 0.0
    for _ in range(10):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); f(); b.record(); torch.cuda.synchronize(); tot += a.elapsed_time(b)
    print("dcn pair %dx%d %d->%d: %.1f us" % (H, W, Cin, Cout, 100 * tot))
