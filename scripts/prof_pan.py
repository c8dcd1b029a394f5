"""Panoptic head alone (BASELINE configs[4]) for one n: python scripts/prof_pan.py [n]   (ncu --metrics gpu__time_duration.sum ...)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import upsnet_b200 as U
dev = torch.device("cuda", 0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
H, W = 1024, 2048
rng = np.random.default_rng(5)
fcn = torch.randn(1, 19, H, W, device=dev) * 3
c = np.stack([rng.uniform(0, W, n), rng.uniform(0, H, n)], 1); s = np.exp(rng.uniform(np.log(16), np.log(512), (n, 2)))
b = np.concatenate([c - s / 2, c + s / 2], 1).astype(np.float32)
b[:, 0::2] = np.clip(b[:, 0::2], 0, W - 1); b[:, 1::2] = np.clip(b[:, 1::2], 0, H - 1)
a = [torch.from_numpy(v).to(dev) for v in (b, (0.6 + 0.4 * (rng.permutation(n) + 1) / (n + 1)).astype(np.float32),
                                           (rng.standard_normal((n, 1, 28, 28)) * 2).astype(np.float32), rng.integers(1, 9, n).astype(np.int64))]
nd = torch.tensor([n], dtype=torch.int32, device=dev)
for _ in range(3):
    out = U.panoptic_fuse(fcn, a[0], a[1], a[2], a[3], 11, n_dev=nd)
torch.cuda.synchronize()
print("kept", int(out[-1].item()))
