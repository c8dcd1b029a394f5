"""ncu driver for the non-conv kernels: FPN ROIAlign (bf16 NHWC, 1000 rois), segmented NMS (5 x 1000), RPN top-k, and
the fused panoptic head (19 x 1024 x 2048, 100 instances).  Usage:
  ncu --set full -k regex:'roi_align|nms_|pan_|topk_' -c 24 python scripts/prof_ops.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import upsnet_b200 as U
from upsnet_b200 import operators as OPS
dev = torch.device("cuda", 0)
rng = np.random.default_rng(0)
def t(a): return torch.from_numpy(np.ascontiguousarray(a)).to(dev)
def cl(x): return x.contiguous(memory_format=torch.channels_last)
# ROIAlign
feats = [cl(torch.randn(1, 256, 256 >> l, 512 >> l, device=dev).bfloat16()) for l in range(4)]
c = rng.uniform(0, [2048, 1024], (1000, 2)); s = np.exp(rng.uniform(np.log(16), np.log(512), (1000, 2)))
rois = np.concatenate([np.zeros((1000, 1)), np.clip(c - s / 2, 0, [2047, 1023]), np.clip(c + s / 2, 0, [2047, 1023])], 1).astype(np.float32)
# NMS
cb = rng.uniform(0, [2048, 1024], (5000, 2)); sb = rng.uniform(16, 256, (5000, 2))
boxes = t(np.concatenate([cb - sb / 2, cb + sb / 2], 1).astype(np.float32))
offs = torch.tensor([0, 1000, 2000, 3000, 4000, 5000], dtype=torch.int32, device=dev)
# top-k
probs = [torch.sigmoid(torch.randn(3, 256 >> l, 512 >> l, device=dev) * 3) for l in range(5)]
# panoptic head
H, W, n = 1024, 2048, 100
fcn = torch.randn(1, 19, H, W, device=dev) * 3
cc = rng.uniform(0, [W, H], (n, 2)); ss = rng.uniform(32, 400, (n, 2))
b = np.concatenate([np.clip(cc - ss / 2, 0, [W - 1, H - 1]), np.clip(cc + ss / 2, 0, [W - 1, H - 1])], 1).astype(np.float32)
prob = ((rng.permutation(n) + 1.0) / (n + 1)).astype(np.float32)
ml = (rng.standard_normal((n, 1, 28, 28)) * 2 + 0.5).astype(np.float32)
cls = rng.integers(1, 9, n).astype(np.int64)
for _ in range(2):
    U.fpn_roi_align(feats, t(rois), 7, 7, [1 / 4., 1 / 8., 1 / 16., 1 / 32.], layout="nhwc")
    OPS.nms_segmented(boxes, offs, 1000, 0.7)
    OPS.rpn_topk(probs, 3, 1000)
    U.panoptic_fuse(fcn, t(b), t(prob), t(ml), t(cls), 11, want_sem=True)
    torch.cuda.synchronize()
