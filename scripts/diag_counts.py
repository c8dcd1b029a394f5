"""Prints the device-side result sizes (detections, panoptic candidates, kept instances) of the synthetic bench
workload -- they decide how much of the static-capacity mask branch / panoptic head is live work."""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import upsnet_b200 as U  # noqa: E402
from upsnet_b200.synthetic import synthetic_input, synthetic_model

U.set_precision("bf16")
dev = torch.device("cuda", 0)
model = synthetic_model(device=dev)
for seed in range(4):
    data = synthetic_input(device=dev, seed=seed)
    out = model(data)
    torch.cuda.synchronize()
    print("seed", seed, "n_det", out["cls_probs"].shape[0], "n_pan_candidates", out["panoptic_cls_inds"].shape[0],
          "labels", torch.unique(out["panoptic_outputs"]).numel())
    ent = next(iter(model._graphs.values()))
    print("   counts(n1,n2,k) =", ent[2]["counts"].tolist())
