import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, upsnet_b200 as U
from upsnet_b200.model import UPSNetConfig
from upsnet_b200.synthetic import synthetic_input, synthetic_model
dev = torch.device("cuda", 0)
U.set_precision("bf16x3")
m = synthetic_model(UPSNetConfig.cityscapes_r50(), seed=0, device=dev)
for s in (0, 100, 101, 102, 103):
    out, _ = m._run_static(synthetic_input(1024, 2048, seed=s, device=dev)["data"], synthetic_input(8, 8)["im_info"][0] * 0 + torch.tensor([1024., 2048., 1.]).numpy())
    torch.cuda.synchronize()
    print("seed", s, "counts (n_det, n_panoptic_candidates, n_kept):", out["counts"].tolist(), "valid rois:", int(out.get("dbg", {}).get("roi_valid", torch.zeros(1)).sum()) if "dbg" in out else "-")
