"""PipelinedEngine with the device-side callers of the forward (rows f2, f3): raw uint8 image in (upsnet_prep_image),
unified 2-channel panoptic map out (upsnet_unified_pan_result) == the host-side reference pipeline around the same model."""
import numpy as np
import pytest
import torch

from oracle import oracle as O

pytestmark = pytest.mark.gpu


def test_pipeline_raw_in_unified_out():
    import upsnet_b200 as U
    from upsnet_b200.model import UPSNetConfig
    from upsnet_b200.synthetic import synthetic_model
    dev = torch.device("cuda", 0)
    m = synthetic_model(UPSNetConfig.cityscapes_r50(), depth=(1, 1, 1, 1), seed=3, device=dev)
    H, W = 256, 384
    rng = np.random.default_rng(9)
    means = U.PipelinedEngine.PIXEL_MEANS
    raws = [torch.from_numpy(rng.integers(0, 256, (H, W, 3)).astype(np.uint8)).pin_memory() for _ in range(3)]
    im_info = np.array([[H, W, 1.0]], np.float32)
    U.set_precision("bf16x3")
    try:
        eng = U.PipelinedEngine(m, im_info, depth=2, with_masks=True, with_unified=True, stuff_area_limit=256)
        tickets, results = [], []
        for r in raws:
            tickets.append(eng.submit(r))
            if len(tickets) > 1:
                results.append({k: v.clone() for k, v in eng.result(tickets[-2]).items()})
        results.append({k: v.clone() for k, v in eng.result(tickets[-1]).items()})
        h2d, d2h = eng.bytes_per_image()
        assert h2d == H * W * 3
        for r, res in zip(raws, results):
            blob, _ = O.prep_image(r.numpy(), means, 1.0)
            with torch.no_grad():
                want = m({"data": torch.from_numpy(blob).to(dev), "im_info": im_info})
            assert torch.equal(res["panoptic_outputs"], want["panoptic_outputs"].cpu())
            assert torch.equal(res["fcn_outputs"], want["fcn_outputs"].cpu())
            assert torch.equal(res["panoptic_cls_inds"], want["panoptic_cls_inds"].cpu())
            assert torch.allclose(res["mask_probs"], want["mask_probs"].cpu())
            uni = O.unified_pan_result(want["fcn_outputs"][0].cpu().numpy(), want["panoptic_outputs"][0].cpu().numpy(),
                                       want["panoptic_cls_inds"].cpu().numpy(), 19, 9, 256)
            assert np.array_equal(res["pan_2ch"].numpy(), uni)
    finally:
        U.set_precision("fp32")
