"""Host logic of the engine (module tree, state_dict names, detection glue, model plumbing) on CPU:
the CUDA entry points are substituted by the oracle-backed CPU implementations of
oracle/cpu_model.py (test infrastructure), everything else is the product code unchanged."""
import numpy as np
import pytest
import torch

from oracle import oracle as O
from oracle.cpu_model import cpu_ops, synthetic_input, synthetic_model


def test_anchors_match_reference(golden_ref):
    from upsnet_b200.detection import generate_anchors
    for s in (4, 8, 16, 32, 64):
        a = generate_anchors(s, np.array((8,)) * s, (0.5, 1, 2))
        assert np.array_equal(a, golden_ref["anchors_%d" % s])


def test_bbox_transform_matches_reference(golden_ref):
    from upsnet_b200.detection import bbox_transform, clip_boxes
    g = golden_ref
    pred = bbox_transform(torch.from_numpy(g["bt_boxes"]), torch.from_numpy(g["bt_deltas"]), (10., 10., 5., 5.))
    np.testing.assert_allclose(pred.numpy(), g["bt_pred"], rtol=1e-5, atol=1e-3)
    np.testing.assert_allclose(clip_boxes(torch.from_numpy(g["bt_pred"].copy()), 600, 900).numpy(), g["bt_clipped"])


def test_state_dict_names_follow_reference():
    from upsnet_b200.model import UPSNetConfig, resnet_upsnet
    m = resnet_upsnet([3, 4, 6, 3], UPSNetConfig.cityscapes_r50())
    keys = set(m.state_dict().keys())
    for k in ["resnet_backbone.conv1.conv1.weight", "resnet_backbone.conv1.bn1.running_mean",
              "resnet_backbone.res2.layers.0.downsample.0.weight", "resnet_backbone.res2.layers.0.downsample.1.weight",
              "resnet_backbone.res3.layers.3.conv3.weight", "resnet_backbone.res5.layers.2.bn2.running_var",
              "fpn.fpn_p5_1x1.weight", "fpn.fpn_p2.bias", "rpn.conv_proposal.0.weight", "rpn.cls_score.weight",
              "rpn.bbox_pred.bias", "rcnn.fc6.0.weight", "rcnn.fc7.0.bias", "rcnn.cls_score.weight",
              "rcnn.bbox_pred.weight", "mask_branch.mask_conv1.0.weight", "mask_branch.mask_deconv1.0.weight",
              "mask_branch.mask_score.bias", "fcn_head.fcn_subnet.conv.0.0.conv_offset.weight",
              "fcn_head.fcn_subnet.conv.0.0.conv.weight", "fcn_head.fcn_subnet.conv.1.0.conv.bias",
              "fcn_head.score.weight"]:
        assert k in keys, k
    assert m.rcnn.fc6[0].weight.shape == (1024, 12544) and m.rcnn.bbox_pred.weight.shape == (36, 1024)
    assert m.fcn_head.fcn_subnet.conv[0][0].conv.weight.shape == (128, 256, 3, 3)
    m2 = resnet_upsnet([3, 4, 23, 3], UPSNetConfig.coco_r101_dcn())
    k2 = set(m2.state_dict().keys())
    assert "resnet_backbone.res3.layers.0.conv2_offset.weight" in k2 and "fpn.fpn_gap.weight" in k2
    assert "resnet_backbone.res2.layers.0.conv2_offset.weight" not in k2
    assert m2.fcn_head.fcn_subnet.conv[2][0].conv.weight.shape == (128, 128, 3, 3)
    assert m2.fcn_head.fcn_subnet.conv[1][0].conv.weight.shape == (128, 256, 3, 3)


def _torch_reference_forward(m, x):
    """Independent restatement of the dense part with plain torch.nn modules (BN NOT folded)."""
    import torch.nn.functional as F
    bb = m.resnet_backbone

    def bott(b, x):
        out = F.relu(b.bn1(b.conv1(x)))
        out = F.relu(b.bn2(b.conv2(out)))
        out = b.bn3(b.conv3(out))
        res = x if b.downsample is None else b.downsample(x)
        return F.relu(out + res)

    c1 = F.max_pool2d(F.relu(bb.conv1.bn1(bb.conv1.conv1(x))), 3, 2, 1)
    outs = []
    for blk in (bb.res2, bb.res3, bb.res4, bb.res5):
        for b in blk.layers:
            c1 = bott(b, c1)
        outs.append(c1)
    return outs


def test_backbone_bn_folding_and_fpn_on_cpu():
    m = synthetic_model(depth=(1, 1, 1, 1), seed=1)
    inp = synthetic_input(64, 96, seed=2)
    with cpu_ops(), torch.no_grad():
        r2, r3, r4, r5 = m.resnet_backbone(inp["data"])
        want = _torch_reference_forward(m, inp["data"])
        for a, b in zip((r2, r3, r4, r5), want):
            assert a.shape == b.shape
            assert (a - b).abs().max() < 1e-3 * max(1.0, b.abs().max())
        p = m.fpn(r2, r3, r4, r5)
        assert [t.shape[-2:] for t in p] == [(16, 24), (8, 12), (4, 6), (2, 3), (1, 2)]


@pytest.mark.parametrize("cfgname", ["cityscapes_r50", "coco_r101_dcn"])
def test_end_to_end_forward_on_cpu(cfgname):
    from upsnet_b200.model import UPSNetConfig
    cfg = getattr(UPSNetConfig, cfgname)()
    m = synthetic_model(cfg, depth=(1, 1, 1, 1), seed=3)
    H, W = 128, 192
    inp = synthetic_input(H, W, seed=4)
    with cpu_ops():
        out = m(inp)
    assert set(out.keys()) == {"cls_probs", "pred_boxes", "mask_probs", "fcn_outputs", "cls_inds",
                               "panoptic_cls_inds", "panoptic_cls_probs", "panoptic_outputs"}
    n = out["pred_boxes"].shape[0]
    assert out["pred_boxes"].shape == (n, 5) and out["mask_probs"].shape == (n, cfg.num_classes, 28, 28)
    assert out["panoptic_outputs"].shape == (1, H, W) and out["panoptic_outputs"].dtype == torch.int64
    assert out["fcn_outputs"].shape == (1, H, W)
    k = out["panoptic_cls_inds"].numel()
    lab = out["panoptic_outputs"]
    num_stuff = cfg.num_seg_classes - cfg.num_classes + 1
    assert ((lab < num_stuff + k) | (lab == 255)).all()
    assert n >= 1  # ties at the max_det score threshold are all kept (mask_roi.py:110-113)


def test_static_engine_equals_dynamic_path_on_cpu():
    """The fixed-shape / device-count engine path must make exactly the decisions of the literal,
    variable-length restatement of the reference glue (detection.MaskROI / ProposalGenerator)."""
    from upsnet_b200.model import UPSNetConfig
    m = synthetic_model(UPSNetConfig.cityscapes_r50(), depth=(1, 1, 1, 1), seed=5)
    inp = synthetic_input(128, 192, seed=6)
    with cpu_ops():
        m.static_engine = True
        a = m(inp)
        m.static_engine = False
        b = m(inp)
    assert a.keys() == b.keys()
    for k in a:
        assert a[k].shape == b[k].shape, k
        assert torch.equal(a[k], b[k]) if a[k].dtype != torch.float32 else torch.allclose(a[k], b[k], atol=1e-4), k
