"""The python-level operators of the hot path against fixtures produced by EXECUTING the reference's own code
(tests/golden/make_reference_modules.py -> reference_modules.npz): MaskRemoval + SegTerm + the
panoptic glue (rows a14-a16), MaskROI (a10) and PyramidProposal (a6).  CPU part: the C oracle, its literal numpy twin and
the product's host logic (torch restatement with the oracle-backed CPU ops).  The CUDA kernels meet the same fixtures in
tests/test_gpu_reference_fixtures.py."""
import os

import numpy as np
import pytest
import torch

from oracle import oracle as O
from oracle.cpu_model import cpu_ops

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def ref():
    return np.load(os.path.join(ROOT, "tests", "golden", "reference_modules.npz"))


def pan_case(ref, i):
    p = "pan%d_" % i
    return {k: ref[p + k] for k in ("fcn", "rois", "prob", "mask_score", "cls", "keep", "panoptic", "kept_cls",
                                    "mask_energy_sum", "mask_energy_cnt", "seg_inst_sum", "seg_inst_cnt")}


def test_fixture_inventory(ref):
    assert int(ref["pan_cases"]) >= 3 and int(ref["mroi_cases"]) >= 3 and int(ref["pp_cases"]) >= 3


@pytest.mark.parametrize("i", range(5))
def test_panoptic_oracle_vs_reference_modules(ref, i):
    """Fused C oracle and its literal numpy twin == reference MaskRemoval (real cv2.resize) + SegTerm + glue, bit for bit."""
    c = pan_case(ref, i)
    boxes, ml = c["rois"][:, 1:], c["mask_score"][:, 0]
    want = c["panoptic"][0]
    # literal numpy twin with the real cv2.resize: every line of the restatement is pinned, bit for bit
    k3, l3 = O.panoptic_head_literal(c["fcn"][0], boxes, c["prob"], ml, c["cls"], 11, resize="cv2")
    assert k3.tolist() == c["keep"].tolist() and np.array_equal(l3, want)
    # oracle of record (OpenCV's documented bilinear formula in un-fused fp32): cv2.resize itself is not bit-reproducible
    # (<= 1.2e-5 abs), so a pixel may flip ONLY where the two best panoptic logits are closer than that
    k2, l2, logits = O.panoptic_head_literal(c["fcn"][0], boxes, c["prob"], ml, c["cls"], 11, return_logits=True)
    keep, labels = O.panoptic_head(c["fcn"][0], boxes, c["prob"], ml, c["cls"], 11)
    assert keep.tolist() == c["keep"].tolist() and k2.tolist() == c["keep"].tolist()
    assert np.array_equal(labels, l2)                       # fused C == literal numpy, always
    diff = np.argwhere(l2 != want)
    assert len(diff) <= 2, len(diff)
    for y, x in diff:
        top2 = np.sort(logits[:, y, x])[-2:]
        assert top2[1] - top2[0] < 5e-5, (y, x, top2)


@pytest.mark.parametrize("i", range(5))
def test_segterm_module_vs_reference(ref, i):
    """The product's SegTerm (API-parity module, pure tensor slicing) reproduces the reference's instance planes."""
    from upsnet_b200.operators import SegTerm
    c = pan_case(ref, i)
    keep = c["keep"]
    rois_k = torch.from_numpy(c["rois"][keep]) * 4.0                       # resnet_upsnet.py:227 passes mask_rois * 4
    seg, inst = SegTerm(19, num_classes=9)(torch.from_numpy(c["cls"][keep]), torch.from_numpy(c["fcn"]), rois_k)
    assert seg.shape[1] == 11
    assert np.array_equal((inst != 0).sum(dim=(0, 2, 3)).numpy(), c["seg_inst_cnt"])
    np.testing.assert_allclose(inst.double().sum(dim=(0, 2, 3)).numpy(), c["seg_inst_sum"], rtol=0, atol=1e-9)


def mroi_case(ref, i):
    p = "mroi%d_" % i
    return {k: ref[p + k] for k in ("rois", "delta", "prob", "agnostic", "score_thresh", "out_scores", "out_boxes", "out_cls")}


def _canon(sc, bx, cls):
    """Order inside a run of EQUAL (class, score) is unspecified in the reference (np.argsort()[::-1] of tied scores,
    nms.py:66 / gpu_nms.pyx:32): compare such runs as sets -- rows sorted by box coordinates inside each run."""
    rows = list(range(len(sc)))
    out, i = [], 0
    while i < len(rows):
        j = i
        while j < len(rows) and cls[j] == cls[i] and sc[j] == sc[i]:
            j += 1
        out += sorted(rows[i:j], key=lambda r: tuple(np.round(bx[r], 1)))
        i = j
    return np.array(out, np.int64)


def _check_mroi(c, sc, bx, cls):
    assert sc.shape[0] == c["out_scores"].shape[0], (sc.shape, c["out_scores"].shape)
    assert np.array_equal(cls, c["out_cls"])
    assert np.array_equal(sc, c["out_scores"])
    a, b = _canon(sc, bx, cls), _canon(c["out_scores"], c["out_boxes"], c["out_cls"])
    np.testing.assert_allclose(bx[a], c["out_boxes"][b], rtol=0, atol=2e-3)      # exp() of the decode differs in the last ulp


@pytest.mark.parametrize("i", range(6))
def test_maskroi_host_logic_vs_reference(ref, i):
    """detection.MaskROI / StaticMaskROI (torch restatement; NMS = the C oracle here) == reference MaskROI.forward:
    candidate threshold, per-class NMS, max_det threshold WITH ties, class-major order, the 'nothing survives' dummy."""
    from upsnet_b200.detection import MaskROI, StaticMaskROI
    c = mroi_case(ref, i)
    info = ref["mroi_im_info"][0]
    args = (100, 9, 0.5, bool(c["agnostic"]), float(c["score_thresh"]), (10., 10., 5., 5.))
    rois, delta, prob = (torch.from_numpy(c[k]) for k in ("rois", "delta", "prob"))
    with cpu_ops():
        sc, bx, cls = MaskROI(*args)(rois, delta, prob, info)
        _check_mroi(c, sc.numpy(), bx.numpy(), cls.numpy())
        valid = torch.ones(rois.shape[0], dtype=torch.bool)
        s_sc, s_bx, s_cls, n = StaticMaskROI(*args)(rois, valid, delta, prob, info)
        n = int(n)
        _check_mroi(c, s_sc[:n].numpy(), s_bx[:n].numpy(), s_cls[:n].numpy())


def pp_case(ref, i):
    p = "pp%d_" % i
    return ([ref[p + "prob%d" % l] for l in range(5)], [ref[p + "delta%d" % l] for l in range(5)], ref[p + "im_info"][0],
            int(ref[p + "pre"]), int(ref[p + "post"]), ref[p + "rois"], ref[p + "scores"])


@pytest.mark.parametrize("i", range(3))
def test_proposals_host_logic_vs_reference(ref, i):
    """detection.ProposalGenerator / StaticProposalGenerator == reference PyramidProposal (individual_proposals=True):
    per-level top-k order, decode + clip, per-level NMS, post-NMS cut, final top-N by score."""
    from upsnet_b200.detection import ProposalGenerator, StaticProposalGenerator
    probs, deltas, info, pre, post, want_rois, want_sc = pp_case(ref, i)
    tp = [torch.from_numpy(p_) for p_ in probs]; td = [torch.from_numpy(d_) for d_ in deltas]
    with cpu_ops():
        rois, sc = ProposalGenerator(pre_nms_top_n=pre, post_nms_top_n=post)(tp, td, info)
        assert rois.shape == want_rois.shape
        assert np.array_equal(sc.numpy(), want_sc)
        np.testing.assert_allclose(rois.numpy(), want_rois, rtol=0, atol=2e-3)
        s_rois, s_sc, ok = StaticProposalGenerator(pre_nms_top_n=pre, post_nms_top_n=post)(tp, td, info)
        n = int(ok.sum())
        assert n == want_rois.shape[0] and bool(ok[:n].all())
        assert np.array_equal(s_sc[:n].numpy(), want_sc)
        np.testing.assert_allclose(s_rois[:n].numpy(), want_rois, rtol=0, atol=2e-3)


def test_mask_term_vs_reference(ref):
    """MaskTerm (training twin, row a17): bilinear (align_corners=False) resize + paste at 1/4 scale."""
    from upsnet_b200.operators import MaskTerm
    seg = torch.zeros(tuple(int(v) for v in ref["mterm_seg_shape"]))
    got = MaskTerm(19, box_scale=1 / 4.0)(torch.from_numpy(ref["mterm_masks"]), torch.from_numpy(ref["mterm_rois"]),
                                         torch.from_numpy(ref["mterm_cls"]), seg)
    assert got.shape == ref["mterm_energy"].shape
    np.testing.assert_allclose(got.numpy(), ref["mterm_energy"], rtol=0, atol=1e-6)


def test_mask_matching_vs_reference(ref):
    from upsnet_b200.operators import MaskMatching
    mm = MaskMatching(19, enable_void=True)
    segs, masks = torch.from_numpy(ref["mmatch_gt_segs"]), torch.from_numpy(ref["mmatch_gt_masks"])
    assert np.array_equal(mm(segs, masks).numpy(), ref["mmatch_all"])
    assert np.array_equal(mm(segs, masks, torch.from_numpy(ref["mmatch_keep"])).numpy(), ref["mmatch_kept"])


@pytest.mark.parametrize("i", range(3))
def test_unified_pan_oracle_vs_reference(ref, i):
    """oracle.unified_pan_result == the reference's get_unified_pan_result (its own source lines, executed)."""
    p = "uni%d_" % i
    got = O.unified_pan_result(ref[p + "seg"], ref[p + "pan"], ref[p + "cls"], 19, 9, int(ref[p + "limit"]))
    assert np.array_equal(got, ref[p + "out"])


@pytest.mark.parametrize("i", range(3))
def test_prep_image_oracle_vs_reference(ref, i):
    p = "prep%d_" % i
    blob, hw = O.prep_image(ref[p + "im"], ref["prep_pixel_means"], float(ref[p + "scale"]))
    assert tuple(hw) == tuple(ref[p + "resized_hw"]) and blob.shape == ref[p + "blob"].shape
    assert np.array_equal(blob, ref[p + "blob"])
