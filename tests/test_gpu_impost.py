"""Row f4 on the GPU: upsnet_im_post_rle / operators.im_post (csrc/impost.cu) against the numpy oracle (bit-exact run
lengths), against the fixtures produced by the reference's own im_post with real cv2, and at the BASELINE image size through
size-independent properties (run lengths sum to H*W, decode == paste)."""
import os

import numpy as np
import pytest
import torch

from oracle import oracle as O

pytestmark = pytest.mark.gpu
G = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_impost.npz"))


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available()
    return torch.device("cuda", 0)


def t(a, dev):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev)


def _run(dev, boxes, masks, cls, H, W, **kw):
    from upsnet_b200 import operators as ops
    counts, run_len, ovf = ops.im_post_rle(t(boxes, dev), t(masks, dev), t(cls, dev), H, W, **kw)
    return counts.cpu().numpy().view(np.uint32), run_len.cpu().numpy(), int(ovf.item())


@pytest.mark.parametrize("ci", [0, 1, 2])
def test_im_post_rle_vs_oracle_and_reference_fixture(dev, ci):
    pre = "c%d_" % ci
    H, W = (int(v) for v in G[pre + "hw"])
    boxes, masks = G[pre + "boxes"], G[pre + "masks"]
    cls = G[pre + "cls"] if masks.shape[1] > 1 else np.zeros(masks.shape[0], np.int64)
    cn, rl, ovf = _run(dev, boxes, masks, cls, H, W)
    assert ovf == 0
    want_imgs = O.im_post_masks(boxes, masks, cls, H, W)
    flips = 0
    for d in range(boxes.shape[0]):
        c = cn[d, :rl[d]]
        assert np.array_equal(c, O.rle_counts(want_imgs[d])), d            # bit-exact against the oracle of record
        flips += int((O.rle_decode(c, H, W) != G[pre + "images"][d]).sum())  # reference-executed (real cv2) images
    assert flips <= 2, flips                                                 # cv2 SIMD rounding at the 0.5 threshold


def test_im_post_dropin_matches_reference_grouping(dev):
    """operators.im_post has im_post's signature and side effects (boxes_all / masks_all per class)."""
    from upsnet_b200 import operators as ops
    pre = "c0_"
    H, W = (int(v) for v in G[pre + "hw"])
    boxes, masks, cls, scores = G[pre + "boxes"], G[pre + "masks"], G[pre + "cls"], G[pre + "scores"]
    boxes_all = [[] for _ in range(9)]
    masks_all = [[] for _ in range(9)]
    pb = np.concatenate([np.zeros((boxes.shape[0], 1), np.float32), boxes], 1)      # model layout: batch index first
    ops.im_post(boxes_all, masks_all, t(scores, dev), t(pb, dev), t(masks, dev), t(cls, dev), 9, (H, W))
    assert np.allclose(boxes_all[1][0], G[pre + "cls_boxes_1"])
    want_imgs = O.im_post_masks(boxes, masks, cls, H, W)
    for idx in range(1, 9):
        sel = np.flatnonzero(cls == idx)
        segs = masks_all[idx][0]
        assert len(segs) == sel.size
        for d, seg in zip(sel, segs):
            assert seg["size"] == [H, W] and isinstance(seg["counts"], str)
            assert seg["counts"] == O.rle_to_string(O.rle_counts(want_imgs[d])).decode()
            if np.array_equal(want_imgs[d], G[pre + "images"][d]):
                assert seg["counts"] == str(G[pre + "counts_str"][d])                # the reference's own string


def test_im_post_full_size_properties(dev):
    """1024 x 2048 (BASELINE configs[1] image), 100 detections incl. a full-image box (runs wrap between columns), a box
    outside-clipped at every border and a 1-pixel box: run lengths sum to H*W and decode to the oracle's paste."""
    rng = np.random.default_rng(8)
    H, W, n = 1024, 2048, 100
    c = np.stack([rng.uniform(0, W, n), rng.uniform(0, H, n)], 1); s = np.exp(rng.uniform(np.log(8), np.log(700), (n, 2)))
    b = np.concatenate([c - s / 2, c + s / 2], 1).astype(np.float32)
    b[:, 0::2] = np.clip(b[:, 0::2], 0, W - 1); b[:, 1::2] = np.clip(b[:, 1::2], 0, H - 1)
    b[0] = [0, 0, W - 1, H - 1]; b[1] = [0, 0, 30, H - 1]; b[2] = [W - 2, H - 2, W - 1, H - 1]; b[3] = [100.2, 200.7, 100.4, 200.9]
    yy, xx = np.mgrid[0:28, 0:28].astype(np.float32)
    masks = np.zeros((n, 9, 28, 28), np.float32)
    cls = rng.integers(1, 9, n).astype(np.int64)
    for i in range(n):
        cx, cy, r = rng.uniform(8, 20), rng.uniform(8, 20), rng.uniform(5, 16)
        masks[i, cls[i]] = np.clip(1.0 / (1.0 + np.exp(((xx - cx) ** 2 + (yy - cy) ** 2 - r * r) / 18.0)) + rng.normal(0, 0.1, (28, 28)), 0, 1)
    masks[0, cls[0]] = 0.8; masks[1, cls[1]] = 0.8
    cn, rl, ovf = _run(dev, b, masks, cls, H, W)
    assert ovf == 0 and (rl > 0).all()
    for d in range(n):
        assert int(cn[d, :rl[d]].astype(np.int64).sum()) == H * W
    for d in (0, 1, 2, 3, 7, 42, 99):
        want = O.im_post_masks(b[d:d + 1], masks[d:d + 1], cls[d:d + 1], H, W)[0]
        assert np.array_equal(cn[d, :rl[d]], O.rle_counts(want)), d


def test_im_post_counts_n_dev_overflow_and_limits(dev):
    import ctypes as C
    from upsnet_b200._lib import lib
    pre = "c2_"
    H, W = (int(v) for v in G[pre + "hw"])
    boxes, masks, cls = G[pre + "boxes"], G[pre + "masks"], G[pre + "cls"]
    nd = torch.tensor([5], dtype=torch.int32, device=dev)
    cn, rl, ovf = _run(dev, boxes, masks, cls, H, W, n_dev=nd)
    assert (rl[5:] == 0).all() and (rl[:5] > 0).all() and ovf == 0
    cn2, rl2, ovf2 = _run(dev, boxes, masks, cls, H, W, cap=4)               # far too small: flagged, sizes still reported
    full = _run(dev, boxes, masks, cls, H, W)[1]
    assert ovf2 == 1 and np.array_equal(rl2, full)
    nb = C.c_size_t(0)
    assert lib().upsnet_im_post_workspace_bytes(10, 64, C.byref(nb)) == 0 and nb.value >= 10 * 64 * 8
    assert lib().upsnet_im_post_workspace_bytes(10, 1, C.byref(nb)) == -1
