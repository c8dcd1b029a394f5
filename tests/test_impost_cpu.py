"""Row f4 (im_post: mask paste + COCO RLE, upsnet_end2end_test.py:95-152), CPU side: the numpy oracle against fixtures made
by executing the reference's own im_post with real cv2 (tests/golden/make_reference_impost.py), and the RLE codec
(restated from pycocotools' maskApi.c) against known-answer vectors and round trips."""
import os

import numpy as np
import pytest

from oracle import oracle as O

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_impost.npz"))


@pytest.mark.parametrize("ci", [0, 1, 2])
def test_oracle_paste_matches_reference_im_post(ci):
    pre = "c%d_" % ci
    H, W = (int(v) for v in G[pre + "hw"])
    masks = G[pre + "masks"]
    cls = G[pre + "cls"] if masks.shape[1] > 1 else np.zeros(masks.shape[0], np.int64)
    got = O.im_post_masks(G[pre + "boxes"], masks, cls, H, W)
    want = G[pre + "images"]
    diff = int((got != want).sum())
    # cv2's SIMD bilinear kernel rounds differently from its documented formula in the last ulp: a pixel whose resized value
    # sits within ~1e-6 of the 0.5 threshold may flip; everything else is identical
    assert diff <= 2, diff
    if diff:
        ref = O.expand_boxes(G[pre + "boxes"], 30.0 / 28.0).astype(np.int32)
        for d, y, x in zip(*np.nonzero(got != want)):
            P = np.zeros((30, 30), np.float32); P[1:-1, 1:-1] = masks[d, int(cls[d])]
            w = max(ref[d, 2] - ref[d, 0] + 1, 1); h = max(ref[d, 3] - ref[d, 1] + 1, 1)
            v = O.resize_linear(P, int(w), int(h))[y - ref[d, 1], x - ref[d, 0]]
            assert abs(float(v) - 0.5) < 5e-6
    # with real cv2 as the resize the restatement is bit-identical to the reference-executed images
    got_cv2 = O.im_post_masks(G[pre + "boxes"], masks, cls, H, W, resize="cv2")
    assert np.array_equal(got_cv2, want)


@pytest.mark.parametrize("ci", [0, 1, 2])
def test_rle_codec_on_reference_images(ci):
    pre = "c%d_" % ci
    H, W = (int(v) for v in G[pre + "hw"])
    for img, s in zip(G[pre + "images"], G[pre + "counts_str"]):
        c = O.rle_counts(img)
        assert int(c.sum()) == H * W and O.rle_to_string(c).decode() == str(s)
        assert np.array_equal(O.rle_from_string(str(s)), c)
        assert np.array_equal(O.rle_decode(c, H, W), img)


def test_rle_known_answers():
    # column-major runs, leading zeros run (possibly empty): maskApi.c rleEncode
    assert O.rle_counts(np.array([[0, 1], [1, 1]], np.uint8)).tolist() == [1, 3]
    assert O.rle_counts(np.array([[1, 0], [1, 0]], np.uint8)).tolist() == [0, 2, 2]
    assert O.rle_counts(np.zeros((3, 5), np.uint8)).tolist() == [15]
    assert O.rle_counts(np.ones((3, 5), np.uint8)).tolist() == [0, 15]
    # rleToString: 5 data bits per char + continuation bit, offset 48; counts past the third are differences to the count two back
    assert O.rle_to_string([1, 3]) == b"13"
    assert O.rle_to_string([0, 15]) == b"0?"
    assert O.rle_to_string([40]) == b"X1"                          # 40 = 0b01000 | 1 << 5
    assert O.rle_to_string([5, 2, 9, 2, 9]) == b"5290" + b"0"      # 4th: 2 - 2 = 0, 5th: 9 - 9 = 0
    assert O.rle_from_string(b"5290" + b"0").tolist() == [5, 2, 9, 2, 9]
    assert O.rle_to_string([3, 7, 1]) == b"371"
    neg = O.rle_to_string([6, 1, 6, 9, 2])                         # 5th count: 2 - 6 = -4 -> sign-extended 5-bit group
    assert O.rle_from_string(neg).tolist() == [6, 1, 6, 9, 2]
    rng = np.random.default_rng(3)
    for _ in range(20):
        h, w = int(rng.integers(1, 40)), int(rng.integers(1, 40))
        m = (rng.uniform(0, 1, (h, w)) > rng.uniform(0.2, 0.8)).astype(np.uint8)
        c = O.rle_counts(m)
        assert np.array_equal(O.rle_decode(O.rle_from_string(O.rle_to_string(c)), h, w), m)


def test_operators_rle_to_string_matches_oracle_codec():
    """The host half of operators.im_post (rleToString over the device's run lengths) is the same codec as the oracle's."""
    from upsnet_b200 import operators as ops
    rng = np.random.default_rng(4)
    for _ in range(30):
        h, w = int(rng.integers(1, 60)), int(rng.integers(1, 60))
        m = (rng.uniform(0, 1, (h, w)) > rng.uniform(0.1, 0.9)).astype(np.uint8)
        c = O.rle_counts(m)
        assert ops.rle_to_string(c) == O.rle_to_string(c)
    assert ops.rle_to_string([0, 15]) == b"0?" and ops.rle_to_string([40]) == b"X1"


def test_workspaces_are_per_engine_lane():
    """operators._Workspace: one scratch buffer per (device, engine lane) -- two lanes running concurrently must never share
    scratch (model._run_static sets WS_SLOT around a lane's work); an outgrown buffer is retired, not freed (captured graphs
    keep its pointer)."""
    import torch
    from upsnet_b200 import operators as ops
    ws = ops._Workspace()
    dev = torch.device("cpu")
    was = ops.WS_SLOT["i"]
    try:
        ops.WS_SLOT["i"] = 0
        a = ws.get(dev, 100)
        assert ws.get(dev, 50) is a
        ops.WS_SLOT["i"] = 1
        b = ws.get(dev, 100)
        assert b is not a and b.data_ptr() != a.data_ptr()
        ops.WS_SLOT["i"] = 0
        a2 = ws.get(dev, 1000)
        assert a2 is not a and a2.numel() >= 1000 and any(r is a for r in ws.retired)
        ops.WS_SLOT["i"] = 1
        assert ws.get(dev, 10) is b
    finally:
        ops.WS_SLOT["i"] = was


def test_workspace_lane_is_thread_local():
    import threading
    from upsnet_b200 import operators as ops
    seen = {}

    def worker():
        seen["initial"] = ops.WS_SLOT["i"]
        ops.WS_SLOT["i"] = 3
        seen["set"] = ops.WS_SLOT["i"]
    was = ops.WS_SLOT["i"]
    ops.WS_SLOT["i"] = 1
    try:
        th = threading.Thread(target=worker); th.start(); th.join()
        assert seen == {"initial": 0, "set": 3} and ops.WS_SLOT["i"] == 1
    finally:
        ops.WS_SLOT["i"] = was
