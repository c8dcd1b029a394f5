"""Fixtures for row f4: the reference's OWN `im_post` (upsnet_end2end_test.py:95-152) executed in the build container.

The function body is compiled from the reference file by AST, unmodified, with these globals: numpy, real cv2, the
reference's `expand_boxes` (bbox/bbox_transform.py:365-381, compiled from its file the same way), a `config` stub carrying
network.mask_size, and `mask_encode` = the numpy restatement of pycocotools.mask.encode (oracle.mask_encode: pycocotools is
not in this image and the reference does not vendor it) wrapped so that the pasted [H,W] images are recorded too.
What the fixture therefore pins to reference-executed code: box expansion + int32 truncation, zero padding, cv2.resize,
the 0.5 threshold, border clipping, per-class grouping / ordering.  The RLE codec stays pinned to its published algorithm.

Run (build container only):  python tests/golden/make_reference_impost.py   ->  tests/golden/reference_impost.npz
"""
import ast
import os
import sys
import types

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import oracle as O  # noqa: E402

REF = "/root/reference/upsnet"


def _fn(path, name):
    src = open(path).read()
    return [n for n in ast.walk(ast.parse(src)) if isinstance(n, ast.FunctionDef) and n.name == name][0]


def reference_im_post():
    ns = {"np": np}
    exec(compile(ast.Module(body=[_fn(REF + "/bbox/bbox_transform.py", "expand_boxes")], type_ignores=[]),
                 "bbox_transform.py:365-381", "exec"), ns)
    recorded = []

    def mask_encode(arr):              # pycocotools.mask.encode signature: [H,W,1] Fortran uint8 -> list of dicts
        recorded.append(np.array(arr[:, :, 0]))
        return [O.mask_encode(arr)]
    config = types.SimpleNamespace(network=types.SimpleNamespace(mask_size=28))
    ns2 = {"np": np, "cv2": cv2, "config": config, "expand_boxes": ns["expand_boxes"], "mask_encode": mask_encode}
    exec(compile(ast.Module(body=[_fn(REF + "/upsnet_end2end_test.py", "im_post")], type_ignores=[]),
                 "upsnet_end2end_test.py:95-152", "exec"), ns2)
    return ns2["im_post"], recorded


def case(rng, n, H, W, C, smin, smax, edge=False):
    c = np.stack([rng.uniform(0, W, n), rng.uniform(0, H, n)], 1)
    s = rng.uniform(smin, smax, (n, 2))
    b = np.concatenate([c - s / 2, c + s / 2], 1).astype(np.float32)
    b[:, 0::2] = np.clip(b[:, 0::2], 0, W - 1); b[:, 1::2] = np.clip(b[:, 1::2], 0, H - 1)
    if edge and n >= 3:
        b[0] = [0, 0, W - 1, H - 1]                       # full image: runs wrap from column to column
        b[1] = [W - 6, H - 5, W - 1, H - 1]               # corner box, clipped expansion
        b[2] = [3.4, 0, 3.6, H - 1]                       # sliver, full height
    # smooth blobs + noise so that masks have several runs per column
    yy, xx = np.mgrid[0:28, 0:28].astype(np.float32)
    masks = np.zeros((n, C, 28, 28), np.float32)
    for i in range(n):
        for k in range(C):
            cx, cy, r = rng.uniform(8, 20), rng.uniform(8, 20), rng.uniform(5, 16)
            blob = 1.0 / (1.0 + np.exp(((xx - cx) ** 2 + (yy - cy) ** 2 - r * r) / 18.0))
            masks[i, k] = np.clip(blob + rng.normal(0, 0.12, (28, 28)), 0, 1)
    if edge and n >= 3:
        masks[0] = 0.9                                   # all ones inside the pad border
    cls = rng.integers(1, C, n).astype(np.int64) if C > 1 else np.zeros(n, np.int64)
    scores = rng.uniform(0.3, 1.0, n).astype(np.float32)
    return b, masks, cls, scores


def main():
    im_post, recorded = reference_im_post()
    rng = np.random.default_rng(2024)
    out = {}
    cases = [dict(n=7, H=96, W=128, C=9, smin=8, smax=70, edge=True), dict(n=5, H=64, W=48, C=1, smin=4, smax=40, edge=False),
             dict(n=12, H=120, W=200, C=9, smin=3, smax=120, edge=True)]
    for ci, c in enumerate(cases):
        b, masks, cls, scores = case(rng, c["n"], c["H"], c["W"], c["C"], c["smin"], c["smax"], c["edge"])
        ncls = 9
        boxes_all = [[] for _ in range(ncls)]
        masks_all = [[] for _ in range(ncls)]
        del recorded[:]
        cls_use = cls if c["C"] > 1 else np.ones(c["n"], np.int64)        # C == 1: class-agnostic masks, every detection class 1
        im_post(boxes_all, masks_all, scores, b, masks, cls_use, ncls, (c["H"], c["W"]))
        # recorded images come class-major (idx = 1..), inside a class in detection order
        order = [d for idx in range(1, ncls) for d in np.flatnonzero(cls_use == idx)]
        imgs = np.zeros((c["n"], c["H"], c["W"]), np.uint8)
        for img, d in zip(recorded, order):
            imgs[d] = img
        strings = [""] * c["n"]
        for idx in range(1, ncls):
            for d, seg in zip(np.flatnonzero(cls_use == idx), masks_all[idx][0]):
                strings[d] = seg["counts"]
        pre = "c%d_" % ci
        out[pre + "boxes"] = b; out[pre + "masks"] = masks; out[pre + "cls"] = cls_use; out[pre + "scores"] = scores
        out[pre + "hw"] = np.array([c["H"], c["W"]]); out[pre + "images"] = imgs
        out[pre + "counts_str"] = np.array(strings)
        out[pre + "cls_boxes_1"] = boxes_all[1][0]
    np.savez_compressed(os.path.join(HERE, "reference_impost.npz"), **out)
    print("wrote reference_impost.npz:", {k: v.shape for k, v in out.items() if k.endswith("images")})


if __name__ == "__main__":
    main()
