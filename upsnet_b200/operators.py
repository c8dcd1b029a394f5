"""Host-side mirror of the reference's operator API (upsnet/operators/modules/*, upsnet/nms/nms.py)
on top of the sm_100a C ABI.  Same class names, constructor arguments, parameter names/shapes
(so reference checkpoints load) and forward signatures; all compute is in libupsnet_b200.so.

Reference interfaces mirrored (paths relative to /root/reference/upsnet/):
  DeformConv, DeformConvWithOffset      operators/modules/deform_conv.py:27-78
  ModDeformConv(+WithOffsetMask)        operators/modules/mod_deform_conv.py:24-81  (exported as
                                        ModulatedDeformConv too, the name BASELINE.json uses)
  RoIAlign / RoIAlignFunction           operators/modules/roialign.py:20-29, functions/roialign.py:21-43
  FPNRoIAlign                           operators/modules/fpn_roi_align.py:22-62
  gpu_nms_wrapper / nms                 nms/nms.py:43-46, nms/gpu_nms.pyx:23-38
  MaskRemoval, SegTerm, PanopticHead    operators/modules/mask_removal.py:23-93,
                                        operators/modules/unary_logits.py:69-105,
                                        models/resnet_upsnet.py:217-247 (F1: no such class there)
Forward only: the inference hot path; backward kernels are the training
config and out of this round's scope.
"""
import ctypes as C
import math

import numpy as np
import torch
import torch.nn as nn
from torch.nn.modules.utils import _pair
from torch.nn.parameter import Parameter

from . import _lib
from ._lib import check, f32c, lib, ptr, require_cuda, stream_ptr

# default arithmetic of the convolution tiles; switched by upsnet_b200.set_precision()
_PRECISION = {"conv": _lib.PREC_FP32_SIMT}
ACT_BF16 = {"on": False}   # engine switch: store activations as bf16 (precision 'bf16' only)
ACT_PAIR = {"on": False}   # engine switch: store activations as hi/lo bf16 pairs (precision 'bf16x3': fp32-grade results)
# pair-stream deformable convs gather from a shared-memory window with the A operand in TMEM (csrc/dcn_win.cu); maps with fewer
# than min_pixels output pixels keep the global-gather kernel (measured: 0.058 vs 0.048 ms at 32x64, 0.060 vs 0.075 ms at 64x128)
DCN_WINDOW = {"on": True, "min_pixels": 4096}
USE_TMA = {"on": True}     # False forces the cp.async gather kernel where the TMA-fed one would qualify (A/B tests)


def set_precision(name, bf16_activations=None, pair_activations=None):
    """'fp32' (CUDA-core fp32 tiles), 'bf16x3' or 'bf16' (tcgen05 tiles).  With 'bf16' the engine also stores
    the NHWC activation stream as bf16 (halves HBM traffic; the gather becomes a cp.async copy) unless
    bf16_activations=False.  With 'bf16x3' -- the configuration that meets "fp32 logits within 1e-3" -- the stream is
    stored as hi/lo bf16 PAIRS (class Pair: same bytes as fp32, ~16 mantissa bits) so that the TMA-fed tcgen05 kernel can
    run the three-term split (hi*hi + lo*hi + hi*lo) without any gather threads; pair_activations=False keeps fp32
    activations and the gather-fed kernel (round-1 behaviour, kept for A/B tests)."""
    _PRECISION["conv"] = {"fp32": _lib.PREC_FP32_SIMT, "bf16x3": _lib.PREC_BF16X3, "bf16": _lib.PREC_BF16}[name]
    ACT_BF16["on"] = (name == "bf16") if bf16_activations is None else (bool(bf16_activations) and name == "bf16")
    ACT_PAIR["on"] = (name == "bf16x3") if pair_activations is None else (bool(pair_activations) and name == "bf16x3")


class Pair:
    """An activation stored as a hi/lo bf16 pair: `store` is bf16 [N,H,W,2C] (NHWC), channels [0,C) = bf16(v) and
    [C,2C) = bf16(v - hi); v = hi + lo is exact in fp32 (UPSNET_DTYPE_PAIR in include/upsnet_b200.h).  Quacks like the
    logical [N,C,H,W] tensor for the few attributes the engine's module code reads."""
    __slots__ = ("store",)
    dtype = "pair"

    def __init__(self, store):
        assert store.dtype == torch.bfloat16 and store.dim() == 4 and store.shape[-1] % 2 == 0 and store.is_contiguous()
        self.store = store

    @staticmethod
    def from_float(x):
        """x: logical [N,C,H,W] float tensor (any memory format) -> Pair."""
        xs = x.float().permute(0, 2, 3, 1)
        hi = xs.to(torch.bfloat16)
        lo = (xs - hi.float()).to(torch.bfloat16)
        return Pair(torch.cat([hi, lo], dim=-1).contiguous())

    @property
    def shape(self):
        n, h, w, c2 = self.store.shape
        return torch.Size((n, c2 // 2, h, w))

    @property
    def device(self):
        return self.store.device

    @property
    def is_cuda(self):
        return self.store.is_cuda

    def dim(self):
        return 4

    def size(self, i=None):
        return self.shape if i is None else self.shape[i]

    def float(self):
        """fp32 logical [N,C,H,W] (channels_last storage) = hi + lo."""
        c = self.store.shape[-1] // 2
        return (self.store[..., :c].float() + self.store[..., c:].float()).permute(0, 3, 1, 2)

    def record_stream(self, s):
        self.store.record_stream(s)

    def subsample2(self):
        """x[:, :, ::2, ::2] (nn.MaxPool2d(kernel_size=1, stride=2): models/fpn.py:75)."""
        return Pair(self.store[:, ::2, ::2, :].contiguous())


def subsample2(x):
    return x.subsample2() if isinstance(x, Pair) else x[:, :, ::2, ::2].contiguous()


def as_float(x):
    """A Pair becomes its fp32 tensor (hi + lo); plain tensors pass through untouched (no dtype round trip)."""
    return x.float() if isinstance(x, Pair) else x


# launch accounting (bench.py reports gpu_launches) and optional per-call CUDA-event timing of the
# kernels behind one C-ABI call (bench.py's roofline leg; off in normal operation)
STATS = {"launches": 0, "trace": None}


class _Timed:
    """with _Timed(kind, n_kernels, work, device): <C-ABI call> -- counts launches; when
    STATS["trace"] is a list also brackets the call with events on the current stream."""

    def __init__(self, kind, n_kernels, work, device):
        self.kind, self.n, self.work, self.device = kind, n_kernels, work, device
        self.ev = None

    def __enter__(self):
        STATS["launches"] += self.n
        if STATS["trace"] is not None:
            self.ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
            self.ev[0].record(torch.cuda.current_stream(self.device))
        return self

    def __exit__(self, *exc):
        if self.ev is not None:
            self.ev[1].record(torch.cuda.current_stream(self.device))
            STATS["trace"].append((self.kind, self.ev[0], self.ev[1], self.work))
        return False


def _conv_out(n, pad, dil, k, stride):
    return (n + 2 * pad - (dil * (k - 1) + 1)) // stride + 1


# ------------------------------------------------------------------------------------------------
# tensor-core (tcgen05) path plumbing: NHWC views and the packed-weight cache
# ------------------------------------------------------------------------------------------------
def _nhwc(x):
    """[N,C,H,W] logical tensor -> contiguous [N,H,W,C] storage (no copy if already channels_last)."""
    xp = x.permute(0, 2, 3, 1)
    return xp if xp.is_contiguous() else xp.contiguous()


import weakref

# keyed on the weight TENSOR OBJECT (weak): a data_ptr key would alias a freed weight whose storage the
# caching allocator handed to a new tensor.  Entries die with their tensor; _version catches in-place updates.
_packed_cache = {}   # id(tensor) -> (weakref to the tensor, version, packed buffer)


def _packed_weight(weight):
    """bf16 hi/lo planes [Cout_pad][kh*kw][Cin] (upsnet_igemm_pack_weight), cached per weight tensor+version."""
    hit = _packed_cache.get(id(weight))
    if hit is not None and hit[0]() is weight and hit[1] == weight._version and hit[2].device == weight.device:
        return hit[2]
    Cout, Cin, kh, kw = weight.shape
    nbytes = C.c_size_t(0)
    check(lib().upsnet_igemm_packed_weight_bytes(Cout, Cin, kh, kw, C.byref(nbytes)), "igemm_packed_weight_bytes")
    buf = torch.empty(nbytes.value, dtype=torch.uint8, device=weight.device)
    w = f32c(weight.detach())
    with torch.cuda.device(weight.device):
        check(lib().upsnet_igemm_pack_weight(ptr(w), Cout, Cin, kh, kw, ptr(buf), stream_ptr(weight.device)),
              "igemm_pack_weight")
    STATS["launches"] += 1
    wid = id(weight)
    _packed_cache[wid] = (weakref.ref(weight, lambda _r, _k=wid: _packed_cache.pop(_k, None)), weight._version, buf)
    return buf


_dcn_packed_cache = {}


def _packed_weight_dcn(weight):
    """upsnet_dcn_pack_weight: bf16 hi/lo planes in the window kernel's K order (16-channel sub-chunk, tap, channel);
    None when the layer shape is not supported by that kernel.  Cached like _packed_weight."""
    hit = _dcn_packed_cache.get(id(weight))
    if hit is not None and hit[0]() is weight and hit[1] == weight._version and (hit[2] is None or hit[2].device == weight.device):
        return hit[2]
    Cout, Cin, kh, kw = weight.shape
    nbytes = C.c_size_t(0)
    rc = lib().upsnet_dcn_packed_weight_bytes(Cout, Cin, kh, kw, C.byref(nbytes))
    buf = None
    if rc == 0:
        buf = torch.empty(nbytes.value, dtype=torch.uint8, device=weight.device)
        w = f32c(weight.detach())
        with torch.cuda.device(weight.device):
            check(lib().upsnet_dcn_pack_weight(ptr(w), Cout, Cin, kh, kw, ptr(buf), stream_ptr(weight.device)), "dcn_pack_weight")
        STATS["launches"] += 1
    elif rc != -2:
        check(rc, "dcn_packed_weight_bytes")
    wid = id(weight)
    _dcn_packed_cache[wid] = (weakref.ref(weight, lambda _r, _k=wid: _dcn_packed_cache.pop(_k, None)), weight._version, buf)
    return buf


# small-N dense 3x3 / stride-1 convs on the pair stream through the window pipeline (csrc/dcn_win.cu DENSE mode): Cout <= max_cout
# measured on B200 (scripts/exp_offset_conv.py): 0.18 vs 0.15 ms for the 256->18 offset conv at 256x512 -- the 32-byte rows of the
# window boxes run into the TMA request rate (~4-5 cycles per box row and SM) -> off by default, kept for the record / tests
DENSE_WINDOW = {"on": False, "max_cout": 64, "min_pixels": 4096}


def _dense_window(x, weight, bias, padding, dilation, relu, nchw_out):
    """upsnet_conv3x3_pair_forward: Pair in; fp32 NCHW (offset maps) or Pair out.  None when the layer does not qualify."""
    N, Cin, H, W = x.shape
    Cout, _, kh, kw = weight.shape
    ph, pw = padding; dh, dw = dilation
    if N * H * W < DENSE_WINDOW["min_pixels"] or Cout > DENSE_WINDOW["max_cout"] or (not nchw_out and Cout % 16):
        return None
    packed = _packed_weight_dcn(weight)
    if packed is None:
        return None
    Ho, Wo = _conv_out(H, ph, dh, 3, 1), _conv_out(W, pw, dw, 3, 1)
    if nchw_out:
        store = torch.empty((N, Cout, Ho, Wo), device=x.device, dtype=torch.float32)
    else:
        store = torch.empty((N, Ho, Wo, 2 * Cout), device=x.device, dtype=torch.bfloat16)
    work = {"flops": 2.0 * N * Ho * Wo * Cout * Cin * 9 * 3, "algo_flops": 2.0 * N * Ho * Wo * Cout * Cin * 9,
            "shape": "N%d %dx%d Cin%d->Cout%d k3 s1 pair->%s (window)" % (N, H, W, Cin, Cout, "float32" if nchw_out else "pair"),
            "bytes": float(x.store.numel() * 2 + 4 * weight.numel() + store.numel() * store.element_size())}
    with torch.cuda.device(x.device), _Timed("conv2d", 1, work, x.device):
        rc = lib().upsnet_conv3x3_pair_forward(ptr(x.store), ptr(packed), ptr(bias), ptr(store), N, H, W, Cin, Cout, ph, pw, dh, dw,
                                               _lib.LAYOUT_NCHW if nchw_out else _lib.LAYOUT_NHWC,
                                               _lib.EPI_RELU if relu else 0, stream_ptr(x.device))
    if rc == -2:
        return None
    check(rc, "conv3x3_pair_forward")
    return store if nchw_out else Pair(store)


def _dcn_window(x, offset, mask, weight, bias, padding, dilation, relu):
    """upsnet_dcn_pair_forward (csrc/dcn_win.cu): Pair in, Pair out, 3x3 / stride 1.  Returns None when the layer does
    not qualify (the caller then takes upsnet_igemm_forward)."""
    N, Cin, H, W = x.shape
    if N * H * W < DCN_WINDOW["min_pixels"]:
        return None
    packed = _packed_weight_dcn(weight)
    if packed is None:
        return None
    Cout, _, kh, kw = weight.shape
    ph, pw = padding; dh, dw = dilation
    Ho, Wo = _conv_out(H, ph, dh, kh, 1), _conv_out(W, pw, dw, kw, 1)
    store = torch.empty((N, Ho, Wo, 2 * Cout), device=x.device, dtype=torch.bfloat16)
    work = {"flops": 2.0 * N * Ho * Wo * Cout * Cin * kh * kw * 3, "algo_flops": 2.0 * N * Ho * Wo * Cout * Cin * kh * kw,
            "shape": "N%d %dx%d Cin%d->Cout%d k%d s1 pair->pair (window)" % (N, H, W, Cin, Cout, kh),
            "bytes": float(x.store.numel() * 2 + 4 * weight.numel() + store.numel() * 2 + offset.numel() * 4)}
    with torch.cuda.device(x.device), _Timed("dcn", 1, work, x.device):
        rc = lib().upsnet_dcn_pair_forward(ptr(x.store), ptr(offset), ptr(mask), ptr(packed), ptr(bias), ptr(store), N, H, W,
                                           Cin, Cout, kh, kw, ph, pw, dh, dw, _lib.EPI_RELU if relu else 0, stream_ptr(x.device))
    if rc == -2:
        return None
    check(rc, "dcn_pair_forward")
    return Pair(store)


_stem_cache = {}   # id(weight) -> (weakref, version, packed bf16 [Cout][kh][8][8])
_stem_ws = None


def stem_conv(x, weight, bias, padding, relu=True, pair=False):
    """k x k / stride-2 convolution of a tiny-Cin fp32 NCHW image (models/resnet.py:155-162 conv1) on the TMA kernel:
    the image is packed to a zero-padded bf16 NHWC8 copy whose 5-D tensor-map boxes are the im2col tiles.
    -> bf16 channels_last activation [N,Cout,Ho,Wo]; pair=True (precision bf16x3): hi and lo copies of the image, three MMAs
    per k-slice, -> Pair.  Raises UpsnetError(UNSUPPORTED) if the driver rejects the map."""
    global _stem_ws
    require_cuda(x, weight, bias)
    x = f32c(x)
    N, Cin, H, W = x.shape
    Cout, _, kh, kw = weight.shape
    dev = x.device
    hit = _stem_cache.get(id(weight))
    if hit is not None and hit[0]() is weight and hit[1] == weight._version and hit[2].device == dev:
        packed = hit[2]
    else:
        nb = C.c_size_t(0)
        check(lib().upsnet_stem_packed_weight_bytes(Cout, kh, C.byref(nb)), "stem_packed_weight_bytes")
        packed = torch.empty(nb.value, dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            check(lib().upsnet_stem_pack_weight(ptr(f32c(weight.detach())), Cout, Cin, kh, kw, ptr(packed), stream_ptr(dev)),
                  "stem_pack_weight")
        STATS["launches"] += 1
        wid = id(weight)
        _stem_cache[wid] = (weakref.ref(weight, lambda _r, _k=wid: _stem_cache.pop(_k, None)), weight._version, packed)
    nb = C.c_size_t(0)
    check(lib().upsnet_stem_workspace_bytes(N, H, W, kh, kw, int(padding), C.byref(nb)), "stem_workspace_bytes")
    if _stem_ws is None:
        _stem_ws = _Workspace()
    ws = _stem_ws.get(dev, nb.value)
    Ho, Wo = (H + 2 * padding - kh) // 2 + 1, (W + 2 * padding - kw) // 2 + 1
    store = torch.empty((N, Ho, Wo, Cout * (2 if pair else 1)), dtype=torch.bfloat16, device=dev)
    fl = 2.0 * N * Ho * Wo * Cout * Cin * kh * kw
    work = {"flops": fl * (3 if pair else 1), "algo_flops": fl,
            "shape": "N%d %dx%d Cin%d->Cout%d k%d s2 float32->%s (stem, TMA)" % (N, H, W, Cin, Cout, kh, "pair" if pair else "bfloat16"),
            "bytes": float(4 * x.numel() + 2 * store.numel())}
    with torch.cuda.device(dev), _Timed("conv2d", 2, work, dev):
        check(lib().upsnet_stem_forward(ptr(x), ptr(packed), ptr(None if bias is None else f32c(bias)), ptr(store), N, Cin, H, W,
                                        Cout, kh, kw, int(padding), (_lib.EPI_RELU if relu else 0) | (_lib.EPI_STEM_PAIR if pair else 0),
                                        ptr(ws), ws.numel(), stream_ptr(dev)), "stem_forward")
    return Pair(store) if pair else store.permute(0, 3, 1, 2)


def _tc_ok(Cin, kh, kw, dg, deform=False):
    return (Cin % 64 == 0 or (Cin <= 8 and not deform)) and dg == 1 and kh * kw <= 49


def _igemm_tc(kind, x, offset, mask, weight, bias, residual, stride, padding, dilation, relu, prec, out_format,
              out_dtype=None, residual_up2=False, pair_group=0, sigmoid_from=None):
    """upsnet_igemm_forward: x logical NCHW (any memory format; fp32 or bf16) or a Pair; result logical NCHW whose
    storage is NHWC (channels_last view, the engine layout) unless out_format == 'nchw'.
    Output: bf16 when the engine stores bf16 activations (ACT_BF16) / a Pair when it stores pairs (ACT_PAIR) and the
    result stays in the NHWC activation stream; fp32 for plane-wise (NCHW) head outputs or when asked via out_dtype.
    pair_group=G (Pair output only): channels are written as [hi G][lo G] groups and the result is returned as the Pair
    of logical shape [N, G, Ho, Wo*Cout/G] that this storage also is (see MaskBranch)."""
    sh, sw = stride; ph, pw = padding; dh, dw = dilation
    N, Cin, H, W = x.shape
    Cout, _, kh, kw = weight.shape
    Ho, Wo = _conv_out(H, ph, dh, kh, sh), _conv_out(W, pw, dw, kw, sw)
    nhwc_out = out_format != "nchw"
    x3 = prec == _lib.PREC_BF16X3
    pair_in = isinstance(x, Pair)
    if pair_in:
        assert x3, "Pair activations belong to precision bf16x3"
    elif x3 and ACT_PAIR["on"] and Cin % 64 == 0:
        x = Pair.from_float(x)                               # entry into the pair stream (API-level callers)
        pair_in = True
    elif x3 and x.dtype != torch.float32:
        x = x.float()                                        # the in-kernel hi/lo split needs fp32 activations
    elif x.dtype not in (torch.float32, torch.bfloat16):
        x = x.float()
    dev = x.device
    pair_out = False
    if out_dtype is None:
        if x3 and ACT_PAIR["on"] and nhwc_out and Cout % 8 == 0:
            pair_out = True
        else:
            out_dtype = torch.bfloat16 if (ACT_BF16["on"] and prec == _lib.PREC_BF16 and nhwc_out) else torch.float32
    elif out_dtype == "pair":
        pair_out = True
    if pair_in:
        xs, x_dt = x.store, _lib.DTYPE_PAIR
    else:
        xs = f32c(x) if Cin % 64 else _nhwc(x)      # tiny-Cin (stem) mode reads the NCHW fp32 image directly
        x_dt = _lib.DTYPE_BF16 if xs.dtype == torch.bfloat16 else _lib.DTYPE_F32
    packed = _packed_weight(weight)
    if pair_out:
        assert nhwc_out
        store = torch.empty((N, Ho, Wo, 2 * Cout), device=dev, dtype=torch.bfloat16)
        res = None
        if residual is not None:
            res = (residual if isinstance(residual, Pair) else Pair.from_float(residual)).store
        y_dt = _lib.DTYPE_PAIR
    elif nhwc_out:
        store = torch.empty((N, Ho, Wo, Cout), device=dev, dtype=out_dtype)
        res = None if residual is None else _nhwc(as_float(residual).to(out_dtype))
        y_dt = _lib.DTYPE_BF16 if out_dtype == torch.bfloat16 else _lib.DTYPE_F32
    else:
        store = torch.empty((N, Cout, Ho, Wo), device=dev, dtype=out_dtype)
        res = None if residual is None else as_float(residual).to(out_dtype).contiguous()
        y_dt = _lib.DTYPE_BF16 if out_dtype == torch.bfloat16 else _lib.DTYPE_F32
    xbytes = xs.numel() * xs.element_size()
    work = {"flops": 2.0 * N * Ho * Wo * Cout * Cin * kh * kw * (3 if x3 else 1),
            "algo_flops": 2.0 * N * Ho * Wo * Cout * Cin * kh * kw,
            "shape": "N%d %dx%d Cin%d->Cout%d k%d s%d %s->%s%s" % (N, H, W, Cin, Cout, kh, sh, "pair" if pair_in else str(xs.dtype)[6:],
                                                              "pair" if pair_out else str(out_dtype)[6:],
                                                              " +res" if residual is not None else ""),
            "bytes": float(xbytes + 4 * weight.numel() +
                           store.numel() * store.element_size() * (2 if residual is not None else 1))}
    flags = (_lib.EPI_RELU if relu else 0) | (_lib.EPI_RES_UP2 if residual_up2 else 0) | \
            (0 if USE_TMA["on"] else _lib.EPI_NO_TMA)
    if pair_group:
        assert pair_out and pair_group % 64 == 0 and Cout % pair_group == 0
        flags |= (pair_group // 64) << 8
    if sigmoid_from is not None:
        assert not pair_out and not nhwc_out and residual is None, "sigmoid epilogue: fp32 NCHW head outputs"
        flags |= _lib.EPI_SIGMOID_FROM(sigmoid_from)
    with torch.cuda.device(dev), _Timed(kind, 1, work, dev):
        check(lib().upsnet_igemm_forward(ptr(xs), ptr(offset), ptr(mask), ptr(packed), ptr(bias), ptr(res),
                                         ptr(store), N, H, W, Cin, Cout, kh, kw, sh, sw, ph, pw, dh, dw,
                                         _lib.LAYOUT_NHWC if nhwc_out else _lib.LAYOUT_NCHW, x_dt, y_dt, flags,
                                         prec, stream_ptr(dev)), kind)
    if pair_out:
        if pair_group:
            return Pair(store.view(N, Ho, Wo * (Cout // pair_group), 2 * pair_group))
        return Pair(store)
    return store.permute(0, 3, 1, 2) if nhwc_out else store


# ------------------------------------------------------------------------------------------------
# functional layer
# ------------------------------------------------------------------------------------------------
def conv2d(x, weight, bias=None, stride=1, padding=0, dilation=1, residual=None, relu=False, precision=None,
           out_format=None, out_dtype=None, residual_up2=False, pair_group=0, sigmoid_from=None):
    """Dense conv + fused bias / residual / ReLU epilogue.  fp32 precision -> upsnet_conv2d_forward
    (NCHW CUDA-core tiles); bf16x3 / bf16 -> upsnet_igemm_forward (tcgen05 tiles, NHWC storage)."""
    require_cuda(x, weight, bias, residual)
    prec = _PRECISION["conv"] if precision is None else precision
    if prec != _lib.PREC_FP32_SIMT and _tc_ok(weight.shape[1], weight.shape[2], weight.shape[3], 1):
        if residual_up2 and out_format == "nchw":
            residual, residual_up2 = torch.nn.functional.interpolate(as_float(residual), scale_factor=2, mode="nearest"), False
        if (DENSE_WINDOW["on"] and isinstance(x, Pair) and prec == _lib.PREC_BF16X3 and ACT_PAIR["on"] and USE_TMA["on"] and
                weight.shape[2] == 3 and weight.shape[3] == 3 and _pair(stride) == (1, 1) and residual is None and
                sigmoid_from is None and not pair_group and
                ((out_format == "nchw" and out_dtype in (None, torch.float32)) or (out_format != "nchw" and out_dtype in (None, "pair")))):
            y = _dense_window(x, weight, None if bias is None else f32c(bias), _pair(padding), _pair(dilation), relu,
                              out_format == "nchw")
            if y is not None:
                return y
        return _igemm_tc("conv2d", x, None, None, weight, None if bias is None else f32c(bias), residual,
                         _pair(stride), _pair(padding), _pair(dilation), relu, prec, out_format, out_dtype,
                         residual_up2, pair_group, sigmoid_from)
    if sigmoid_from is not None:      # CUDA-core path: the conv entry has no sigmoid epilogue
        y = conv2d(x, weight, bias, stride, padding, dilation, residual, relu, precision, out_format, out_dtype, residual_up2)
        y[:, sigmoid_from:] = torch.sigmoid(y[:, sigmoid_from:])
        return y
    if isinstance(x, Pair):
        x = x.float()
    if isinstance(residual, Pair):
        residual = residual.float()
    if residual_up2:   # CUDA-core path: materialise the nearest-neighbour upsampling (models/fpn.py:33-34)
        residual = torch.nn.functional.interpolate(residual, scale_factor=2, mode="nearest")
    x, weight = f32c(x), f32c(weight)
    bias = None if bias is None else f32c(bias)
    residual = None if residual is None else f32c(residual)
    sh, sw = _pair(stride); ph, pw = _pair(padding); dh, dw = _pair(dilation)
    N, Cin, H, W = x.shape
    Cout, Cin_w, kh, kw = weight.shape
    assert Cin_w == Cin, "groups != 1 is not supported"
    Ho, Wo = _conv_out(H, ph, dh, kh, sh), _conv_out(W, pw, dw, kw, sw)
    y = torch.empty((N, Cout, Ho, Wo), device=x.device, dtype=torch.float32)
    if residual is not None:
        assert residual.shape == y.shape
    prec = _lib.PREC_FP32_SIMT
    work = {"flops": 2.0 * N * Ho * Wo * Cout * Cin * kh * kw, "algo_flops": 2.0 * N * Ho * Wo * Cout * Cin * kh * kw,
            "bytes": 4.0 * (x.numel() + weight.numel() + y.numel() * (2 if residual is not None else 1))}
    with torch.cuda.device(x.device), _Timed("conv2d_simt", 1, work, x.device):
        check(lib().upsnet_conv2d_forward(ptr(x), ptr(weight), ptr(bias), ptr(residual), ptr(y), N, Cin, H, W,
                                          Cout, kh, kw, sh, sw, ph, pw, dh, dw, _lib.EPI_RELU if relu else 0,
                                          prec, stream_ptr(x.device)), "conv2d")
    return y


def linear(x, weight, bias=None, relu=False, precision=None, out_dtype=None):
    """y = x @ weight.T + bias as a 1x1 convolution over N 'images' of 1x1 pixels.  x: [N,K] tensor, or a Pair of
    logical shape [N,K,1,1] (then the result is a Pair too unless out_dtype says otherwise)."""
    if isinstance(x, Pair):
        y = conv2d(x, _as_1x1(weight), bias, relu=relu, precision=precision, out_dtype=out_dtype)
        return y if isinstance(y, Pair) else y.reshape(y.shape[0], weight.shape[0])
    N, K = x.shape
    y = conv2d(x.reshape(N, K, 1, 1), _as_1x1(weight), bias, relu=relu, precision=precision, out_dtype=out_dtype)
    return y if isinstance(y, Pair) else y.reshape(N, weight.shape[0])


_view_cache = {}


def _as_1x1(weight):
    """[Cout,K] -> [Cout,K,1,1] view, cached per weight tensor so the packed-weight cache (keyed on tensor
    identity) hits on every call."""
    hit = _view_cache.get(id(weight))
    if hit is not None and hit[0]() is weight:
        return hit[1]
    v = weight.reshape(weight.shape[0], weight.shape[1], 1, 1)
    wid = id(weight)
    _view_cache[wid] = (weakref.ref(weight, lambda _r, _k=wid: _view_cache.pop(_k, None)), v)
    return v


def deform_conv(data, offset, weight, bias=None, stride=1, padding=0, dilation=1, deformable_groups=1,
                mask=None, relu=False, precision=None, out_format=None, out_dtype=None):
    """DeformConvFunction.forward (functions/deform_conv.py:26-57); with `mask` (already 2*sigmoid)
    ModDeformConvFunction.forward (functions/mod_deform_conv.py:25-59).  One fused launch."""
    require_cuda(data, offset, weight, bias, mask)
    prec = _PRECISION["conv"] if precision is None else precision
    use_tc = prec != _lib.PREC_FP32_SIMT and _tc_ok(weight.shape[1], weight.shape[2], weight.shape[3], deformable_groups, True)
    offset = f32c(offset)
    bias = None if bias is None else f32c(bias)
    mask = None if mask is None else f32c(mask)
    sh, sw = _pair(stride); ph, pw = _pair(padding); dh, dw = _pair(dilation)
    if use_tc:
        N, Cin, H, W = data.shape
        kh, kw = weight.shape[2], weight.shape[3]
        Ho, Wo = _conv_out(H, ph, dh, kh, sh), _conv_out(W, pw, dw, kw, sw)
        assert tuple(offset.shape) == (N, 2 * kh * kw, Ho, Wo), offset.shape
        if mask is not None:
            assert tuple(mask.shape) == (N, kh * kw, Ho, Wo), mask.shape
        if (DCN_WINDOW["on"] and isinstance(data, Pair) and prec == _lib.PREC_BF16X3 and (sh, sw) == (1, 1) and
                out_format != "nchw" and out_dtype in (None, "pair") and ACT_PAIR["on"]):
            y = _dcn_window(data, offset, mask, weight, bias, (ph, pw), (dh, dw), relu)
            if y is not None:
                return y
        return _igemm_tc("dcn", data, offset, mask, weight, bias, None, (sh, sw), (ph, pw), (dh, dw), relu, prec,
                         out_format, out_dtype)
    data, weight = f32c(as_float(data)), f32c(weight)
    N, Cin, H, W = data.shape
    Cout, Cin_w, kh, kw = weight.shape
    assert Cin_w == Cin, "groups != 1 is not supported (the reference ignores `groups`)"
    Ho, Wo = _conv_out(H, ph, dh, kh, sh), _conv_out(W, pw, dw, kw, sw)
    assert tuple(offset.shape) == (N, 2 * kh * kw * deformable_groups, Ho, Wo), offset.shape
    if mask is not None:
        assert tuple(mask.shape) == (N, kh * kw * deformable_groups, Ho, Wo), mask.shape
    y = torch.empty((N, Cout, Ho, Wo), device=data.device, dtype=torch.float32)
    prec = _lib.PREC_FP32_SIMT
    work = {"flops": 2.0 * N * Ho * Wo * Cout * Cin * kh * kw, "algo_flops": 2.0 * N * Ho * Wo * Cout * Cin * kh * kw,
            "bytes": 4.0 * (data.numel() + offset.numel() + weight.numel() + y.numel() +
                            (mask.numel() if mask is not None else 0))}
    with torch.cuda.device(data.device), _Timed("dcn_simt", 1, work, data.device):
        check(lib().upsnet_dcn_forward(ptr(data), ptr(offset), ptr(mask), ptr(weight), ptr(bias), ptr(y), N, Cin,
                                       H, W, Cout, kh, kw, sh, sw, ph, pw, dh, dw, deformable_groups,
                                       _lib.EPI_RELU if relu else 0, prec, stream_ptr(data.device)),
              "deform_conv")
    return y


def roi_align(features, rois, pooled_height, pooled_width, spatial_scale, sampling_ratio=2, layout="nchw"):
    require_cuda(features, rois)
    features, rois = f32c(features), f32c(rois)
    assert rois.dim() == 2 and rois.shape[1] == 5
    R = rois.shape[0]
    if layout == "nchw":
        B, Cc, H, W = features.shape
        out = torch.empty((R, Cc, pooled_height, pooled_width), device=features.device, dtype=torch.float32)
        lay = _lib.LAYOUT_NCHW
    else:
        B, H, W, Cc = features.shape
        out = torch.empty((R, pooled_height, pooled_width, Cc), device=features.device, dtype=torch.float32)
        lay = _lib.LAYOUT_NHWC
    if R == 0:
        return out
    with torch.cuda.device(features.device), _Timed("roi_align", 1, {"bytes": 4.0 * out.numel()}, features.device):
        check(lib().upsnet_roi_align_forward(ptr(features), B, Cc, H, W, lay, 0, ptr(rois), R, pooled_height,
                                             pooled_width, sampling_ratio, float(spatial_scale), ptr(out),
                                             stream_ptr(features.device)), "roi_align")
    return out


def max_pool2d(x, kernel_size, stride, padding):
    """nn.MaxPool2d on the engine's activation stream (models/resnet.py:163).  x is a logical NCHW tensor; the
    kernel works on NHWC storage (channels_last tensors are used as they are) in bf16 or fp32."""
    require_cuda(x)
    N, Cc, H, W = x.shape
    Ho = (H + 2 * padding - kernel_size) // stride + 1
    Wo = (W + 2 * padding - kernel_size) // stride + 1
    if isinstance(x, Pair):
        store = torch.empty((N, Ho, Wo, 2 * Cc), device=x.device, dtype=torch.bfloat16)
        with torch.cuda.device(x.device), _Timed("maxpool", 1, {"bytes": float((x.store.numel() + store.numel()) * 2)}, x.device):
            check(lib().upsnet_maxpool2d_nhwc(ptr(x.store), ptr(store), N, H, W, Cc, kernel_size, stride, padding,
                                              _lib.DTYPE_PAIR, stream_ptr(x.device)), "maxpool2d")
        return Pair(store)
    if x.dtype not in (torch.float32, torch.bfloat16):
        x = x.float()
    xs = _nhwc(x)
    store = torch.empty((N, Ho, Wo, Cc), device=x.device, dtype=x.dtype)
    with torch.cuda.device(x.device), _Timed("maxpool", 1, {"bytes": float((xs.numel() + store.numel()) * x.element_size())}, x.device):
        check(lib().upsnet_maxpool2d_nhwc(ptr(xs), ptr(store), N, H, W, Cc, kernel_size, stride, padding,
                                          1 if x.dtype == torch.bfloat16 else 0, stream_ptr(x.device)), "maxpool2d")
    return store.permute(0, 3, 1, 2)


def upsample_bilinear(x, factor):
    """nn.Upsample(scale_factor=factor, mode='bilinear', align_corners=False) on contiguous NCHW fp32 planes
    (models/fcn.py:88-101: the semantic logits)."""
    require_cuda(x)
    x = f32c(x)
    N, Cc, H, W = x.shape
    y = torch.empty((N, Cc, H * factor, W * factor), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device), _Timed("upsample", 1, {"bytes": 4.0 * (x.numel() + y.numel())}, x.device):
        check(lib().upsnet_upsample_bilinear_nchw(ptr(x), ptr(y), N * Cc, H, W, int(factor), stream_ptr(x.device)), "upsample")
    return y


def fcn_score_fuse(s2, s3, s4, s5):
    """s2 + up2(s3) + up4(s4) + up8(s5) on contiguous NCHW fp32 score maps (models/fcn.py:94-101 after the engine's
    score-before-upsample rewrite), one launch instead of three F.interpolate + three adds."""
    require_cuda(s2, s3, s4, s5)
    s2, s3, s4, s5 = f32c(s2), f32c(s3), f32c(s4), f32c(s5)
    N, Cc, H, W = s2.shape
    assert s3.shape == (N, Cc, H // 2, W // 2) and s4.shape == (N, Cc, H // 4, W // 4) and s5.shape == (N, Cc, H // 8, W // 8)
    out = torch.empty_like(s2)
    with torch.cuda.device(s2.device), _Timed("upsample", 1, {"bytes": 4.0 * (2 * s2.numel() + s3.numel() + s4.numel() + s5.numel())}, s2.device):
        check(lib().upsnet_fcn_score_fuse(ptr(s2), ptr(s3), ptr(s4), ptr(s5), ptr(out), N * Cc, H, W, stream_ptr(s2.device)),
              "fcn_score_fuse")
    return out


def fpn_roi_align(feats, rois, pooled_height, pooled_width, spatial_scales, sampling_ratio=2, layout="nchw",
                  return_levels=False):
    """FPNRoIAlign.forward in one launch (level assignment on device, output already in roi order)."""
    assert len(feats) == 4 and len(spatial_scales) == 4
    require_cuda(rois, *feats)
    rois = f32c(rois)
    R = rois.shape[0]
    if isinstance(feats[0], Pair):
        # hi/lo pair features -> Pair result: [R,C,PH,PW] pair pixels, or (layout 'flat_pair') the [R, PH*PW*C, 1, 1]
        # pair of the flattened (ph, pw, c) roi feature that the RCNN fc6 consumes as a 1x1 'image'
        assert all(isinstance(f, Pair) for f in feats) and layout in ("auto", "nhwc", "flat_pair") and not return_levels
        B, Cc = feats[0].shape[0], feats[0].shape[1]
        flat = layout == "flat_pair"
        if flat:
            out = torch.empty((R, 1, 1, 2 * pooled_height * pooled_width * Cc), device=rois.device, dtype=torch.bfloat16)
        else:
            out = torch.empty((R, pooled_height, pooled_width, 2 * Cc), device=rois.device, dtype=torch.bfloat16)
        if R > 0:
            fp = (C.c_void_p * 4)(*[f.store.data_ptr() for f in feats])
            hs = (C.c_int * 4)(*[f.shape[2] for f in feats]); ws = (C.c_int * 4)(*[f.shape[3] for f in feats])
            sc = (C.c_float * 4)(*[float(s_) for s_ in spatial_scales])
            with torch.cuda.device(rois.device), _Timed("roi_align_fpn", 1, {"bytes": 2.0 * out.numel()}, rois.device):
                check(lib().upsnet_roi_align_fpn_forward(fp, hs, ws, sc, B, Cc, _lib.LAYOUT_FLAT_PAIR if flat else _lib.LAYOUT_NHWC,
                                                         _lib.DTYPE_PAIR, ptr(rois), R, pooled_height, pooled_width,
                                                         sampling_ratio, ptr(out), ptr(None), stream_ptr(rois.device)),
                      "fpn_roi_align")
        return Pair(out)
    if layout == "auto":
        # engine tensors: logical NCHW; if every level is stored channels_last use the NHWC kernel
        # (coalesced channel vectors) and hand back a logical-NCHW view of the NHWC result
        cl = all(f.dim() == 4 and f.permute(0, 2, 3, 1).is_contiguous() and not f.is_contiguous() for f in feats)
        if cl:
            out = fpn_roi_align([f.permute(0, 2, 3, 1) for f in feats], rois, pooled_height, pooled_width,
                                spatial_scales, sampling_ratio, "nhwc", return_levels)
            return (out[0].permute(0, 3, 1, 2), out[1]) if return_levels else out.permute(0, 3, 1, 2)
        layout = "nchw"
    bf16 = layout == "nhwc" and all(f.dtype == torch.bfloat16 for f in feats)
    feats = [f.contiguous() for f in feats] if bf16 else [f32c(f) for f in feats]
    odt = torch.bfloat16 if bf16 else torch.float32
    if layout == "nchw":
        B, Cc = feats[0].shape[0], feats[0].shape[1]
        Hs = [f.shape[2] for f in feats]; Ws = [f.shape[3] for f in feats]
        out = torch.empty((R, Cc, pooled_height, pooled_width), device=rois.device, dtype=torch.float32)
        lay = _lib.LAYOUT_NCHW
    else:
        B, Cc = feats[0].shape[0], feats[0].shape[3]
        Hs = [f.shape[1] for f in feats]; Ws = [f.shape[2] for f in feats]
        out = torch.empty((R, pooled_height, pooled_width, Cc), device=rois.device, dtype=odt)
        lay = _lib.LAYOUT_NHWC
    levels = torch.empty((R,), device=rois.device, dtype=torch.int32) if return_levels else None
    fp = (C.c_void_p * 4)(*[f.data_ptr() for f in feats])
    hs = (C.c_int * 4)(*Hs); ws = (C.c_int * 4)(*Ws)
    sc = (C.c_float * 4)(*[float(s) for s in spatial_scales])
    with torch.cuda.device(rois.device), _Timed("roi_align_fpn", 1, {"bytes": 4.0 * out.numel()}, rois.device):
        check(lib().upsnet_roi_align_fpn_forward(fp, hs, ws, sc, B, Cc, lay, 1 if bf16 else 0, ptr(rois), R, pooled_height,
                                                 pooled_width, sampling_ratio, ptr(out), ptr(levels),
                                                 stream_ptr(rois.device)), "fpn_roi_align")
    return (out, levels) if return_levels else out


# ------------------------------------------------------------------------------------------------
# NMS
# ------------------------------------------------------------------------------------------------
import threading


class _WsSlot(threading.local):
    """Engine lane whose scratch buffers the C-ABI calls of THIS THREAD use (model._run_static / PipelinedEngine set it around
    a lane's work).  Thread-local: the reference's thread-per-GPU DataParallel usage (one Python thread per device) must not
    see another thread's lane.  Dict-style access (`WS_SLOT["i"]`) is kept for the callers."""
    i = 0

    def __getitem__(self, k):
        assert k == "i"
        return self.i

    def __setitem__(self, k, v):
        assert k == "i"
        self.i = int(v)


WS_SLOT = _WsSlot()


class _Workspace:
    """Caller-owned, grow-only device scratch (the C ABI never allocates).  A buffer that is outgrown is RETIRED, not
    freed: captured CUDA graphs (model._run_static) have its raw pointer baked in and keep writing to it on replay, so
    handing the block back to the caching allocator would let a replay corrupt whatever tensor reuses it.
    Growth is geometric (>= 1.5x), which bounds the retired bytes by about twice the final size."""

    def __init__(self):
        self.buf = {}
        self.retired = []

    def get(self, device, nbytes):
        key = (device, WS_SLOT["i"])     # one scratch set per engine lane: lanes run concurrently on their own streams
        b = self.buf.get(key)
        if b is None or b.numel() < nbytes:
            if b is not None:
                self.retired.append(b)
                nbytes = max(int(nbytes), int(b.numel() * 1.5))
            b = torch.empty(max(int(nbytes), 1), dtype=torch.uint8, device=device)
            self.buf[key] = b
        return b


_nms_ws = _Workspace()
_nms_ws_side = _Workspace()   # second scratch buffer: an NMS running concurrently on another stream must not share the first
_pan_ws = _Workspace()


def nms_segmented(boxes_sorted, seg_offsets, max_seg_len, thresh, side=False):
    """boxes_sorted [total,4] fp32 sorted by descending score inside each segment; seg_offsets int32
    [S+1] (device).  Returns (keep [S,max_seg_len] int32 positions relative to segment start,
    counts [S] int32) -- both on the device, no host synchronisation."""
    require_cuda(boxes_sorted, seg_offsets)
    boxes_sorted = f32c(boxes_sorted)
    assert seg_offsets.dtype == torch.int32
    S = seg_offsets.numel() - 1
    dev = boxes_sorted.device
    nbytes = C.c_size_t(0)
    check(lib().upsnet_nms_workspace_bytes(S, max_seg_len, C.byref(nbytes)), "nms_workspace_bytes")
    ws = (_nms_ws_side if side else _nms_ws).get(dev, nbytes.value)
    keep = torch.empty((S, max_seg_len), dtype=torch.int32, device=dev)
    cnt = torch.empty((S,), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev), _Timed("nms", 2, {"bytes": 20.0 * boxes_sorted.shape[0]}, dev):
        check(lib().upsnet_nms_segmented(ptr(boxes_sorted), ptr(seg_offsets), S, max_seg_len, float(thresh),
                                         ptr(keep), ptr(cnt), ptr(ws), ws.numel(), stream_ptr(dev)), "nms")
    return keep, cnt


def rpn_decode(bbox_preds, top_idx, shapes, strides, base_anchors, A, im_h, im_w):
    """Fused anchors + bbox_transform + clip_boxes for the pre-NMS top-k of every level (one launch).
    bbox_preds[l] fp32 [4A,h,w] (contiguous), top_idx[l] int64 [k_l] flat (y,x,a) indices, base_anchors float64
    [L,A,4] on the device.  -> boxes [sum k_l, 4]."""
    L = len(bbox_preds)
    dev = bbox_preds[0].device
    require_cuda(*bbox_preds, *top_idx, base_anchors)
    assert base_anchors.dtype == torch.float64 and base_anchors.is_contiguous()
    bbox_preds = [f32c(b) for b in bbox_preds]
    top_idx = [t.contiguous() for t in top_idx]
    assert all(t.dtype == torch.int64 for t in top_idx)
    ks = [int(t.numel()) for t in top_idx]
    out = torch.empty((sum(ks), 4), dtype=torch.float32, device=dev)
    vp, ci = C.c_void_p, C.c_int
    with torch.cuda.device(dev), _Timed("rpn_decode", 1, {"bytes": 56.0 * sum(ks)}, dev):
        check(lib().upsnet_rpn_decode((vp * L)(*[b.data_ptr() for b in bbox_preds]), (vp * L)(*[t.data_ptr() for t in top_idx]),
                                      (ci * L)(*ks), (ci * L)(*[int(s[0]) for s in shapes]), (ci * L)(*[int(s[1]) for s in shapes]),
                                      (ci * L)(*[int(s) for s in strides]), ptr(base_anchors), L, int(A), float(im_h), float(im_w),
                                      ptr(out), stream_ptr(dev)), "rpn_decode")
    return out


_topk_ws = None


def rpn_topk(probs, A, pre_nms_top_n):
    """Top pre_nms_top_n anchors of every level in one call.  probs[l] fp32 [A,h,w] -> (scores [sum k_l], flat (y,x,a)
    indices int64 [sum k_l], [k_l]): sorted by descending score, ties by ascending index."""
    global _topk_ws
    L = len(probs)
    dev = probs[0].device
    require_cuda(*probs)
    probs = [f32c(pr) for pr in probs]
    ks = [min(int(pre_nms_top_n), int(pr.numel())) for pr in probs]
    out_s = torch.empty((sum(ks),), dtype=torch.float32, device=dev)
    out_i = torch.empty((sum(ks),), dtype=torch.int64, device=dev)
    if _topk_ws is None:
        _topk_ws = _Workspace()
    nbytes = C.c_size_t(0)
    check(lib().upsnet_rpn_topk_workspace_bytes(L, C.byref(nbytes)), "rpn_topk_workspace_bytes")
    ws = _topk_ws.get(dev, nbytes.value)
    vp, ci = C.c_void_p, C.c_int
    with torch.cuda.device(dev), _Timed("rpn_topk", 7, {"bytes": 4.0 * 6 * sum(pr.numel() for pr in probs)}, dev):
        check(lib().upsnet_rpn_topk((vp * L)(*[pr.data_ptr() for pr in probs]), (ci * L)(*[int(pr.shape[-2]) for pr in probs]),
                                    (ci * L)(*[int(pr.shape[-1]) for pr in probs]), L, int(A), int(pre_nms_top_n),
                                    ptr(out_s), ptr(out_i), ptr(ws), ws.numel(), stream_ptr(dev)), "rpn_topk")
    return out_s, out_i, ks


def rpn_collect(keep, cnt, offs, boxes, scores, post_nms_top_n):
    """Fused proposal collect after the per-level NMS -> (rois [post,5], scores [post], valid [post] bool)."""
    require_cuda(keep, cnt, offs, boxes, scores)
    dev = boxes.device
    S, M = keep.shape
    post = int(post_nms_top_n)
    rois = torch.empty((post, 5), dtype=torch.float32, device=dev)
    out_s = torch.empty((post,), dtype=torch.float32, device=dev)
    ok = torch.empty((post,), dtype=torch.bool, device=dev)
    with torch.cuda.device(dev), _Timed("rpn_collect", 1, {"bytes": 28.0 * post}, dev):
        check(lib().upsnet_rpn_collect(ptr(keep), ptr(cnt), ptr(offs), ptr(f32c(boxes)), ptr(f32c(scores)), S, M, post,
                                       ptr(rois), ptr(out_s), ptr(ok), stream_ptr(dev)), "rpn_collect")
    return rois, out_s, ok


def maskroi_prepare(rois, roi_valid, bbox_delta, cls_prob, class_agnostic, score_thresh, weights, im_h, im_w):
    """Fused MaskROI front half -> (sc [n], cls int32 [n], bx [n,4], offs int32 [nseg+1]), n = R*(C-1):
    candidates first in (segment, score desc, index) order, decoded and clipped."""
    require_cuda(rois, roi_valid, bbox_delta, cls_prob)
    rois, bbox_delta, cls_prob = f32c(rois), f32c(bbox_delta), f32c(cls_prob)
    assert roi_valid.dtype == torch.bool and roi_valid.is_contiguous()
    R, Cn = cls_prob.shape
    n = R * (Cn - 1)
    dev = rois.device
    nseg = 1 if class_agnostic else Cn - 1
    sc = torch.empty((n,), dtype=torch.float32, device=dev)
    cls = torch.empty((n,), dtype=torch.int32, device=dev)
    bx = torch.empty((n, 4), dtype=torch.float32, device=dev)
    offs = torch.empty((nseg + 1,), dtype=torch.int32, device=dev)
    w4 = (C.c_float * 4)(*[float(w) for w in weights])
    with torch.cuda.device(dev), _Timed("maskroi", 1, {"bytes": 4.0 * (rois.numel() + bbox_delta.numel() + cls_prob.numel())}, dev):
        check(lib().upsnet_maskroi_prepare(ptr(rois), ptr(roi_valid), ptr(bbox_delta), ptr(cls_prob), R, Cn,
                                           1 if class_agnostic else 0, float(score_thresh), w4, float(im_h), float(im_w),
                                           ptr(sc), ptr(cls), ptr(bx), ptr(offs), stream_ptr(dev)), "maskroi_prepare")
    return sc, cls, bx, offs


def maskroi_finish(keep, cnt, offs, sc, cls, bx, top_n, cap):
    """Fused MaskROI back half -> (scores [cap], boxes [cap,5], cls int64 [cap], n int32 device scalar, flags int32 device
    scalar).  flags != 0 means the static buffers truncated what the reference would have kept: bit 0 = more NMS survivors
    than candidate slots, bit 1 = a tie at the top-n threshold larger than the output slack."""
    require_cuda(keep, cnt, offs, sc, cls, bx)
    dev = sc.device
    nseg, M = keep.shape
    out_sc = torch.empty((cap,), dtype=torch.float32, device=dev)
    out_bx = torch.empty((cap, 5), dtype=torch.float32, device=dev)
    out_cls = torch.empty((cap,), dtype=torch.int64, device=dev)
    n_out = torch.empty((2,), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev), _Timed("maskroi", 1, {"bytes": 32.0 * cap}, dev):
        check(lib().upsnet_maskroi_finish(ptr(keep), ptr(cnt), ptr(offs), ptr(sc), ptr(cls), ptr(bx), nseg, M, int(top_n),
                                          int(cap), ptr(out_sc), ptr(out_bx), ptr(out_cls), ptr(n_out), stream_ptr(dev)),
              "maskroi_finish")
    return out_sc, out_bx, out_cls, n_out[0], n_out[1]


def nms(boxes, scores, thresh):
    """Device API: boxes [N,4], scores [N] (CUDA) -> int64 indices of kept boxes, descending score
    (== order[keep] of nms/gpu_nms.pyx:32-38).  One D2H of the count only."""
    require_cuda(boxes, scores)
    n = boxes.shape[0]
    if n == 0:
        return torch.empty((0,), dtype=torch.int64, device=boxes.device)
    _, order = torch.sort(scores.float(), descending=True, stable=True)
    seg = torch.tensor([0, n], dtype=torch.int32, device=boxes.device)
    keep, cnt = nms_segmented(boxes.float()[order], seg, n, thresh)
    k = int(cnt.item())
    return order[keep[0, :k].long()]


def gpu_nms(dets, thresh, device_id=0):
    """nms/gpu_nms.pyx:23-38: dets np.float32 [N,5] on the HOST -> list[int]."""
    dets = np.ascontiguousarray(dets, dtype=np.float32)
    if dets.shape[0] == 0:
        return []
    d = torch.from_numpy(dets).to(torch.device("cuda", device_id))
    return nms(d[:, :4], d[:, 4], thresh).cpu().tolist()


def gpu_nms_wrapper(thresh, device_id):
    """nms/nms.py:43-46."""
    def _nms(dets):
        return gpu_nms(dets, thresh, device_id)
    return _nms


# ------------------------------------------------------------------------------------------------
# modules (reference names / signatures / parameter names)
# ------------------------------------------------------------------------------------------------
class DeformConv(nn.Module):
    """operators/modules/deform_conv.py:27-64.  Parameters are created on CUDA like the reference."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1,
                 deformable_groups=1, bias=True):
        super().__init__()
        assert in_channels % groups == 0, 'in_channels must be divisible by groups'
        assert out_channels % groups == 0, 'out_channels must be divisible by groups'
        assert out_channels % deformable_groups == 0, 'out_channels must be divisible by deformable groups'
        self.in_channels, self.out_channels = in_channels, out_channels
        self.kernel_size, self.stride = _pair(kernel_size), _pair(stride)
        self.padding, self.dilation = _pair(padding), _pair(dilation)
        self.groups, self.deformable_groups = groups, deformable_groups
        dev = torch.device("cuda") if torch.cuda.is_available() else torch.device("cpu")
        self.weight = Parameter(torch.empty(out_channels, in_channels // groups, *self.kernel_size, device=dev))
        if bias:
            self.bias = Parameter(torch.empty(out_channels, device=dev))
        else:
            self.register_parameter('bias', None)
        self.reset_parameters()

    def reset_parameters(self):
        n = self.in_channels
        for k in self.kernel_size:
            n *= k
        stdv = 1. / math.sqrt(n)
        self.weight.data.uniform_(-stdv, stdv)
        if self.bias is not None:
            self.bias.data.uniform_(-stdv, stdv)

    def forward(self, data, offset):
        if torch.is_grad_enabled() and self.deformable_groups == 1 and \
                any(t is not None and t.requires_grad for t in (data, offset, self.weight, self.bias)):
            from .training import DeformConvFunction      # training path: hand-written backward kernels (csrc/backward.cu)
            return DeformConvFunction.apply(data, offset, self.weight, self.bias, self.stride, self.padding, self.dilation)
        return deform_conv(data, offset, self.weight, self.bias, self.stride, self.padding, self.dilation,
                           self.deformable_groups)


class DeformConvWithOffset(nn.Module):
    """operators/modules/deform_conv.py:67-78 (submodules `conv_offset`, `conv`)."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1,
                 deformable_groups=1, bias=True):
        super().__init__()
        self.conv_offset = nn.Conv2d(in_channels, kernel_size * kernel_size * 2 * deformable_groups,
                                     kernel_size=3, stride=1, padding=1)
        self.conv_offset.weight.data.zero_()
        self.conv_offset.bias.data.zero_()
        self.conv = DeformConv(in_channels, out_channels, kernel_size=kernel_size, stride=stride,
                               padding=padding, dilation=dilation, groups=groups,
                               deformable_groups=deformable_groups, bias=bias)

    def forward(self, x):
        offset = conv2d(x, self.conv_offset.weight, self.conv_offset.bias, 1, 1, 1, out_format="nchw")
        return self.conv(x, offset)


class ModDeformConv(DeformConv):
    """operators/modules/mod_deform_conv.py:24-67: forward(data, offset_mask) with
    offset = cat(chunk0, chunk1), mask = 2*sigmoid(chunk2)."""

    def forward(self, data, offset_mask):
        offset_1, offset_2, mask = torch.chunk(offset_mask, 3, dim=1)
        offset = torch.cat((offset_1, offset_2), dim=1)
        mask = torch.sigmoid(mask) * 2
        if torch.is_grad_enabled() and self.deformable_groups == 1 and \
                any(t is not None and t.requires_grad for t in (data, offset_mask, self.weight, self.bias)):
            from .training import ModDeformConvFunction
            return ModDeformConvFunction.apply(data, offset, mask, self.weight, self.bias, self.stride, self.padding, self.dilation)
        return deform_conv(data, offset, self.weight, self.bias, self.stride, self.padding, self.dilation,
                           self.deformable_groups, mask=mask)


ModulatedDeformConv = ModDeformConv


class ModDeformConvWithOffsetMask(nn.Module):
    """operators/modules/mod_deform_conv.py:70-81 (submodules `conv_offset_mask`, `conv`)."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1,
                 deformable_groups=1, bias=True):
        super().__init__()
        self.conv_offset_mask = nn.Conv2d(in_channels, kernel_size * kernel_size * 3 * deformable_groups,
                                          kernel_size=3, stride=1, padding=1)
        self.conv_offset_mask.weight.data.zero_()
        self.conv_offset_mask.bias.data.zero_()
        self.conv = ModDeformConv(in_channels, out_channels, kernel_size=kernel_size, stride=stride,
                                  padding=padding, dilation=dilation, groups=groups,
                                  deformable_groups=deformable_groups, bias=bias)

    def forward(self, x):
        om = conv2d(x, self.conv_offset_mask.weight, self.conv_offset_mask.bias, 1, 1, 1, out_format="nchw")
        return self.conv(x, om)


class RoIAlignFunction:
    """functions/roialign.py:21-43 call shape: RoIAlignFunction(ph, pw, scale)(features, rois)."""

    def __init__(self, pooled_height, pooled_width, spatial_scale, sampling_ratio=2):
        self.pooled_width, self.pooled_height = int(pooled_width), int(pooled_height)
        self.spatial_scale, self.sampling_ratio = float(spatial_scale), sampling_ratio

    def __call__(self, features, rois):
        if not features.is_cuda:
            raise Exception('not implemented')
        if torch.is_grad_enabled() and features.requires_grad:
            from .training import RoIAlignFunction as _F      # training path: upsnet_roi_align_backward
            return _F.apply(features, rois, self.pooled_height, self.pooled_width, self.spatial_scale, self.sampling_ratio)
        return roi_align(features, rois, self.pooled_height, self.pooled_width, self.spatial_scale,
                         self.sampling_ratio)


class RoIAlign(nn.Module):
    """operators/modules/roialign.py:20-29."""

    def __init__(self, pooled_height, pooled_width, spatial_scale):
        super().__init__()
        self.pooled_width, self.pooled_height = int(pooled_width), int(pooled_height)
        self.spatial_scale = float(spatial_scale)

    def forward(self, features, rois):
        return RoIAlignFunction(self.pooled_height, self.pooled_width, self.spatial_scale)(features, rois)


ROIAlign = RoIAlign


class FPNRoIAlign(nn.Module):
    """operators/modules/fpn_roi_align.py:22-62; forward([P2..P5], rois[N,5]) -> [N,C,ph,pw]."""

    def __init__(self, pooled_height, pooled_width, spatial_scale, with_expand=False):
        super().__init__()
        self.pooled_width, self.pooled_height = int(pooled_width), int(pooled_height)
        self.spatial_scale = spatial_scale
        self.with_expand = with_expand

    def forward(self, feat, rois):
        return fpn_roi_align(list(feat), rois, self.pooled_height, self.pooled_width, self.spatial_scale,
                             layout="auto")


# ------------------------------------------------------------------------------------------------
# panoptic head
# ------------------------------------------------------------------------------------------------
def panoptic_fuse(fcn_output, mask_rois, cls_prob, mask_logit, cls_idx, num_stuff, fraction_threshold=0.3,
                  want_sem=False, n_dev=None, workspace_bytes=None, up4=False):
    """Fused MaskRemoval + SegTerm + void/argmax (upsnet_panoptic_head).
    fcn_output [1,S,H,W]; mask_rois [n,4]; cls_prob [n]; mask_logit [n,1,28,28] or [n,28,28];
    cls_idx int64 [n].  Returns (keep_inds int64 [k], panoptic_output int64 [1,H,W][, sem int64 [1,H,W]]).
    With n_dev (int32 device scalar, actual count <= n) nothing is read back: keep_inds is the padded
    [n] buffer and the device count k is returned as an extra last element (static-shape engine path).
    up4=True: `fcn_output` is the quarter-resolution score map [1,S,H/4,W/4] (FCNHead's 'fcn_score'); its x4 bilinear
    up-sampling (models/fcn.py:88-101) is evaluated inside the kernel, bit-identical to upsample_bilinear(score, 4)."""
    require_cuda(fcn_output, mask_rois, cls_prob, mask_logit, cls_idx)
    assert fcn_output.dim() == 4 and fcn_output.shape[0] == 1, "only support batch size = 1"
    fcn = f32c(fcn_output)
    _, S, H, W = fcn.shape
    if up4:
        Hs, Ws, H, W = H, W, 4 * H, 4 * W
    boxes, prob, ml = f32c(mask_rois), f32c(cls_prob).reshape(-1), f32c(mask_logit)
    cls = cls_idx.to(torch.int64).contiguous()
    n = boxes.shape[0]
    assert boxes.shape == (n, 4) and prob.numel() == n and ml.numel() == n * 784 and cls.numel() == n
    dev = fcn.device
    num_thing = S - num_stuff
    nbytes = C.c_size_t(0)
    check(lib().upsnet_panoptic_workspace_bytes(n, H, W, num_thing, C.byref(nbytes)), "panoptic_workspace_bytes")
    if workspace_bytes is None:
        ws = _pan_ws.get(dev, nbytes.value)
    else:
        # caller-chosen (smaller) workspace: the instances' bit windows are then processed in several rounds of
        # consecutive score ranks -- same results (upsnet_panoptic_workspace_min_bytes is the floor)
        ws = torch.empty(int(workspace_bytes), dtype=torch.uint8, device=dev)
    if n_dev is not None:
        assert n_dev.dtype == torch.int32 and n_dev.is_cuda
    keep = torch.zeros((max(n, 1),), dtype=torch.int64, device=dev)
    k = torch.empty((1,), dtype=torch.int32, device=dev)
    labels = torch.empty((1, H, W), dtype=torch.int64, device=dev)
    sem = torch.empty((1, H, W), dtype=torch.int64, device=dev) if want_sem else None
    work = {"bytes": 4.0 * S * H * W / (16 if up4 else 1) + 8.0 * H * W * (2 if want_sem else 1) + n * (4.0 * 784 + 24)}
    with torch.cuda.device(dev), _Timed("panoptic_head", 5, work, dev):
        if up4:
            check(lib().upsnet_panoptic_head_up4(ptr(fcn), S, Hs, Ws, ptr(boxes), ptr(prob), ptr(ml), ptr(cls), n, ptr(n_dev),
                                                 num_stuff, float(fraction_threshold), ptr(keep), ptr(k), ptr(labels),
                                                 ptr(sem), ptr(ws), ws.numel(), stream_ptr(dev)), "panoptic_head_up4")
        else:
            check(lib().upsnet_panoptic_head(ptr(fcn), S, H, W, ptr(boxes), ptr(prob), ptr(ml), ptr(cls), n, ptr(n_dev),
                                             num_stuff, float(fraction_threshold), ptr(keep), ptr(k), ptr(labels),
                                             ptr(sem), ptr(ws), ws.numel(), stream_ptr(dev)), "panoptic_head")
    if n_dev is not None:
        return (keep, labels, sem, k) if want_sem else (keep, labels, k)
    keep = keep[:int(k.item())]
    return (keep, labels, sem) if want_sem else (keep, labels)


class MaskRemoval(nn.Module):
    """operators/modules/mask_removal.py:23-93, same signature and return values:
    forward(mask_rois[n,4], cls_prob[n], mask_prob[n,1,28,28], cls_idx[n], im_shape) ->
    (keep_inds LongTensor [k], mask_energy [1,k,H,W]).  Runs the device kernels of the fused head and, for
    API parity, materialises mask_energy (the fused PanopticHead never does)."""

    def __init__(self, fraction_threshold=0.3):
        super().__init__()
        self.fraction_threshold = fraction_threshold

    def forward(self, mask_rois, cls_prob, mask_prob, cls_idx, im_shape):
        require_cuda(mask_rois, cls_prob, mask_prob, cls_idx)
        boxes, prob, ml = f32c(mask_rois), f32c(cls_prob).reshape(-1), f32c(mask_prob)
        cls = cls_idx.to(torch.int64).contiguous()
        n, (H, W) = boxes.shape[0], (int(im_shape[0]), int(im_shape[1]))
        dev = boxes.device
        num_thing = max(int(cls.max().item()), 1)          # mask_image planes: np.max(cls_idx) (host read, as the reference)
        nbytes = C.c_size_t(0)
        check(lib().upsnet_panoptic_workspace_bytes(n, H, W, num_thing, C.byref(nbytes)), "panoptic_workspace_bytes")
        ws = _pan_ws.get(dev, nbytes.value)
        keep = torch.zeros((max(n, 1),), dtype=torch.int64, device=dev)
        k = torch.empty((1,), dtype=torch.int32, device=dev)
        energy = torch.empty((n, H, W), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev), _Timed("mask_removal", 5, {"bytes": 4.0 * n * H * W}, dev):
            check(lib().upsnet_mask_removal(ptr(boxes), ptr(prob), ptr(ml), ptr(cls), n, None, H, W, num_thing,
                                            float(self.fraction_threshold), ptr(keep), ptr(k), ptr(energy), ptr(ws),
                                            ws.numel(), stream_ptr(dev)), "mask_removal")
        kk = int(k.item())
        return keep[:kk], energy[:kk].unsqueeze(0)


class SegTerm(nn.Module):
    """operators/modules/unary_logits.py:69-105: (stuff logits view, per-instance boxed copy of the instance's
    thing-class logit).  Pure tensor slicing on the device, same integer conventions as the reference
    (int() truncation, numpy half-to-even round, python-slice clamping).  API parity only: the fused
    PanopticHead evaluates the same windows inside pan_fuse without building [1,k,H,W]."""

    def __init__(self, num_seg_classes, box_scale=1 / 4.0, class_mapping=None, thresh=0.3, num_classes=None):
        super().__init__()
        num_classes = num_classes if num_classes is not None else 9
        self.class_mapping = dict(zip(range(1, num_classes), range(num_seg_classes - num_classes + 1, num_seg_classes))) \
            if class_mapping is None else class_mapping
        self.num_seg_classes, self.num_inst_classes, self.box_scale = num_seg_classes, len(self.class_mapping), box_scale

    def forward(self, cls_indices, seg_score, boxes):
        assert seg_score.shape[0] == 1, "only support batch size = 1"
        cls_np = cls_indices.detach().cpu().numpy()
        seg_energy = seg_score[[0], :-self.num_inst_classes, :, :]
        b = boxes.detach().cpu().numpy()[:, 1:] * self.box_scale
        if cls_np.size == 0:
            return seg_energy, torch.ones_like(seg_energy[[0], [0], :, :]).view(1, 1, seg_energy.shape[2], seg_energy.shape[3]) * -10
        inst = torch.zeros((1, cls_np.shape[0], seg_score.shape[2], seg_score.shape[3]), device=seg_score.device)
        for i in range(cls_np.shape[0]):
            if cls_np[i] == 0:
                continue
            y0, y1 = int(b[i][1]), int(b[i][3].round() + 1)
            x0, x1 = int(b[i][0]), int(b[i][2].round() + 1)
            inst[0, i, y0:y1, x0:x1] = seg_score[0, self.class_mapping[int(cls_np[i])], y0:y1, x0:x1]
        return seg_energy, inst


class MaskTerm(nn.Module):
    """operators/modules/unary_logits.py:24-66 (training twin of MaskRemoval's paste: every roi's 28x28 mask logit is
    bilinearly resized (align_corners=False, ATen rule) to its box at 1/4 scale and pasted into [1,n,h,w]).  Tensor ops
    on the device, differentiable through torch autograd like the reference's F.upsample."""

    def __init__(self, num_seg_classes, box_scale=1 / 4.0, class_mapping=None, num_classes=9, mask_size=28):
        super().__init__()
        self.num_seg_classes, self.box_scale, self.mask_size = num_seg_classes, box_scale, mask_size
        self.class_mapping = dict(zip(range(1, num_classes), range(num_seg_classes - num_classes + 1, num_seg_classes))) \
            if class_mapping is None else class_mapping

    def forward(self, masks, boxes, cls_indices, seg_score):
        assert seg_score.shape[0] == 1, "only support batch size = 1"
        boxes = boxes[:, 1:] * self.box_scale
        H, W = int(seg_score.shape[2]), int(seg_score.shape[3])
        energy = torch.zeros((1, masks.shape[0], H, W), device=seg_score.device)
        ref = boxes.long().tolist()                                          # one host read for all boxes
        for i, (bx0, by0, bx1, by1) in enumerate(ref):
            w, h = max(bx1 - bx0 + 1, 1), max(by1 - by0 + 1, 1)
            m = torch.nn.functional.interpolate(masks[i, 0].view(1, 1, self.mask_size, self.mask_size), size=(h, w),
                                                mode="bilinear", align_corners=False)
            x0, x1, y0, y1 = max(bx0, 0), min(bx1 + 1, W), max(by0, 0), min(by1 + 1, H)
            if x1 > x0 and y1 > y0:
                energy[0, i, y0:y1, x0:x1] = m[0, 0, (y0 - by0):(y1 - by0), (x0 - bx0):(x1 - bx0)]
        return energy


class MaskMatching(nn.Module):
    """operators/modules/mask_matching.py:27-62: panoptic ground truth = stuff labels kept, every (kept) gt instance mask
    painted with its channel index, the rest void (255) or the extra 'unmatched' channel."""

    def __init__(self, num_seg_classes, enable_void, class_mapping=None, num_classes=9):
        super().__init__()
        self.class_mapping = dict(zip(range(1, num_classes), range(num_seg_classes - num_classes + 1, num_seg_classes))) \
            if class_mapping is None else class_mapping
        self.num_seg_classes, self.num_classes = num_seg_classes, num_classes
        self.num_inst_classes, self.enable_void = len(self.class_mapping), enable_void

    def forward(self, gt_segs, gt_masks, keep_inds=None):
        matched = torch.ones_like(gt_segs) * -1
        matched = torch.where(gt_segs <= self.num_seg_classes - self.num_classes, gt_segs, matched)
        matched = torch.where(gt_segs >= 255, gt_segs, matched)
        if keep_inds is not None:
            gt_masks = gt_masks[keep_inds]
        base = self.num_seg_classes - self.num_inst_classes
        for i in range(gt_masks.shape[0]):
            matched[(gt_masks[[i], :, :] != 0) & (gt_masks[[i], :, :] != 255)] = i + base
        matched[matched == -1] = (base + gt_masks.shape[0]) if keep_inds is not None else 255
        return matched


class PanopticHead(nn.Module):
    """The parameter-free panoptic head of models/resnet_upsnet.py:217-247 as one module (the
    reference has no such class: SURVEY.md F1).  forward takes what lines 220-227 consume."""

    def __init__(self, num_seg_classes, num_classes, fraction_threshold=0.3):
        super().__init__()
        self.num_seg_classes, self.num_classes = num_seg_classes, num_classes
        self.num_stuff = num_seg_classes - num_classes + 1  # unary_logits.py:72
        self.fraction_threshold = fraction_threshold

    def forward(self, fcn_output, mask_rois, cls_prob, mask_score, cls_idx, want_sem=False):
        """mask_rois [n,5] (batch,x1,y1,x2,y2) or [n,4]; mask_score [n,1,28,28] = logit of the predicted
        class (resnet_upsnet.py:220).  Returns dict(keep_inds, panoptic_outputs[, fcn_outputs])."""
        boxes = mask_rois[:, 1:] if mask_rois.shape[1] == 5 else mask_rois
        out = panoptic_fuse(fcn_output, boxes, cls_prob, mask_score, cls_idx, self.num_stuff,
                            self.fraction_threshold, want_sem)
        res = {'keep_inds': out[0], 'panoptic_outputs': out[1]}
        if want_sem:
            res['fcn_outputs'] = out[2]
        return res


# ------------------------------------------------------------------------------------------------
# callers either side of the forward: unified panoptic result, input pipeline
# ------------------------------------------------------------------------------------------------
_uni_ws = _Workspace()


def unified_pan_result(seg, pan, cls_inds, num_seg_classes, num_classes, stuff_area_limit=4 * 64 * 64, k_dev=None,
                       check_errors=True):
    """dataset/base_dataset.py:332-371 get_unified_pan_result for one image, on the device.
    seg / pan: int64 [H,W] or [1,H,W] ('fcn_outputs' / 'panoptic_outputs' of the forward); cls_inds int64 [k]
    ('panoptic_cls_inds').  -> uint8 [H,W,3] (semantic class, instance number, 0).  With check_errors (one int D2H) an
    instance label without a cls_inds entry raises IndexError like the reference."""
    require_cuda(seg, pan, cls_inds)
    seg = seg.reshape(seg.shape[-2:]).to(torch.int64).contiguous()
    pan = pan.reshape(pan.shape[-2:]).to(torch.int64).contiguous()
    cls = cls_inds.to(torch.int64).contiguous()
    H, W = pan.shape
    dev = pan.device
    nb = C.c_size_t(0)
    check(lib().upsnet_unified_pan_workspace_bytes(int(num_seg_classes), C.byref(nb)), "unified_pan_workspace_bytes")
    ws = _uni_ws.get(dev, nb.value)
    out = torch.empty((H, W, 3), dtype=torch.uint8, device=dev)
    err = torch.zeros((1,), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev), _Timed("unified_pan", 5, {"bytes": 19.0 * H * W}, dev):
        check(lib().upsnet_unified_pan_result(ptr(seg), ptr(pan), ptr(cls), int(cls.numel()), ptr(k_dev), H, W,
                                              int(num_seg_classes), int(num_classes), int(stuff_area_limit), ptr(out), ptr(err),
                                              ptr(ws), ws.numel(), stream_ptr(dev)), "unified_pan_result")
    if check_errors:
        e = int(err.item())
        if e & 2:
            raise IndexError("panoptic label without an entry in cls_inds (base_dataset.py:350)")
        if e & 1:
            raise ValueError("label out of range in seg / pan")
    return out


def prep_image(image_hwc_u8, pixel_means, scale=1.0, stride=32):
    """dataset/base_dataset.py:143-174 prep_im_for_blob + :898-923 im_list_to_blob on the device: uint8 [h,w,3] (BGR)
    -> fp32 blob [1,3,Hp,Wp] (mean-subtracted, bilinearly resized by `scale`, zero-padded to a multiple of `stride`)
    and the resized (h, w)."""
    require_cuda(image_hwc_u8)
    assert image_hwc_u8.dtype == torch.uint8 and image_hwc_u8.dim() == 3 and image_hwc_u8.shape[2] == 3
    im = image_hwc_u8.contiguous()
    h, w = int(im.shape[0]), int(im.shape[1])
    ho, wo = int(np.rint(h * scale)), int(np.rint(w * scale))              # cvRound(src * f) (cv2.resize dsize rule)
    Hp, Wp = int(math.ceil(ho / float(stride)) * stride), int(math.ceil(wo / float(stride)) * stride)
    blob = torch.empty((1, 3, Hp, Wp), dtype=torch.float32, device=im.device)
    pm = (C.c_double * 3)(*[float(v) for v in pixel_means])
    with torch.cuda.device(im.device), _Timed("prep_image", 1, {"bytes": 3.0 * h * w + 12.0 * Hp * Wp}, im.device):
        check(lib().upsnet_prep_image(ptr(im), h, w, float(scale), ho, wo, Hp, Wp, pm, ptr(blob), stream_ptr(im.device)), "prep_image")
    return blob, (ho, wo)


# ------------------------------------------------------------------------------------------------
# Row f4: im_post (upsnet_end2end_test.py:95-152) -- mask paste + COCO RLE
# ------------------------------------------------------------------------------------------------
_impost_ws = _Workspace()


def rle_to_string(cnts):
    """pycocotools maskApi.c rleToString: the compressed `counts` bytes of a COCO RLE from its run lengths (host side; a
    pure function of the numbers the device kernel produces)."""
    out = bytearray()
    cnts = [int(c) for c in cnts]
    for i, x in enumerate(cnts):
        if i > 2:
            x -= cnts[i - 2]
        more = True
        while more:
            c = x & 0x1f
            x >>= 5
            more = (x != -1) if (c & 0x10) else (x != 0)
            if more:
                c |= 0x20
            out.append(c + 48)
    return bytes(out)


def im_post_rle(pred_boxes, pred_masks, cls_inds, im_h, im_w, n_dev=None, cap=None):
    """upsnet_im_post_rle: the COCO run lengths of every detection's pasted mask, computed on the device without
    materialising the [H,W] images.  pred_boxes [n,4] or [n,5] (batch index first, like the model's `pred_boxes`),
    pred_masks [n,C,M,M] probabilities, cls_inds [n].  Returns (counts uint32 [n,cap] as int64-viewable tensor, run_len
    int32 [n]) on the device; raises if a detection needs more than `cap` counts (default 16 per image column)."""
    require_cuda(pred_boxes, pred_masks, cls_inds)
    boxes = f32c(pred_boxes[:, 1:] if pred_boxes.shape[1] == 5 else pred_boxes)
    masks = f32c(pred_masks)
    n, Cc, M, M2 = masks.shape
    assert M == M2 and boxes.shape == (n, 4) and cls_inds.numel() == n
    cls = cls_inds.to(torch.int64).contiguous()
    dev = masks.device
    cap = int(cap) if cap else 16 * int(im_w) + 64
    nb = C.c_size_t(0)
    check(lib().upsnet_im_post_workspace_bytes(n, cap, C.byref(nb)), "im_post_workspace_bytes")
    ws = _impost_ws.get(dev, nb.value)
    counts = torch.empty((max(n, 1), cap), dtype=torch.int32, device=dev)      # uint32 payload
    run_len = torch.zeros((max(n, 1),), dtype=torch.int32, device=dev)
    ovf = torch.zeros((1,), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev), _Timed("im_post", 1, {"bytes": 4.0 * masks.numel()}, dev):
        check(lib().upsnet_im_post_rle(ptr(masks), Cc, M, ptr(boxes), ptr(cls), n, ptr(n_dev), int(im_h), int(im_w), ptr(counts),
                                       cap, ptr(run_len), ptr(ovf), ptr(ws), ws.numel(), stream_ptr(dev)), "im_post_rle")
    return counts[:n], run_len[:n], ovf


def im_post(boxes_all, masks_all, scores, pred_boxes, pred_masks, cls_inds, num_classes, im_info):
    """Drop-in for upsnet_end2end_test.py:95-152 `im_post` (same arguments, same side effects on boxes_all / masks_all):
    device tensors in, per-class lists of [x1,y1,x2,y2,score] arrays and COCO RLE dicts out.  One kernel launch and one
    D2H copy of the run lengths per image instead of n x (cv2.resize + paste + pycocotools encode) on the host."""
    H, W = int(im_info[0]), int(im_info[1])
    counts, run_len, ovf = im_post_rle(pred_boxes, pred_masks, cls_inds, H, W)
    if int(ovf.item()):
        raise UpsnetError("im_post: a mask needs more RLE counts than the buffer holds (pass a larger cap)")
    rl = run_len.cpu().numpy()
    cn = counts.cpu().numpy().view(np.uint32)
    boxes_np = (pred_boxes[:, 1:] if pred_boxes.shape[1] == 5 else pred_boxes).float().cpu().numpy()
    sc = scores.float().cpu().numpy().reshape(-1, 1)
    ci = cls_inds.cpu().numpy()
    for idx in range(1, num_classes):
        sel = np.flatnonzero(ci == idx)
        cls_boxes = np.hstack([boxes_np[sel], sc[sel]]) if sel.size else np.zeros((0, 5), np.float32)
        segms = [{"size": [H, W], "counts": rle_to_string(cn[d, :rl[d]]).decode()} for d in sel]
        boxes_all[idx].append(cls_boxes)
        masks_all[idx].append(segms)
