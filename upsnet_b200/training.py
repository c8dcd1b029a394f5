"""Training-side pieces of the hot path (BASELINE config #4: UPSNet-50 end2end_train, bf16, 8 GPUs, NCCL all-reduce):

* autograd Functions for the custom operators with hand-written sm_100a BACKWARD kernels (csrc/backward.cu):
    DeformConvFunction / ModDeformConvFunction   operators/functions/deform_conv.py:26-108, mod_deform_conv.py:25-118
    RoIAlignFunction                             operators/functions/roialign.py:21-58
  The dense GEMMs of the deformable backward (d(weight) = dY col^T, d(col) = W^T dY) are library calls (torch.mm), like the
  reference's; the gather / scatter / coordinate-gradient kernels are ours.
* FlatBucketAllReduce: the gradient all-reduce of `upsnet_end2end_train.py:121` (hvd.DistributedOptimizer) as flat bf16
  buckets over torch.distributed (NCCL over NVLink on the B200 box, gloo in the CPU tests): gradients are packed per bucket,
  reduced asynchronously while the rest of backward runs, averaged and unpacked before the optimiser step.

Scope note: this is the operator / communication layer of the training configuration.  Losses, target assignment and the
optimiser are plain torch in the reference and stay that way; the dense backward convolutions are library calls.
"""
import ctypes as C

import torch
import torch.distributed as dist
from torch.nn.modules.utils import _pair

from . import _lib
from ._lib import check, f32c, lib, ptr, require_cuda, stream_ptr


def _conv_out(n, pad, dil, k, stride):
    return (n + 2 * pad - (dil * (k - 1) + 1)) // stride + 1


class _DeformConvBase(torch.autograd.Function):
    @staticmethod
    def _geom(x, weight, stride, padding, dilation):
        sh, sw = _pair(stride); ph, pw = _pair(padding); dh, dw = _pair(dilation)
        N, Cin, H, W = x.shape
        Cout, _, kh, kw = weight.shape
        return (N, Cin, H, W, Cout, kh, kw, sh, sw, ph, pw, dh, dw, _conv_out(H, ph, dh, kh, sh), _conv_out(W, pw, dw, kw, sw))

    @staticmethod
    def _forward(ctx, x, offset, mask, weight, bias, stride, padding, dilation):
        from . import operators as ops
        require_cuda(x, offset, weight, bias, mask)
        ctx.save_for_backward(x, offset, mask if mask is not None else x.new_empty(0), weight)
        ctx.has_mask, ctx.has_bias = mask is not None, bias is not None
        ctx.conv = (stride, padding, dilation)
        # fp32 CUDA-core tiles: the training forward keeps fp32 semantics of the reference (.data<float>())
        return ops.deform_conv(x, offset, weight, bias, stride, padding, dilation, 1, mask=mask, precision=_lib.PREC_FP32_SIMT)

    @staticmethod
    def _backward(ctx, grad_out):
        x, offset, mask, weight = ctx.saved_tensors
        mask = mask if ctx.has_mask else None
        stride, padding, dilation = ctx.conv
        N, Cin, H, W, Cout, kh, kw, sh, sw, ph, pw, dh, dw, Ho, Wo = _DeformConvBase._geom(x, weight, stride, padding, dilation)
        x, offset, grad_out = f32c(x), f32c(offset), f32c(grad_out)
        mask = None if mask is None else f32c(mask)
        dev = x.device
        K, P = Cin * kh * kw, Ho * Wo
        w2 = weight.reshape(Cout, K).float()
        dx = torch.empty_like(x)
        doff = torch.empty_like(offset)
        dmask = torch.empty_like(mask) if mask is not None else None
        dw_ = torch.zeros((Cout, K), dtype=torch.float32, device=dev)
        col = torch.empty((K, P), dtype=torch.float32, device=dev)
        g = (Cin, H, W, kh, kw, sh, sw, ph, pw, dh, dw)
        st = stream_ptr(dev)
        with torch.cuda.device(dev):
            for n in range(N):          # the reference loops over the batch as well (functions/deform_conv.py:84-104)
                m_n = None if mask is None else mask[n]
                go = grad_out[n].reshape(Cout, P)
                check(lib().upsnet_dcn_im2col(ptr(x[n]), ptr(offset[n]), ptr(m_n), *g, ptr(col), st), "dcn_im2col")
                dw_.addmm_(go, col.t())                                         # d(weight) += dY col^T
                dcol = torch.mm(w2.t(), go)                                     # d(col) = W^T dY
                check(lib().upsnet_dcn_col2im(ptr(dcol), ptr(offset[n]), ptr(m_n), *g, ptr(dx[n]), st), "dcn_col2im")
                check(lib().upsnet_dcn_col2im_coord(ptr(dcol), ptr(x[n]), ptr(offset[n]), ptr(m_n), *g, ptr(doff[n]),
                                                    ptr(None if dmask is None else dmask[n]), st), "dcn_col2im_coord")
        dbias = grad_out.sum(dim=(0, 2, 3)) if ctx.has_bias else None
        return dx, doff, dmask, dw_.view_as(weight), dbias


class DeformConvFunction(_DeformConvBase):
    """y = DeformConv(x, offset; weight, bias) with hand-written backward kernels (K1-K3)."""

    @staticmethod
    def forward(ctx, x, offset, weight, bias=None, stride=1, padding=0, dilation=1):
        return _DeformConvBase._forward(ctx, x, offset, None, weight, bias, stride, padding, dilation)

    @staticmethod
    def backward(ctx, grad_out):
        dx, doff, _, dw_, db = _DeformConvBase._backward(ctx, grad_out)
        return dx, doff, dw_, db, None, None, None


class ModDeformConvFunction(_DeformConvBase):
    """v2: mask is the already-activated modulation (2*sigmoid in ModDeformConv.forward); backward kernels K4-K6."""

    @staticmethod
    def forward(ctx, x, offset, mask, weight, bias=None, stride=1, padding=0, dilation=1):
        return _DeformConvBase._forward(ctx, x, offset, mask, weight, bias, stride, padding, dilation)

    @staticmethod
    def backward(ctx, grad_out):
        dx, doff, dmask, dw_, db = _DeformConvBase._backward(ctx, grad_out)
        return dx, doff, dmask, dw_, db, None, None, None


class RoIAlignFunction(torch.autograd.Function):
    """functions/roialign.py:21-58: forward = upsnet_roi_align_forward (NCHW fp32), backward = upsnet_roi_align_backward."""

    @staticmethod
    def forward(ctx, features, rois, pooled_height, pooled_width, spatial_scale, sampling_ratio=2):
        from . import operators as ops
        ctx.save_for_backward(rois)
        ctx.cfg = (tuple(features.shape), int(pooled_height), int(pooled_width), float(spatial_scale), int(sampling_ratio))
        return ops.roi_align(features, rois, pooled_height, pooled_width, spatial_scale, sampling_ratio)

    @staticmethod
    def backward(ctx, grad_out):
        (rois,) = ctx.saved_tensors
        (B, Cc, H, W), ph, pw, scale, sr = ctx.cfg
        grad_out, rois = f32c(grad_out), f32c(rois)
        dfeat = torch.empty((B, Cc, H, W), dtype=torch.float32, device=grad_out.device)
        with torch.cuda.device(grad_out.device):
            check(lib().upsnet_roi_align_backward(ptr(grad_out), ptr(rois), rois.shape[0], B, Cc, H, W, ph, pw, sr, scale,
                                                  ptr(dfeat), stream_ptr(grad_out.device)), "roi_align_backward")
        return dfeat, None, None, None, None, None


# ------------------------------------------------------------------------------------------------
# gradient all-reduce (config #4: "bf16, 8xB200, NCCL allreduce over NVLink")
# ------------------------------------------------------------------------------------------------
class FlatBucketAllReduce:
    """Averages the gradients of `params` over the process group through flat buckets.

    * buckets are filled in REVERSE parameter order (the order backward produces gradients), `bucket_bytes` each;
    * `reduce_dtype` (bf16 on the NCCL path: half the NVLink bytes; the accumulation of 8 ranks in bf16 costs ~3 bits, the
      configuration BASELINE.json names) -- gradients are packed with a cast, reduced with SUM, and unpacked with 1/world;
    * `start()` launches every bucket's all_reduce asynchronously (NCCL: on its own stream, overlapping the optimiser's
      host work and, when called from autograd hooks, the rest of backward); `finish()` waits and writes p.grad back.
    No data-path collective exists at inference; this is the one real exchange step of the path."""

    def __init__(self, params, bucket_bytes=25 << 20, reduce_dtype=torch.bfloat16, group=None):
        self.params = [p for p in params if p.requires_grad]
        self.group, self.dtype = group, reduce_dtype
        esize = torch.empty((), dtype=reduce_dtype).element_size()
        self.buckets, cur, cur_n = [], [], 0
        for p in reversed(self.params):
            cur.append(p); cur_n += p.numel()
            if cur_n * esize >= bucket_bytes:
                self.buckets.append(cur); cur, cur_n = [], 0
        if cur:
            self.buckets.append(cur)
        self._flat, self._work = [None] * len(self.buckets), []

    def start(self):
        assert dist.is_available() and dist.is_initialized()
        self._work = []
        for bi, bucket in enumerate(self.buckets):
            grads = [(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1) for p in bucket]
            flat = torch.cat([g.to(self.dtype) for g in grads])
            self._flat[bi] = flat
            self._work.append(dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True))
        return self

    def finish(self):
        world = dist.get_world_size(self.group)
        for bi, bucket in enumerate(self.buckets):
            self._work[bi].wait()
            flat, off = self._flat[bi], 0
            for p in bucket:
                n = p.numel()
                g = (flat[off:off + n].to(p.dtype) / world).view_as(p)
                if p.grad is None:
                    p.grad = g.clone()
                else:
                    p.grad.copy_(g)
                off += n
        self._work = []
        return self

    def __call__(self):
        return self.start().finish()
