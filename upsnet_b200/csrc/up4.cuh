// up4.cuh -- x4 bilinear up-sampling (align_corners = False) of one row quad, shared by upsample_bilinear_nchw_kernel
// (pool.cu) and the fused panoptic head (panoptic.cu: pan_fuse_kernel<true> evaluates the up-sampled semantic logits on
// the fly instead of reading a materialised [S,4H,4W] tensor).  Both kernels inline THIS function with explicit
// fused-multiply-adds, so the fused head sees bit-identical logits to the ones the stand-alone kernel writes
// (reference: nn.Upsample(scale_factor=4, mode='bilinear'), models/fcn.py:88-101).
#pragma once
#include <cuda_runtime.h>

namespace ups {

struct Up4Row { int y0, y1; float ly, hy; };

// source rows / weights of output row yo (H = source height)
__device__ __forceinline__ Up4Row up4_row(int yo, int H) {
  Up4Row r;
  const float sy = fmaxf(__fmaf_rn(0.25f, (float)yo + 0.5f, -0.5f), 0.f);
  r.y0 = (int)sy;
  r.y1 = r.y0 + (r.y0 < H - 1 ? 1 : 0);
  r.ly = sy - (float)r.y0;
  r.hy = 1.f - r.ly;
  return r;
}

// outputs 4q..4q+3 of one output row from source rows r0 / r1 (W = source width): they read source columns q-1, q, q+1
// only (fractions .625 .875 | .125 .375).  Left border (q == 0): the source index clamps at 0, i.e. weight 1 on column 0.
__device__ __forceinline__ void up4_quad(const float* __restrict__ r0, const float* __restrict__ r1, int q, int W, float ly,
                                         float hy, float (&o)[4]) {
  const int xm = max(q - 1, 0), xp = min(q + 1, W - 1);
  const float a0 = __ldg(r0 + xm), a1 = __ldg(r0 + q), a2 = __ldg(r0 + xp);
  const float b0 = __ldg(r1 + xm), b1 = __ldg(r1 + q), b2 = __ldg(r1 + xp);
  const float l0 = q == 0 ? 0.f : 0.625f, l1 = q == 0 ? 0.f : 0.875f;
  const float t0 = __fmaf_rn(l0, a1, __fmul_rn(1.f - l0, a0)), u0 = __fmaf_rn(l0, b1, __fmul_rn(1.f - l0, b0));
  const float t1 = __fmaf_rn(l1, a1, __fmul_rn(1.f - l1, a0)), u1 = __fmaf_rn(l1, b1, __fmul_rn(1.f - l1, b0));
  const float t2 = __fmaf_rn(0.125f, a2, __fmul_rn(0.875f, a1)), u2 = __fmaf_rn(0.125f, b2, __fmul_rn(0.875f, b1));
  const float t3 = __fmaf_rn(0.375f, a2, __fmul_rn(0.625f, a1)), u3 = __fmaf_rn(0.375f, b2, __fmul_rn(0.625f, b1));
  o[0] = __fmaf_rn(ly, u0, __fmul_rn(hy, t0));
  o[1] = __fmaf_rn(ly, u1, __fmul_rn(hy, t1));
  o[2] = __fmaf_rn(ly, u2, __fmul_rn(hy, t2));
  o[3] = __fmaf_rn(ly, u3, __fmul_rn(hy, t3));
}

}  // namespace ups
