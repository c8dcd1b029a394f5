// backward.cu -- backward kernels of the custom operators for the training configuration (BASELINE config #4,
// SURVEY 2.2 K1-K6, K8).  fp32 NCHW, the layout of the reference's training path; the forward of these operators is
// the fused tensor-core kernel (igemm_tc.cu) or, for bit-stable training, the fp32 tiles.
//
// Deformable convolution (v1 / v2), structure of operators/functions/deform_conv.py:59-108 (mod_deform_conv.py:61-118):
//   d(weight) = dY[Cout,P] * col^T[P,Cin*KHW]      col from upsnet_dcn_im2col      (K1 / K4)
//   d(col)    = W^T[Cin*KHW,Cout] * dY[Cout,P]     library GEMM on the host side
//   d(x)      = scatter(d(col))                    upsnet_dcn_col2im               (K2 / K5: red.global.add.f32)
//   d(offset), d(mask) = reduce over Cin           upsnet_dcn_col2im_coord         (K3 / K6: warp per (tap,pixel), lanes over
//                                                                                   channels, shuffle reduction)
// ROIAlign: upsnet_roi_align_backward (K8): scatter of every output gradient to its gh*gw*4 bilinear taps.
// All kernels are HBM / atomic-throughput bound streaming kernels; col is Cin*KHW*P*4 bytes like the reference's.
#include "common.cuh"

namespace ups {

struct DcnGeom {
  int Cin, H, W, kh, kw, sh, sw, ph, pw, dh, dw, Ho, Wo;
};

// zero-padded bilinear sample (deform_conv_kernel.cu:89-118 deformable_im2col_bilinear): h in (-1, H), w in (-1, W)
__device__ __forceinline__ float dcn_bilinear(const float* __restrict__ pl, int H, int W, float h, float w) {
  const int hl = (int)floorf(h), wl = (int)floorf(w), hh = hl + 1, wh = wl + 1;
  const float lh = h - hl, lw = w - wl, ch = 1.f - lh, cw = 1.f - lw;
  const float v1 = (hl >= 0 && wl >= 0) ? __ldg(pl + hl * W + wl) : 0.f;
  const float v2 = (hl >= 0 && wh <= W - 1) ? __ldg(pl + hl * W + wh) : 0.f;
  const float v3 = (hh <= H - 1 && wl >= 0) ? __ldg(pl + hh * W + wl) : 0.f;
  const float v4 = (hh <= H - 1 && wh <= W - 1) ? __ldg(pl + hh * W + wh) : 0.f;
  return ch * cw * v1 + ch * lw * v2 + lh * cw * v3 + lh * lw * v4;
}

__device__ __forceinline__ void dcn_sample_pos(const DcnGeom& g, const float* __restrict__ offset, int tap, int p, float& h, float& w) {
  const int HoWo = g.Ho * g.Wo;
  const int ho = p / g.Wo, wo = p - ho * g.Wo;
  const int ki = tap / g.kw, kj = tap - ki * g.kw;
  h = (float)(ho * g.sh - g.ph + ki * g.dh) + __ldg(offset + (size_t)(2 * tap) * HoWo + p);
  w = (float)(wo * g.sw - g.pw + kj * g.dw) + __ldg(offset + (size_t)(2 * tap + 1) * HoWo + p);
}

// K1 / K4: col[(c*KHW + tap), p] for one image; thread = (c, p), p fastest (coalesced offset / col accesses)
__global__ void __launch_bounds__(256)
dcn_im2col_kernel(const float* __restrict__ x, const float* __restrict__ offset, const float* __restrict__ mask, DcnGeom g,
                  float* __restrict__ col) {
  const int HoWo = g.Ho * g.Wo, KHW = g.kh * g.kw;
  const long long total = (long long)g.Cin * HoWo;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int p = (int)(t % HoWo), c = (int)(t / HoWo);
    const float* pl = x + (size_t)c * g.H * g.W;
    for (int tap = 0; tap < KHW; ++tap) {
      float h, w;
      dcn_sample_pos(g, offset, tap, p, h, w);
      float v = 0.f;
      if (h > -1.f && w > -1.f && h < (float)g.H && w < (float)g.W) v = dcn_bilinear(pl, g.H, g.W, h, w);
      if (mask) v *= __ldg(mask + (size_t)tap * HoWo + p);
      col[((size_t)c * KHW + tap) * HoWo + p] = v;
    }
  }
}

// K2 / K5: d(x)[c, y, x] += bilinear weight * d(col)[(c,tap), p] (* mask) for the <= 4 in-range corners of the sample
__global__ void __launch_bounds__(256)
dcn_col2im_kernel(const float* __restrict__ dcol, const float* __restrict__ offset, const float* __restrict__ mask, DcnGeom g,
                  float* __restrict__ dx) {
  const int HoWo = g.Ho * g.Wo, KHW = g.kh * g.kw;
  const long long total = (long long)g.Cin * KHW * HoWo;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int p = (int)(t % HoWo);
    const int ct = (int)(t / HoWo), tap = ct % KHW, c = ct / KHW;
    float h, w;
    dcn_sample_pos(g, offset, tap, p, h, w);
    if (!(h > -1.f && w > -1.f && h < (float)g.H && w < (float)g.W)) continue;
    float gv = dcol[t];
    if (mask) gv *= __ldg(mask + (size_t)tap * HoWo + p);
    const int hl = (int)floorf(h), wl = (int)floorf(w), hh = hl + 1, wh = wl + 1;
    const float lh = h - hl, lw = w - wl, ch = 1.f - lh, cw = 1.f - lw;
    float* pl = dx + (size_t)c * g.H * g.W;
    if (hl >= 0 && wl >= 0) atomicAdd(pl + hl * g.W + wl, ch * cw * gv);
    if (hl >= 0 && wh <= g.W - 1) atomicAdd(pl + hl * g.W + wh, ch * lw * gv);
    if (hh <= g.H - 1 && wl >= 0) atomicAdd(pl + hh * g.W + wl, lh * cw * gv);
    if (hh <= g.H - 1 && wh <= g.W - 1) atomicAdd(pl + hh * g.W + wh, lh * lw * gv);
  }
}

// K3 / K6: one WARP per (tap, pixel); lanes stride over the channels, three partial sums (d/dh, d/dw, d/dmask) reduced by
// shuffles.  d(sample)/dh = (x[hh,.] - x[hl,.]) blended along w, d/dw likewise (deform_conv_kernel.cu:149-186).
__global__ void __launch_bounds__(256)
dcn_col2im_coord_kernel(const float* __restrict__ dcol, const float* __restrict__ x, const float* __restrict__ offset,
                        const float* __restrict__ mask, DcnGeom g, float* __restrict__ doffset, float* __restrict__ dmask) {
  const int HoWo = g.Ho * g.Wo, KHW = g.kh * g.kw;
  const int lane = threadIdx.x & 31;
  const long long warps = ((long long)gridDim.x * blockDim.x) >> 5;
  const long long total = (long long)KHW * HoWo;
  for (long long t = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; t < total; t += warps) {
    const int p = (int)(t % HoWo), tap = (int)(t / HoWo);
    float h, w;
    dcn_sample_pos(g, offset, tap, p, h, w);
    const bool inside = h > -1.f && w > -1.f && h < (float)g.H && w < (float)g.W;
    float gh = 0.f, gw = 0.f, gm = 0.f;
    if (inside) {
      const int hl = (int)floorf(h), wl = (int)floorf(w), hh = hl + 1, wh = wl + 1;
      const float lh = h - hl, lw = w - wl, ch = 1.f - lh, cw = 1.f - lw;
      const bool t_ok = hl >= 0, b_ok = hh <= g.H - 1, l_ok = wl >= 0, r_ok = wh <= g.W - 1;
      const float m = mask ? __ldg(mask + (size_t)tap * HoWo + p) : 1.f;
      for (int c = lane; c < g.Cin; c += 32) {
        const float* pl = x + (size_t)c * g.H * g.W;
        const float v1 = (t_ok && l_ok) ? __ldg(pl + hl * g.W + wl) : 0.f;
        const float v2 = (t_ok && r_ok) ? __ldg(pl + hl * g.W + wh) : 0.f;
        const float v3 = (b_ok && l_ok) ? __ldg(pl + hh * g.W + wl) : 0.f;
        const float v4 = (b_ok && r_ok) ? __ldg(pl + hh * g.W + wh) : 0.f;
        const float gc = dcol[((size_t)c * KHW + tap) * HoWo + p];
        gh += gc * m * (cw * (v3 - v1) + lw * (v4 - v2));
        gw += gc * m * (ch * (v2 - v1) + lh * (v4 - v3));
        gm += gc * (ch * cw * v1 + ch * lw * v2 + lh * cw * v3 + lh * lw * v4);
      }
    }
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) {
      gh += __shfl_xor_sync(0xffffffffu, gh, s);
      gw += __shfl_xor_sync(0xffffffffu, gw, s);
      gm += __shfl_xor_sync(0xffffffffu, gm, s);
    }
    if (lane == 0) {
      doffset[(size_t)(2 * tap) * HoWo + p] = gh;
      doffset[(size_t)(2 * tap + 1) * HoWo + p] = gw;
      if (dmask) dmask[(size_t)tap * HoWo + p] = gm;
    }
  }
}

// K8: thread per output gradient element (n, c, ph, pw) -> atomicAdd to the four taps of each of its gh*gw samples
// (roi_align_kernel.cu:238-348; sample positions exactly as the forward, roi_align.cu roi_sample)
__global__ void __launch_bounds__(256)
roi_align_backward_kernel(const float* __restrict__ dout, const float* __restrict__ rois, int R, int C, int H, int W, int PH,
                          int PW, int sr, float scale, float* __restrict__ dfeat) {
  const long long total = (long long)R * C * PH * PW;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int pw = (int)(idx % PW), ph = (int)((idx / PW) % PH);
    const int c = (int)((idx / PW / PH) % C), n = (int)(idx / PW / PH / C);
    const float* r = rois + (size_t)n * 5;
    const int b = (int)roundf(r[0]);
    const float rsw = r[1] * scale, rsh = r[2] * scale, rew = r[3] * scale, reh = r[4] * scale;
    const float rw = fmaxf(rew - rsw, 1.f), rh = fmaxf(reh - rsh, 1.f);
    const float bsh = rh / (float)PH, bsw = rw / (float)PW;
    const int gh = sr > 0 ? sr : (int)ceilf(rh / PH), gw = sr > 0 ? sr : (int)ceilf(rw / PW);
    const float gv = dout[idx] / (float)(gh * gw);
    float* d = dfeat + ((size_t)b * C + c) * H * W;
    for (int iy = 0; iy < gh; ++iy) {
      float y = rsh + ph * bsh + (float)(iy + .5f) * bsh / (float)gh;
      for (int ix = 0; ix < gw; ++ix) {
        float x = rsw + pw * bsw + (float)(ix + .5f) * bsw / (float)gw;
        float yy = y;
        if (yy < -1.0f || yy > (float)H || x < -1.0f || x > (float)W) continue;
        if (yy <= 0) yy = 0;
        if (x <= 0) x = 0;
        int yl = (int)yy, xl = (int)x, yh, xh;
        if (yl >= H - 1) { yh = yl = H - 1; yy = (float)yl; } else yh = yl + 1;
        if (xl >= W - 1) { xh = xl = W - 1; x = (float)xl; } else xh = xl + 1;
        const float ly = yy - yl, lx = x - xl, hy = 1.f - ly, hx = 1.f - lx;
        atomicAdd(d + yl * W + xl, gv * hy * hx);
        atomicAdd(d + yl * W + xh, gv * hy * lx);
        atomicAdd(d + yh * W + xl, gv * ly * hx);
        atomicAdd(d + yh * W + xh, gv * ly * lx);
      }
    }
  }
}

static int dcn_geom(DcnGeom* g, int Cin, int H, int W, int kh, int kw, int sh, int sw, int ph, int pw, int dh, int dw) {
  if (Cin <= 0 || H <= 0 || W <= 0 || kh <= 0 || kw <= 0 || sh <= 0 || sw <= 0 || ph < 0 || pw < 0 || dh <= 0 || dw <= 0)
    return UPSNET_E_BADARG;
  g->Cin = Cin; g->H = H; g->W = W; g->kh = kh; g->kw = kw; g->sh = sh; g->sw = sw; g->ph = ph; g->pw = pw; g->dh = dh; g->dw = dw;
  g->Ho = conv_out_size(H, ph, dh, kh, sh);
  g->Wo = conv_out_size(W, pw, dw, kw, sw);
  if (g->Ho <= 0 || g->Wo <= 0) return UPSNET_E_BADARG;
  if ((long long)Cin * kh * kw * g->Ho * g->Wo >= (1ll << 40)) return UPSNET_E_UNSUPPORTED;
  return 0;
}

static unsigned blocks_for(long long total) {
  long long b = (total + 255) / 256;
  if (b > (long long)kNumSMs * 32) b = (long long)kNumSMs * 32;
  return (unsigned)(b > 0 ? b : 1);
}

}  // namespace ups

extern "C" int upsnet_dcn_im2col(const float* x, const float* offset, const float* mask, int Cin, int H, int W, int kh, int kw,
                                 int sh, int sw, int ph, int pw, int dh, int dw, float* col, void* stream) {
  using namespace ups;
  if (!x || !offset || !col) return UPSNET_E_BADARG;
  DcnGeom g;
  const int rc = dcn_geom(&g, Cin, H, W, kh, kw, sh, sw, ph, pw, dh, dw);
  if (rc) return rc;
  dcn_im2col_kernel<<<blocks_for((long long)Cin * g.Ho * g.Wo), 256, 0, (cudaStream_t)stream>>>(x, offset, mask, g, col);
  UPS_CHECK_LAUNCH();
  return 0;
}

extern "C" int upsnet_dcn_col2im(const float* dcol, const float* offset, const float* mask, int Cin, int H, int W, int kh,
                                 int kw, int sh, int sw, int ph, int pw, int dh, int dw, float* dx, void* stream) {
  using namespace ups;
  if (!dcol || !offset || !dx) return UPSNET_E_BADARG;
  DcnGeom g;
  const int rc = dcn_geom(&g, Cin, H, W, kh, kw, sh, sw, ph, pw, dh, dw);
  if (rc) return rc;
  UPS_CUDA(cudaMemsetAsync(dx, 0, sizeof(float) * (size_t)Cin * H * W, (cudaStream_t)stream));
  dcn_col2im_kernel<<<blocks_for((long long)Cin * kh * kw * g.Ho * g.Wo), 256, 0, (cudaStream_t)stream>>>(dcol, offset, mask, g, dx);
  UPS_CHECK_LAUNCH();
  return 0;
}

extern "C" int upsnet_dcn_col2im_coord(const float* dcol, const float* x, const float* offset, const float* mask, int Cin,
                                       int H, int W, int kh, int kw, int sh, int sw, int ph, int pw, int dh, int dw,
                                       float* doffset, float* dmask, void* stream) {
  using namespace ups;
  if (!dcol || !x || !offset || !doffset || (mask && !dmask)) return UPSNET_E_BADARG;
  DcnGeom g;
  const int rc = dcn_geom(&g, Cin, H, W, kh, kw, sh, sw, ph, pw, dh, dw);
  if (rc) return rc;
  dcn_col2im_coord_kernel<<<blocks_for((long long)kh * kw * g.Ho * g.Wo * 32), 256, 0, (cudaStream_t)stream>>>(
      dcol, x, offset, mask, g, doffset, mask ? dmask : nullptr);
  UPS_CHECK_LAUNCH();
  return 0;
}

extern "C" int upsnet_roi_align_backward(const float* dout, const float* rois, int R, int B, int C, int H, int W, int PH,
                                         int PW, int sampling_ratio, float spatial_scale, float* dfeat, void* stream) {
  using namespace ups;
  if (!dout || !dfeat || (!rois && R > 0) || R < 0 || B <= 0 || C <= 0 || H <= 0 || W <= 0 || PH <= 0 || PW <= 0)
    return UPSNET_E_BADARG;
  UPS_CUDA(cudaMemsetAsync(dfeat, 0, sizeof(float) * (size_t)B * C * H * W, (cudaStream_t)stream));
  if (R == 0) return 0;
  roi_align_backward_kernel<<<blocks_for((long long)R * C * PH * PW), 256, 0, (cudaStream_t)stream>>>(
      dout, rois, R, C, H, W, PH, PW, sampling_ratio, spatial_scale, dfeat);
  UPS_CHECK_LAUNCH();
  return 0;
}
