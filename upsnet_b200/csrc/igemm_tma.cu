// igemm_tma.cu -- TMA-fed tcgen05 implicit-GEMM convolution for sm_100a (dense, stride 1, bf16 NHWC in / out).
//
// The layers that dominate the UPSNet backbone / FPN / RPN / mask head / FC path (1x1 and 3x3, stride 1,
// Cin % 64 == 0, Cout % 64 == 0; reference: models/resnet.py, models/fpn.py, models/rcnn.py) need no gather
// arithmetic at all once the activations are NHWC bf16: a tile of output pixels is a bw x bh x bn BOX of the
// 4-D tensor (C, W, H, N), and filter tap (ki, kj) of that tile is the SAME box shifted by (kj*dil - pad,
// ki*dil - pad) with out-of-range pixels reading zero.  That is exactly one tiled-mode TMA load per k-block:
//
//   A[<=128 pixels x 64 ch]  cp.async.bulk.tensor.4d  box (64, bw, bh, bn) at (c0, w0 + kj*dw - pw, h0 + ki*dh - ph, n0)
//   B[BN couts x 64 k]       cp.async.bulk.tensor.2d  box (64, BN) of the packed weights [Cout][tap*Cin + c]
//   D[128 x BN] fp32 in TMEM (two buffers)  +=  A * B^T     tcgen05.mma kind::f16, one elected thread
//   epilogue (8 warps, thread = accumulator row): tcgen05.ld -> +bias (+residual slab, TMA-prefetched) -> ReLU
//            -> bf16 -> SWIZZLE_128B slab in smem -> cp.async.bulk.tensor.4d store (clips the box at the borders)
//
// Both operands land in the K-major SWIZZLE_128B layout the UMMA descriptors expect (the TMA swizzle mode and
// the smem descriptor's layout type are the same permutation), so no thread touches the operands: 3 service
// warps (TMA loads, MMA issue, residual loads) + 8 epilogue warps per persistent CTA, one CTA per SM.
// Roofline: tensor pipe for the 3x3 layers (2*P*Cout*Cin*9 flop), HBM for the 1x1 (+residual) layers
// (x + residual + y bytes); per-SM L2->smem operand traffic is (128 + BN) * 128 B per k-block.
#include <cuda.h>   // CUtensorMap + enums only; the encoder is resolved at run time (no libcuda link dependency)
#include <cuda_bf16.h>
#include <cstdlib>

#include "common.cuh"
#include "tc_ptx.cuh"
#include "tc_params.cuh"

namespace ups {

constexpr int TM_EPI_WARPS = 8;
constexpr int TM_WARP_TMA = 8, TM_WARP_MMA = 9, TM_WARP_RES = 10;
constexpr int TM_THREADS = 11 * 32;
constexpr int TM_MAX_STAGES = 8;
constexpr int TM_SLAB_BYTES = 128 * 128;   // 128 rows x 64 bf16

struct TmaGeom {
  const float* bias;
  int N, Ho, Wo, Cout, Cin;
  int kw, KHW, ph, pw, dh, dw;
  int bw, bh, bn;                       // M-tile box: pixels along W, along H, images (bw*bh*bn <= 128)
  int tiles_w, tiles_h, tiles_n, n_tiles;
  int BN, stages, relu, has_res;
  int res_up2;                          // residual = half-resolution map, nearest 2x up-sampling (FPN top-down)
  // direct-store epilogue (small / odd Cout, fp32 or NCHW outputs: offset convs, RPN / score / mask-logit heads)
  int direct, y_bf16, out_nhwc;
  void* y;
  int stem;                             // stride-2 tiny-Cin stem: A boxes come from the packed / padded image (5-D map)
  // halo mode (k x k, stride 1): one (16 x PH)-pixel input PATCH per (tile, channel chunk) feeds all taps -- the A operand
  // of tap (ky, kx) is the patch seen through a descriptor that starts (ky*dh*16 + kx*dw) rows further (tile = 8 x 16 px)
  int halo, patch_rows, pstages, kh;
  int rotate;                           // tile-dependent start of the K loop (see producer)
  int wres, wtiles;                     // halo mode with the n-tile's whole weight set (wtiles tiles) resident in smem, loaded once per CTA
  int sig_from;                         // direct epilogue: channels >= sig_from get a logistic sigmoid (-1 = none)
  int dbg;                              // timing experiments only (UPSNET_TMA_DEBUG): 1 alternate accumulators, 2 one MMA per k-block, 3 no MMAs
  // PAIR mode (precision bf16x3 on the TMA kernel): activations are hi/lo bf16 PAIRS -- an NHWC tensor with 2*C channels,
  // channels [0,C) = bf16(x), [C,2C) = bf16(x - hi) -- weights are the packed hi/lo planes, every k-slice issues three
  // MMAs (lo*hi, hi*lo, hi*hi) and the epilogue splits the fp32 result into a pair again.
  int x3;
  int x_lo;                             // channel coordinate of the input's lo plane (= Cin)
  int w_lo;                             // row coordinate of the weight lo plane (= Cout_pad)
  int pg;                               // output / residual pair group G: channels stored [hi G][lo G] per group (G = Cout normally)
  int res_inplace;                      // residual slabs are TMA-loaded into the (double-buffered) output slabs and updated in place
  int opairs;                           // output slab pairs that alternate (2, or 1 when shared memory is short)
  // 2-CTA kernel, N tile <= 64: the hi*hi and hi*lo products share ONE instruction -- B operand = [W_hi ; W_lo] (N = 2 BN, CTA 0
  // stages the hi rows, CTA 1 the lo rows), accumulated in columns [0, 2 BN); lo*hi goes to columns [0, BN); the epilogue adds
  // the two column groups.  A tcgen05.mma of M = 256 takes ~100 cycles whatever N <= 128 is (measured: the 18-channel offset
  // convs and the 64->64 3x3 convs run at 12 instructions x ~100 cycles per k-block), so two instructions per K slice instead
  // of three is a third less time on these instruction-bound layers.
  int wide;
};

// ---- PTX: TMA (bulk tensor) copies ----
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_load_5d(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1, int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* tm, uint32_t src, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
      ::"l"(reinterpret_cast<uint64_t>(tm)), "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* tm) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(tm)) : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int count) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, const uint4& v) {
  asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// pair tensors: channel coordinate of channel n's hi value when channels are stored [hi G][lo G] per group of G (lo = +G)
__device__ __forceinline__ int pair_chan(int n, int G) { return (n / G) * 2 * G + (n % G); }

struct TmaSmem {
  uint32_t stages, out, res, a_bytes, b_bytes, stage_bytes, total;
  uint32_t patch, patch_bytes;                               // halo mode: patch ring in front of a B-only ring
  uint32_t a_half, b_half, patch_half;                       // pair mode: offset of the lo tile inside an A / B / patch slot
  uint32_t oslabs, res_slab;                                 // pair mode: number of (hi, lo) output slab pairs; bytes per residual slab
};
__host__ __device__ inline TmaSmem tma_smem_layout(int BN, int stages, bool has_res, int patch_rows = 0, int pstages = 0,
                                                  bool direct = false, bool x3 = false, bool res_up2 = false,
                                                  bool res_inplace = false, int opairs = 2) {
  TmaSmem s;
  const uint32_t mul = x3 ? 2u : 1u;
  s.a_half = 128 * 128; s.b_half = (uint32_t)BN * 128; s.patch_half = (uint32_t)patch_rows * 128u;
  s.a_bytes = patch_rows ? 0u : s.a_half * mul;
  s.b_bytes = s.b_half * mul;
  s.stage_bytes = s.a_bytes + s.b_bytes;
  s.patch = 1024;                                            // barriers live in the first KB
  s.patch_bytes = s.patch_half * mul;                        // multiple of 2048 (16-pixel patch rows)
  s.stages = s.patch + s.patch_bytes * (uint32_t)pstages;
  s.out = s.stages + s.stage_bytes * (uint32_t)stages;
  s.res_slab = (x3 && res_up2) ? 4096u : (uint32_t)TM_SLAB_BYTES;
  if (x3) {
    // pair mode: the eight epilogue warps work on ONE 64-channel slab pair at a time (two pairs alternate so that a TMA
    // store can still be reading the first while the second is written); in-place residual: one pair per slab and buffer
    s.oslabs = direct ? 0u : (res_inplace ? 2u * (uint32_t)(BN / 64) : (uint32_t)opairs);
    s.res = s.out + s.oslabs * 2u * TM_SLAB_BYTES;
    s.total = s.res + ((has_res && !res_inplace) ? 2u * (uint32_t)(BN / 64) * 2u * s.res_slab : 0u);
    return s;
  }
  s.oslabs = 0;
  const uint32_t out_slabs = direct ? 0 : (BN == 64 ? 1 : 2);   // one output slab per epilogue group (TMA-store epilogue only)
  s.res = s.out + out_slabs * TM_SLAB_BYTES;
  s.total = s.res + (has_res ? 2u * (uint32_t)(BN / 64) * TM_SLAB_BYTES : 0u);   // two residual buffers (prefetch)
  return s;
}

// Roles (11 warps): warps 0-7 epilogue (TMEM lane quadrant = warp & 3; column half = warp >> 2), warp 8 TMA
// operand loads, warp 9 MMA issue + TMEM ownership, warp 10 residual-slab loads.
// Barriers: full[s]/empty[s] smem ring (TMA <-> MMA), tfull[b]/tempty[b] TMEM buffers (MMA <-> epilogue),
// rfull/rempty residual slabs (TMA <-> epilogue).
__global__ void __launch_bounds__(TM_THREADS, 1)
igemm_tma_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_w,
                 const __grid_constant__ CUtensorMap tm_y, const __grid_constant__ CUtensorMap tm_r, const TmaGeom g) {
  extern __shared__ __align__(1024) uint8_t smem_dyn[];
  const uint32_t raw = smem_u32(smem_dyn);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* sm = smem_dyn + (base - raw);
  const TmaSmem L = tma_smem_layout(g.BN, g.wres ? g.wtiles : g.stages, g.has_res != 0, g.halo ? g.patch_rows : 0, g.pstages,
                                    g.direct != 0, g.x3 != 0, g.res_up2 != 0, g.res_inplace != 0, g.opairs);
  const uint32_t bar_full = base, bar_empty = base + 8 * TM_MAX_STAGES;
  const uint32_t bar_tfull = bar_empty + 8 * TM_MAX_STAGES, bar_tempty = bar_tfull + 16;
  const uint32_t bar_rfull = bar_tempty + 16, bar_rempty = bar_rfull + 16;     // two residual buffers
  const uint32_t bar_pfull = bar_rempty + 16, bar_pempty = bar_pfull + 32;     // up to four patch slots (halo mode)
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(sm + 8 * (2 * TM_MAX_STAGES + 16));

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int cchunks = g.Cin / 64;
  const int num_kb = g.KHW * cchunks;
  const long long m_tiles = (long long)g.tiles_w * g.tiles_h * g.tiles_n;
  const long long num_tiles = m_tiles * g.n_tiles;
  const uint32_t box_bytes = (uint32_t)(g.bw * g.bh * g.bn) * 128u;
  // wide-B (pair mode, N tile <= 64, per-tap boxes / stem): B operand = the contiguous [W_hi ; W_lo] tile of the stage (N = 2 BN)
  const uint32_t acc_cols = (uint32_t)(g.wide ? 2 * g.BN : g.BN);      // TMEM columns of one accumulator buffer
  uint32_t tmem_cols = 32;
  while (tmem_cols < 2 * acc_cols) tmem_cols <<= 1;

  if (warp == TM_WARP_MMA) {
    if (lane == 0) {
      for (int s = 0; s < g.stages; ++s) {
        mbar_init(bar_full + 8 * s, 1);
        mbar_init(bar_empty + 8 * s, 1);
      }
      for (int b = 0; b < 2; ++b) {
        mbar_init(bar_tfull + 8 * b, 1);
        mbar_init(bar_tempty + 8 * b, TM_EPI_WARPS);
      }
      for (int b = 0; b < 2; ++b) {
        mbar_init(bar_rfull + 8 * b, 1);
        mbar_init(bar_rempty + 8 * b, g.res_inplace ? 1 : TM_EPI_WARPS);
      }
      for (int b = 0; b < 4; ++b) {
        mbar_init(bar_pfull + 8 * b, 1);
        mbar_init(bar_pempty + 8 * b, 1);
      }
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc(smem_u32(tmem_ptr_smem), tmem_cols);
  } else if (warp == TM_WARP_TMA && lane == 0) {
    prefetch_tmap(&tm_x);
    prefetch_tmap(&tm_w);
    prefetch_tmap(&tm_y);
    if (g.has_res || (g.stem && g.x3)) prefetch_tmap(&tm_r);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == TM_WARP_TMA) {
    // =============================== OPERAND LOADS ===============================
    if (lane == 0) {
      // ring position / phase are running counters: no division on the per-k-block path of this single thread
      uint32_t s = 0, ph = 0, sp = 0, php = 0;
      bool first_tile = true;
      uint32_t a_dst = base + L.stages;
      const uint32_t tx_bytes = (g.x3 ? 2u * box_bytes : box_bytes) + L.b_bytes;
      for (long long tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int nt = (int)(tile % g.n_tiles);
        const long long mt = tile / g.n_tiles;
        const int w0 = (int)(mt % g.tiles_w) * g.bw;
        const int h0 = (int)((mt / g.tiles_w) % g.tiles_h) * g.bh;
        const int i0 = (int)(mt / ((long long)g.tiles_w * g.tiles_h)) * g.bn;
        const int n0 = nt * g.BN;
        if (g.halo) {
          // Units (tile, channel chunk) in order.  The PATCH of unit u+1 is requested before the weight tiles of unit u
          // (its latency would otherwise be exposed once per unit: the weight ring is shorter than a unit's KHW tiles).
          // Resident-weight mode: the n-tile's whole weight set is loaded once, up front.
          if (first_tile) {
            first_tile = false;
            if (g.wres) {
              mbar_arrive_expect_tx(bar_full, (uint32_t)(g.KHW * cchunks) * L.b_bytes);
              for (int cc = 0; cc < cchunks; ++cc)
                for (int tap = 0; tap < g.KHW; ++tap) {
                  const uint32_t wd = base + L.stages + (uint32_t)(cc * g.KHW + tap) * L.b_bytes;
                  tma_load_2d(wd, &tm_w, bar_full, tap * g.Cin + cc * 64, n0);
                  if (g.x3) tma_load_2d(wd + L.b_half, &tm_w, bar_full, tap * g.Cin + cc * 64, g.w_lo + n0);
                }
            }
            mbar_wait(bar_pempty + 8 * sp, php ^ 1u);
            mbar_arrive_expect_tx(bar_pfull + 8 * sp, L.patch_bytes);
            tma_load_4d(base + L.patch + sp * L.patch_bytes, &tm_x, bar_pfull + 8 * sp, 0, w0 - g.pw, h0 - g.ph, i0);
            if (g.x3) tma_load_4d(base + L.patch + sp * L.patch_bytes + L.patch_half, &tm_x, bar_pfull + 8 * sp, g.x_lo, w0 - g.pw, h0 - g.ph, i0);
            if (++sp == (uint32_t)g.pstages) { sp = 0; php ^= 1u; }
          }
          const long long ntile = tile + gridDim.x;
          for (int cc = 0; cc < cchunks; ++cc) {
            // next unit: next chunk of this tile, or chunk 0 of this CTA's next tile
            if (cc + 1 < cchunks || ntile < num_tiles) {
              int nw0 = w0, nh0 = h0, ni0 = i0, ncc = cc + 1;
              if (cc + 1 == cchunks) {
                const long long nmt = ntile / g.n_tiles;
                nw0 = (int)(nmt % g.tiles_w) * g.bw;
                nh0 = (int)((nmt / g.tiles_w) % g.tiles_h) * g.bh;
                ni0 = (int)(nmt / ((long long)g.tiles_w * g.tiles_h)) * g.bn;
                ncc = 0;
              }
              mbar_wait(bar_pempty + 8 * sp, php ^ 1u);
              mbar_arrive_expect_tx(bar_pfull + 8 * sp, L.patch_bytes);
              tma_load_4d(base + L.patch + sp * L.patch_bytes, &tm_x, bar_pfull + 8 * sp, ncc * 64, nw0 - g.pw, nh0 - g.ph, ni0);
              if (g.x3) tma_load_4d(base + L.patch + sp * L.patch_bytes + L.patch_half, &tm_x, bar_pfull + 8 * sp, g.x_lo + ncc * 64, nw0 - g.pw, nh0 - g.ph, ni0);
              if (++sp == (uint32_t)g.pstages) { sp = 0; php ^= 1u; }
            }
            if (g.wres) continue;
            for (int t_ = 0; t_ < g.KHW; ++t_) {
              int tap = t_ + (g.rotate ? (int)(mt % g.KHW) : 0);     // same rotation as the MMA loop
              if (tap >= g.KHW) tap -= g.KHW;
              const uint32_t bf = bar_full + 8 * s;
              mbar_wait(bar_empty + 8 * s, ph ^ 1u);
              mbar_arrive_expect_tx(bf, L.b_bytes);
              tma_load_2d(a_dst, &tm_w, bf, tap * g.Cin + cc * 64, n0);
              if (g.x3) tma_load_2d(a_dst + L.b_half, &tm_w, bf, tap * g.Cin + cc * 64, g.w_lo + n0);
              a_dst += L.stage_bytes;
              if (++s == (uint32_t)g.stages) { s = 0; ph ^= 1u; a_dst = base + L.stages; }
            }
          }
          continue;
        }
        // Optional (UPSNET_TMA_ROTATE=1, default off): start the K loop at a tile-dependent k-block and wrap.  Measured
        // on B200: SLOWER by 2.5 % end to end -- lock-step CTAs asking for the same weight tile at the same moment is
        // what lets L2 merge their requests, so the natural order stays the default.
        int kbr = g.rotate ? (int)((mt * 5) % num_kb) : 0;
        int tap0 = kbr / cchunks;
        int cc = kbr - tap0 * cchunks, ki = tap0 / g.kw, kj = tap0 - ki * g.kw;
        int cw = w0 - g.pw + kj * g.dw, ch = h0 - g.ph + ki * g.dh;         // box origin of the current tap
        for (int kb = 0; kb < num_kb; ++kb) {
          const uint32_t bf = bar_full + 8 * s;
          mbar_wait(bar_empty + 8 * s, ph ^ 1u);
          mbar_arrive_expect_tx(bf, tx_bytes);
          if (g.stem)   // k-block = filter row ky: 8 pixels x 8 channels per output pixel, input row 2*ho + ky of the padded image
            tma_load_5d(a_dst, &tm_x, bf, 0, w0, kbr & 1, h0 + (kbr >> 1), i0);
          else
            tma_load_4d(a_dst, &tm_x, bf, cc * 64, cw, ch, i0);
          tma_load_2d(a_dst + L.a_bytes, &tm_w, bf, kbr * 64, n0);
          if (g.x3) {
            if (g.stem)   // pair stem: the lo plane of the packed image has its own tensor map (tm_r is free: no residual)
              tma_load_5d(a_dst + L.a_half, &tm_r, bf, 0, w0, kbr & 1, h0 + (kbr >> 1), i0);
            else
              tma_load_4d(a_dst + L.a_half, &tm_x, bf, g.x_lo + cc * 64, cw, ch, i0);
            tma_load_2d(a_dst + L.a_bytes + L.b_half, &tm_w, bf, kbr * 64, g.w_lo + n0);
          }
          if (++kbr == num_kb) { kbr = 0; cc = 0; kj = 0; cw = w0 - g.pw; ch = h0 - g.ph; }
          else if (++cc == cchunks) {
            cc = 0; cw += g.dw;
            if (++kj == g.kw) { kj = 0; cw = w0 - g.pw; ch += g.dh; }
          }
          a_dst += L.stage_bytes;
          if (++s == (uint32_t)g.stages) { s = 0; ph ^= 1u; a_dst = base + L.stages; }
        }
      }
    }
    __syncwarp();
  } else if (warp == TM_WARP_MMA) {
    // =============================== MMA ISSUER ===============================
    if (lane == 0) {
      const uint32_t idesc = umma_idesc(128, g.BN), idesc_w = umma_idesc(128, 2 * g.BN);
      // smem descriptors: constant high word (SBO = 1024 B, version 1, SWIZZLE_128B), the low word carries
      // (address >> 4) and is advanced by running adds -- the issue loop of this single thread paces every tile
      // whose MMAs are short (N <= 128), so it is kept to a handful of instructions per k-block.
      const uint32_t desc_hi = (uint32_t)(1024 >> 4) | (1u << 14) | (2u << 29);
      const uint32_t a_lo0 = ((base + L.stages) >> 4) & 0x3fffu, stage16 = L.stage_bytes >> 4, a16 = L.a_bytes >> 4;
      const uint32_t ah16 = L.a_half >> 4, bh16 = L.b_half >> 4, ph16 = L.patch_half >> 4;   // pair mode: lo-tile offsets
      uint32_t s = 0, ph = 0, a_lo = a_lo0, ti_local = 0, sp = 0, php = 0;
      for (long long tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++ti_local) {
        const uint32_t buf = ti_local & 1u, use = ti_local >> 1;
        mbar_wait(bar_tempty + 8 * buf, (use & 1u) ^ 1u);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + buf * acc_cols;
        uint32_t acc = 0u;
        if (g.halo) {
          // A descriptors: 8-row groups = 8 consecutive patch pixels of one patch row, group stride = one patch row
          // (16 px = 2048 B); a tap shifts the START by whole 128-byte rows.  Measured on B200: the SWIZZLE_128B XOR is
          // taken from the shared-memory ADDRESS bits [7,10) -- exactly what the TMA write used -- so a start that is
          // not 1024-byte aligned needs no base-offset (field = 0; a non-zero value double-shifts and corrupts the tile).
          const uint32_t hi_a0 = (uint32_t)(2048 >> 4) | (1u << 14) | (2u << 29);
          if (g.wres && ti_local == 0) {     // the resident weight set: one barrier, once per CTA
            mbar_wait(bar_full, 0u);
            tc_fence_after();
          }
          for (int cc = 0; cc < cchunks; ++cc) {
            mbar_wait(bar_pfull + 8 * sp, php);
            tc_fence_after();
            const uint32_t patch = base + L.patch + sp * L.patch_bytes;
            int ky = 0, kx = 0;
            if (g.wres) {
              uint32_t b_lo = a_lo0 + (uint32_t)(cc * g.KHW) * (L.b_bytes >> 4);
              for (int tap = 0; tap < g.KHW; ++tap) {
                const uint32_t a_start = patch + (uint32_t)((ky * g.dh * 16 + kx * g.dw) * 128);
                const uint32_t pa_lo = (a_start >> 4) & 0x3fffu;
#pragma unroll
                for (uint32_t k = 0; k < 4; ++k) {
                  if (g.x3) {
                    umma_bf16_lohi2(tmem_d, pa_lo + ph16 + 2 * k, hi_a0, b_lo + 2 * k, desc_hi, idesc, acc);
                    umma_bf16_lohi2(tmem_d, pa_lo + 2 * k, hi_a0, b_lo + bh16 + 2 * k, desc_hi, idesc, 1u);
                    acc = 1u;
                  }
                  umma_bf16_lohi2(tmem_d, pa_lo + 2 * k, hi_a0, b_lo + 2 * k, desc_hi, idesc, acc);
                  acc = 1u;
                }
                b_lo += L.b_bytes >> 4;
                if (++kx == g.kw) { kx = 0; ++ky; }
              }
              umma_commit(bar_pempty + 8 * sp);
              if (++sp == (uint32_t)g.pstages) { sp = 0; php ^= 1u; }
              continue;
            }
            const int trot = g.rotate ? (int)((tile / g.n_tiles) % g.KHW) : 0;
            ky = trot / g.kw; kx = trot - ky * g.kw;
            for (int t_ = 0; t_ < g.KHW; ++t_) {
              mbar_wait(bar_full + 8 * s, ph);
              tc_fence_after();
              const uint32_t a_start = patch + (uint32_t)((ky * g.dh * 16 + kx * g.dw) * 128);
              const uint32_t pa_lo = (a_start >> 4) & 0x3fffu, pa_hi = hi_a0;
#pragma unroll
              for (uint32_t k = 0; k < 4; ++k) {
                if (g.x3) {
                  umma_bf16_lohi2(tmem_d, pa_lo + ph16 + 2 * k, pa_hi, a_lo + 2 * k, desc_hi, idesc, acc);
                  umma_bf16_lohi2(tmem_d, pa_lo + 2 * k, pa_hi, a_lo + bh16 + 2 * k, desc_hi, idesc, 1u);
                  acc = 1u;
                }
                umma_bf16_lohi2(tmem_d, pa_lo + 2 * k, pa_hi, a_lo + 2 * k, desc_hi, idesc, acc);
                acc = 1u;
              }
              umma_commit(bar_empty + 8 * s);
              a_lo += stage16;
              if (++s == (uint32_t)g.stages) { s = 0; ph ^= 1u; a_lo = a_lo0; }
              if (++kx == g.kw) { kx = 0; if (++ky == g.kh) ky = 0; }
            }
            umma_commit(bar_pempty + 8 * sp);     // every tap of this chunk has been issued: the patch slot may be refilled
            if (++sp == (uint32_t)g.pstages) { sp = 0; php ^= 1u; }
          }
          umma_commit(bar_tfull + 8 * buf);
          continue;
        }
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(bar_full + 8 * s, ph);
          tc_fence_after();
          const uint32_t b_lo = a_lo + a16;
#pragma unroll
          for (uint32_t k = 0; k < 4; ++k) {       // 16 bf16 = 32 bytes = 2 descriptor units inside the swizzle span
            if (g.dbg == 3 || (g.dbg == 2 && k)) continue;
            const uint32_t td = (g.dbg == 1 && (k & 1)) ? (tmem_base + (buf ^ 1u) * acc_cols) : tmem_d;
            if (g.wide) {    // two instructions per K slice: hi * [hi ; lo] -> columns [0, 2 BN), lo * hi -> columns [0, BN)
              umma_bf16_lohi(td, a_lo + 2 * k, b_lo + 2 * k, desc_hi, idesc_w, acc);
              umma_bf16_lohi(td, a_lo + ah16 + 2 * k, b_lo + 2 * k, desc_hi, idesc, 1u);
              acc = 1u;
              continue;
            }
            if (g.x3) {      // (hi + lo) * (hi + lo) without the lo * lo term: relative error ~2^-16
              umma_bf16_lohi(td, a_lo + ah16 + 2 * k, b_lo + 2 * k, desc_hi, idesc, acc);
              umma_bf16_lohi(td, a_lo + 2 * k, b_lo + bh16 + 2 * k, desc_hi, idesc, 1u);
              acc = 1u;
            }
            umma_bf16_lohi(td, a_lo + 2 * k, b_lo + 2 * k, desc_hi, idesc, acc);
            acc = 1u;
          }
          umma_commit(bar_empty + 8 * s);
          a_lo += stage16;
          if (++s == (uint32_t)g.stages) { s = 0; ph ^= 1u; a_lo = a_lo0; }
        }
        umma_commit(bar_tfull + 8 * buf);
      }
    }
    __syncwarp();
  } else if (warp == TM_WARP_RES) {
    // =============================== RESIDUAL LOADS ===============================
    if (lane == 0 && g.has_res) {
      uint32_t ti_local = 0;
      const int slabs = g.BN / 64;
      for (long long tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++ti_local) {
        const int nt = (int)(tile % g.n_tiles);
        const long long mt = tile / g.n_tiles;
        const int w0 = (int)(mt % g.tiles_w) * g.bw;
        const int h0 = (int)((mt / g.tiles_w) % g.tiles_h) * g.bh;
        const int i0 = (int)(mt / ((long long)g.tiles_w * g.tiles_h)) * g.bn;
        // residual slabs are double-buffered: tile t+1's residual streams in from HBM while tile t's epilogue runs
        const uint32_t rb = ti_local & 1u, ruse = ti_local >> 1;
        const uint32_t rdst = base + L.res + rb * (uint32_t)slabs * TM_SLAB_BYTES;
        mbar_wait(bar_rempty + 8 * rb, (ruse & 1u) ^ 1u);
        if (g.x3) {
          // pair mode: (hi, lo) slab per 64 channels; in-place mode lands them in this tile's output slab pairs
          const uint32_t lo_off = g.res_inplace ? (uint32_t)TM_SLAB_BYTES : L.res_slab;
          const uint32_t pdst = g.res_inplace ? base + L.out + rb * (uint32_t)slabs * 2u * TM_SLAB_BYTES
                                              : base + L.res + rb * (uint32_t)slabs * 2u * L.res_slab;
          mbar_arrive_expect_tx(bar_rfull + 8 * rb, (g.res_up2 ? box_bytes / 4 : box_bytes) * (uint32_t)slabs * 2u);
          for (int s = 0; s < slabs; ++s) {
            const int c = pair_chan(nt * g.BN + s * 64, g.pg);
            tma_load_4d(pdst + (uint32_t)s * 2u * lo_off, &tm_r, bar_rfull + 8 * rb, c, w0 >> g.res_up2, h0 >> g.res_up2, i0);
            tma_load_4d(pdst + (uint32_t)s * 2u * lo_off + lo_off, &tm_r, bar_rfull + 8 * rb, c + g.pg, w0 >> g.res_up2,
                        h0 >> g.res_up2, i0);
          }
          continue;
        }
        // res_up2: the box of the half-resolution map that covers this tile is (bw/2, bh/2) at (w0/2, h0/2)
        mbar_arrive_expect_tx(bar_rfull + 8 * rb, (g.res_up2 ? box_bytes / 4 : box_bytes) * (uint32_t)slabs);
        for (int s = 0; s < slabs; ++s)
          tma_load_4d(rdst + s * TM_SLAB_BYTES, &tm_r, bar_rfull + 8 * rb, nt * g.BN + s * 64, w0 >> g.res_up2,
                      h0 >> g.res_up2, i0);
      }
    }
    __syncwarp();
  } else {
    // =============================== EPILOGUE (warps 0-7) ===============================
    const int q = warp & 3, half = warp >> 2;
    const int row = q * 32 + lane;
    const int units = g.BN / 32;                       // 32-column units of the accumulator
    const int upw = units >= 4 ? units / 2 : 1;        // units per half (BN = 64: one each)
    const int bar_id = g.BN == 64 ? 1 : 1 + half;
    const int bar_cnt = g.BN == 64 ? 256 : 128;
    const bool leader = (g.BN == 64 ? warp == 0 : q == 0) && lane == 0;
    const uint32_t out_slab = base + L.out + (g.BN == 64 ? 0u : (uint32_t)half * TM_SLAB_BYTES);
    const uint32_t sw_row = (uint32_t)row * 128u;
    const uint32_t rx = (uint32_t)(row & 7);
    int rrow = row;                                    // row of the residual slab this accumulator row reads
    if (g.res_up2) {
      const int w = row % g.bw, h = (row / g.bw) % g.bh, n = row / (g.bw * g.bh);
      rrow = (w >> 1) + (g.bw >> 1) * ((h >> 1) + (g.bh >> 1) * n);
    }
    const uint32_t rs_row = (uint32_t)rrow * 128u, rrx = (uint32_t)(rrow & 7);
    uint32_t ti_local = 0, oc = 0;
    for (long long tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++ti_local) {
      const int nt = (int)(tile % g.n_tiles);
      const long long mt = tile / g.n_tiles;
      const int w0 = (int)(mt % g.tiles_w) * g.bw;
      const int h0 = (int)((mt / g.tiles_w) % g.tiles_h) * g.bh;
      const int i0 = (int)(mt / ((long long)g.tiles_w * g.tiles_h)) * g.bn;
      const int n0 = nt * g.BN;
      const uint32_t buf = ti_local & 1u, use = ti_local >> 1;
      mbar_wait(bar_tfull + 8 * buf, use & 1u);
      tc_fence_after();
      if (g.has_res) mbar_wait(bar_rfull + 8 * buf, use & 1u);
      const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16) + buf * acc_cols;
      if (g.direct) {
        // thread = accumulator row = one output pixel of the box; the two halves split the columns
        const int wq = row % g.bw, hq = (row / g.bw) % g.bh, nq = row / (g.bw * g.bh);
        const int wo = w0 + wq, ho = h0 + hq, ni = i0 + nq;
        const bool ok = nq < g.bn && wo < g.Wo && ho < g.Ho && ni < g.N;
        const size_t HoWo = (size_t)g.Ho * g.Wo;
        const size_t pix = ((size_t)ni * g.Ho + ho) * g.Wo + wo;
        const int cbeg = half * (g.BN / 2), cend = cbeg + g.BN / 2;
        for (int cb = cbeg; cb < cend; cb += 16) {
          uint32_t v[16];
          if (g.wide) {       // hi*hi + lo*hi in column cb, hi*lo in column BN + cb
            uint32_t v2[16];
            tmem_ld16_issue(trow + (uint32_t)cb, v);
            tmem_ld16_issue(trow + (uint32_t)(g.BN + cb), v2);
            tmem_ld_wait();
#pragma unroll
            for (int e = 0; e < 16; ++e) v[e] = __float_as_uint(__uint_as_float(v[e]) + __uint_as_float(v2[e]));
          } else {
            tmem_ld16(trow + (uint32_t)cb, v);         // warp-collective
          }
          const int co0 = n0 + cb;
          if (!ok || co0 >= g.Cout) continue;
#pragma unroll
          for (int e = 0; e < 16; ++e) {
            const int co = co0 + e;
            if (co >= g.Cout) break;
            float o = __uint_as_float(v[e]);
            if (g.bias) o += __ldg(g.bias + co);
            if (g.relu) o = fmaxf(o, 0.f);
            if (g.sig_from >= 0 && co >= g.sig_from) o = 1.f / (1.f + expf(-o));
            const size_t oi = g.out_nhwc ? pix * g.Cout + co : ((size_t)ni * g.Cout + co) * HoWo + (size_t)ho * g.Wo + wo;
            if (g.y_bf16) reinterpret_cast<__nv_bfloat16*>(g.y)[oi] = __float2bfloat16_rn(o);
            else reinterpret_cast<float*>(g.y)[oi] = o;
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_tempty + 8 * buf);
        continue;
      }
      if (g.x3) {
        // ---- pair epilogue: all eight warps share one 64-channel (hi, lo) slab pair at a time; half h owns its columns
        //      [32h, 32h + 32).  fp32 result -> hi = bf16(o), lo = bf16(o - hi) -> two swizzled slabs -> two TMA stores ----
        const int nslab = g.BN / 64;
        const uint32_t jb = (uint32_t)half * 4u;
        const bool lead = warp == 0 && lane == 0;
        for (int sl = 0; sl < nslab; ++sl, ++oc) {
          const int u = 2 * sl + half;
          uint32_t v0[16], v1[16];
          tmem_ld16_issue(trow + (uint32_t)(u * 32), v0);
          tmem_ld16_issue(trow + (uint32_t)(u * 32 + 16), v1);
          uint32_t ob;
          if (g.res_inplace) {
            ob = base + L.out + (buf * (uint32_t)nslab + (uint32_t)sl) * 2u * TM_SLAB_BYTES;   // holds this slab's residual
          } else {
            ob = base + L.out + (g.opairs == 2 ? (oc & 1u) : 0u) * 2u * TM_SLAB_BYTES;
            if (lead) {                             // the stores that last used this pair have finished reading it
              if (g.opairs == 2) bulk_wait_read1(); else bulk_wait_read0();
            }
            named_bar_sync(1, 256);
          }
          tmem_ld_wait();
          float o[32];
#pragma unroll
          for (int e = 0; e < 16; ++e) { o[e] = __uint_as_float(v0[e]); o[16 + e] = __uint_as_float(v1[e]); }
          if (g.wide) {       // + the hi*lo column group
            tmem_ld16_issue(trow + (uint32_t)(g.BN + u * 32), v0);
            tmem_ld16_issue(trow + (uint32_t)(g.BN + u * 32 + 16), v1);
            tmem_ld_wait();
#pragma unroll
            for (int e = 0; e < 16; ++e) { o[e] += __uint_as_float(v0[e]); o[16 + e] += __uint_as_float(v1[e]); }
          }
          if (g.bias) {
            const float4* bp = reinterpret_cast<const float4*>(g.bias + n0 + u * 32);
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const float4 b4 = __ldg(bp + e);
              o[4 * e] += b4.x; o[4 * e + 1] += b4.y; o[4 * e + 2] += b4.z; o[4 * e + 3] += b4.w;
            }
          }
          if (g.has_res) {
            uint32_t rh, rl, xr;
            if (g.res_inplace) { rh = ob + sw_row; rl = rh + TM_SLAB_BYTES; xr = rx; }
            else { rh = base + L.res + (buf * (uint32_t)nslab + (uint32_t)sl) * 2u * L.res_slab + rs_row; rl = rh + L.res_slab; xr = rrx; }
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              const uint4 hv = lds128(rh + (((jb + c) ^ xr) << 4)), lv = lds128(rl + (((jb + c) ^ xr) << 4));
              const uint32_t hw[4] = {hv.x, hv.y, hv.z, hv.w}, lw[4] = {lv.x, lv.y, lv.z, lv.w};
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                o[c * 8 + 2 * e] += __uint_as_float(hw[e] << 16) + __uint_as_float(lw[e] << 16);
                o[c * 8 + 2 * e + 1] += __uint_as_float(hw[e] & 0xffff0000u) + __uint_as_float(lw[e] & 0xffff0000u);
              }
            }
          }
          if (g.relu) {
#pragma unroll
            for (int e = 0; e < 32; ++e) o[e] = fmaxf(o[e], 0.f);
          }
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            uint32_t hw[4], lw[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float a = o[c * 8 + 2 * e], b = o[c * 8 + 2 * e + 1];
              hw[e] = pack_bf16x2(a, b);
              lw[e] = pack_bf16x2(a - __uint_as_float(hw[e] << 16), b - __uint_as_float(hw[e] & 0xffff0000u));
            }
            sts128(ob + sw_row + (((jb + c) ^ rx) << 4), make_uint4(hw[0], hw[1], hw[2], hw[3]));
            sts128(ob + TM_SLAB_BYTES + sw_row + (((jb + c) ^ rx) << 4), make_uint4(lw[0], lw[1], lw[2], lw[3]));
          }
          fence_proxy_async();
          named_bar_sync(1, 256);
          if (lead) {
            const int c = pair_chan(n0 + sl * 64, g.pg);
            tma_store_4d(&tm_y, ob, c, w0, h0, i0);
            tma_store_4d(&tm_y, ob + TM_SLAB_BYTES, c + g.pg, w0, h0, i0);
            bulk_commit();
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(bar_tempty + 8 * buf);
          if (g.has_res && !g.res_inplace) mbar_arrive(bar_rempty + 8 * buf);
        }
        if (g.res_inplace && lead) {     // the slab pairs of this tile may be refilled once their stores have been read out
          bulk_wait_read0();
          mbar_arrive(bar_rempty + 8 * buf);
        }
        continue;
      }
      for (int ui = 0; ui < upw; ++ui) {
        const int u = half * upw + ui;
        const int slab = u >> 1;
        const uint32_t jb = (uint32_t)(u & 1) * 4u;     // first 16-byte chunk of this unit inside the slab row
        uint32_t v0[16], v1[16];
        tmem_ld16_issue(trow + (uint32_t)(u * 32), v0);
        tmem_ld16_issue(trow + (uint32_t)(u * 32 + 16), v1);
        if ((u & 1) == 0 || g.BN == 64) {
          // the output slab is about to be overwritten: its previous TMA store must have finished reading it
          if (leader) bulk_wait_read0();
          named_bar_sync(bar_id, bar_cnt);
        }
        tmem_ld_wait();
        float o[32];
#pragma unroll
        for (int e = 0; e < 16; ++e) { o[e] = __uint_as_float(v0[e]); o[16 + e] = __uint_as_float(v1[e]); }
        if (g.bias) {
          const float4* bp = reinterpret_cast<const float4*>(g.bias + n0 + u * 32);
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const float4 b4 = __ldg(bp + e);
            o[4 * e] += b4.x; o[4 * e + 1] += b4.y; o[4 * e + 2] += b4.z; o[4 * e + 3] += b4.w;
          }
        }
        if (g.has_res) {
          const uint32_t rs = base + L.res + (buf * (uint32_t)(g.BN / 64) + (uint32_t)slab) * TM_SLAB_BYTES + rs_row;
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const uint4 rv = lds128(rs + (((jb + c) ^ rrx) << 4));
            const uint32_t rw[4] = {rv.x, rv.y, rv.z, rv.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              o[c * 8 + 2 * e] += __uint_as_float(rw[e] << 16);
              o[c * 8 + 2 * e + 1] += __uint_as_float(rw[e] & 0xffff0000u);
            }
          }
        }
        if (g.relu) {
#pragma unroll
          for (int e = 0; e < 32; ++e) o[e] = fmaxf(o[e], 0.f);
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint4 w;
          w.x = pack_bf16x2(o[c * 8], o[c * 8 + 1]); w.y = pack_bf16x2(o[c * 8 + 2], o[c * 8 + 3]);
          w.z = pack_bf16x2(o[c * 8 + 4], o[c * 8 + 5]); w.w = pack_bf16x2(o[c * 8 + 6], o[c * 8 + 7]);
          sts128(out_slab + sw_row + (((jb + c) ^ rx) << 4), w);
        }
        if ((u & 1) == 1 || g.BN == 64) {
          fence_proxy_async();                 // generic-proxy smem writes -> visible to the TMA store
          named_bar_sync(bar_id, bar_cnt);
          if (leader) {
            tma_store_4d(&tm_y, out_slab, n0 + slab * 64, w0, h0, i0);
            bulk_commit();
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(bar_tempty + 8 * buf);
        if (g.has_res) mbar_arrive(bar_rempty + 8 * buf);
      }
    }
    if (leader || (g.x3 && warp == 0 && lane == 0)) bulk_wait0();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == TM_WARP_MMA) {
    tc_fence_after();
    tmem_dealloc(tmem_base, tmem_cols);
  }
}

// ----------------------------------------------------------------------------------------------
// 2-CTA variant (cta_group::2) for the latency-bound layers of the PAIR stream.
//
// A cluster of two CTAs (one TPC) works on an M = 256 tile: CTA r owns output rows [128 r, 128 r + 128) -- m-tile 2p + r of
// the same n-tile -- and stages its OWN A tiles (hi, lo) but only HALF of the weight tile (rows [n0 + r*BN/2, +BN/2) of both
// planes); `tcgen05.mma.cta_group::2` (issued by the leader CTA's single MMA thread, M = 256) reads the B halves from both
// CTAs' shared memory.  A pair k-block is then 48 KB per CTA instead of 64 KB, so FOUR ring stages fit next to the output
// slab pair instead of three -- the ring is latency-bound (profiles/r2_mma_probe_pair.md), stages are throughput.
// Protocol (CUTLASS sm100 2-SM scheme): every CTA's TMA thread issues its loads with `.cta_group::2`, their complete_tx
// bytes go to the LEADER's full barrier (which expects the bytes of both CTAs); `tcgen05.commit ... multicast::cluster`
// releases the stage / publishes the accumulator in BOTH CTAs; both CTAs' epilogue warps arrive on the leader's
// tmem-empty barrier (remote mbarrier arrive for CTA 1).  Per-tap boxes only, no residual (the +res layers keep the 1-CTA
// kernel with in-place slab pairs); slab (pair) or direct (fp32) epilogue as in the 1-CTA kernel.
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa_rank(uint32_t addr, uint32_t rank) {   // shared::cta address -> shared::cluster address in CTA `rank`
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void tma2_load_4d(uint32_t dst, const CUtensorMap* tm, uint32_t bar_leader, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(bar_leader), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma2_load_2d(uint32_t dst, const CUtensorMap* tm, uint32_t bar_leader, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(bar_leader), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma2_bf16_lohi(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t desc_hi, uint32_t idesc,
                                                uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "mov.b64 da, {%1, %5};\n\t"
      "mov.b64 db, {%2, %5};\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(accumulate), "r"(desc_hi) : "memory");
}
__device__ __forceinline__ void umma2_commit_mc(uint32_t bar) {      // arrive on `bar` (same offset) in both CTAs of the pair
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"((uint16_t)3) : "memory");
}

struct Tma2Smem { uint32_t stages, out, res, res_slab, a_half, b_half, a_bytes, b_bytes, stage_bytes, total; };
__host__ __device__ inline Tma2Smem tma2_smem_layout(int BN, int stages, bool direct, int opairs, bool has_res = false,
                                                    bool res_up2 = false, bool res_inplace = false, bool wide = false) {
  Tma2Smem s;
  s.a_half = 128 * 128; s.b_half = (uint32_t)(BN / 2) * 128;      // this CTA's half of the weight tile, per plane
  // wide: [Y: BN rows = this CTA's half of [W_hi ; W_lo]] [X: BN/2 rows = this CTA's half of W_hi]
  s.a_bytes = 2 * s.a_half; s.b_bytes = (wide ? 3u : 2u) * s.b_half;
  s.stage_bytes = s.a_bytes + s.b_bytes;
  s.stages = 1024;
  s.out = s.stages + s.stage_bytes * (uint32_t)stages;
  s.res_slab = res_up2 ? 4096u : (uint32_t)TM_SLAB_BYTES;
  const uint32_t pairs = direct ? 0u : (res_inplace ? 2u * (uint32_t)(BN / 64) : (uint32_t)opairs);
  s.res = s.out + pairs * 2u * TM_SLAB_BYTES;
  s.total = s.res + ((has_res && !res_inplace) ? 2u * (uint32_t)(BN / 64) * 2u * s.res_slab : 0u);
  return s;
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(TM_THREADS, 1)
igemm_tma2_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_w,
                  const __grid_constant__ CUtensorMap tm_y, const __grid_constant__ CUtensorMap tm_r, const TmaGeom g) {
  extern __shared__ __align__(1024) uint8_t smem_dyn[];
  const uint32_t raw = smem_u32(smem_dyn);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* sm = smem_dyn + (base - raw);
  const Tma2Smem L = tma2_smem_layout(g.BN, g.stages, g.direct != 0, g.opairs, g.has_res != 0, g.res_up2 != 0, g.res_inplace != 0,
                                      g.wide != 0);
  const uint32_t acc_cols = (uint32_t)(g.wide ? 2 * g.BN : g.BN);      // TMEM columns of one accumulator buffer
  const uint32_t bar_full = base, bar_empty = base + 8 * TM_MAX_STAGES;
  const uint32_t bar_tfull = bar_empty + 8 * TM_MAX_STAGES, bar_tempty = bar_tfull + 16;
  const uint32_t bar_rfull = bar_tempty + 16, bar_rempty = bar_rfull + 16;     // residual slabs: CTA-local protocol
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(sm + 8 * (2 * TM_MAX_STAGES + 16));

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader_cta = rank == 0;
  const int cchunks = g.Cin / 64;
  const int num_kb = g.KHW * cchunks;
  const long long m_tiles = (long long)g.tiles_w * g.tiles_h * g.tiles_n;
  const long long p_tiles = (m_tiles + 1) / 2;                      // pairs of m-tiles
  const long long num_tiles = p_tiles * g.n_tiles;
  const long long cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;
  const uint32_t box_bytes = (uint32_t)(g.bw * g.bh * g.bn) * 128u;
  uint32_t tmem_cols = 32;
  while (tmem_cols < 2 * acc_cols) tmem_cols <<= 1;

  if (warp == TM_WARP_MMA) {
    if (lane == 0) {
      for (int s = 0; s < g.stages; ++s) {
        mbar_init(bar_full + 8 * s, 1);
        mbar_init(bar_empty + 8 * s, 1);
      }
      for (int b = 0; b < 2; ++b) {
        mbar_init(bar_tfull + 8 * b, 1);
        mbar_init(bar_tempty + 8 * b, 2 * TM_EPI_WARPS);      // the epilogue warps of BOTH CTAs (used in the leader only)
        mbar_init(bar_rfull + 8 * b, 1);
        mbar_init(bar_rempty + 8 * b, g.res_inplace ? 1 : TM_EPI_WARPS);
      }
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc2(smem_u32(tmem_ptr_smem), tmem_cols);
  } else if (warp == TM_WARP_TMA && lane == 0) {
    prefetch_tmap(&tm_x);
    prefetch_tmap(&tm_w);
    prefetch_tmap(&tm_y);
    if (g.has_res) prefetch_tmap(&tm_r);
  }
  tc_fence_before();
  cluster_sync_all();          // barriers of both CTAs initialised, TMEM allocated in both
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == TM_WARP_TMA) {
    if (lane == 0) {
      uint32_t s = 0, ph = 0;
      uint32_t a_dst = base + L.stages;
      const uint32_t tx_both = 2u * (2u * box_bytes + L.b_bytes);       // both CTAs' A (hi, lo) boxes + B parts
      for (long long tile = cluster_id; tile < num_tiles; tile += num_clusters) {
        const int nt = (int)(tile % g.n_tiles);
        const long long mt = 2 * (tile / g.n_tiles) + rank;
        const int w0 = (int)(mt % g.tiles_w) * g.bw;
        const int h0 = (int)((mt / g.tiles_w) % g.tiles_h) * g.bh;
        const int i0 = (int)(mt / ((long long)g.tiles_w * g.tiles_h)) * g.bn;     // >= N for the odd tail tile: all-zero boxes
        const int nrow = nt * g.BN + (int)rank * (g.BN / 2);                      // this CTA's half of the weight rows
        int cc = 0, kj = 0, cw = w0 - g.pw, ch = h0 - g.ph;
        for (int kb = 0; kb < num_kb; ++kb) {
          const uint32_t bf_local = bar_full + 8 * s;
          const uint32_t bf = mapa_rank(bf_local, 0);                 // complete_tx goes to the leader CTA's barrier
          mbar_wait(bar_empty + 8 * s, ph ^ 1u);
          if (leader_cta) mbar_arrive_expect_tx(bf_local, tx_both);
          tma2_load_4d(a_dst, &tm_x, bf, cc * 64, cw, ch, i0);
          tma2_load_4d(a_dst + L.a_half, &tm_x, bf, g.x_lo + cc * 64, cw, ch, i0);
          if (g.wide) {
            const int yrow = (rank ? g.w_lo : 0) + nt * g.BN;              // CTA 0: the hi rows of the n-tile, CTA 1: its lo rows
            tma2_load_2d(a_dst + L.a_bytes, &tm_w, bf, kb * 64, yrow);
            tma2_load_2d(a_dst + L.a_bytes + L.b_half, &tm_w, bf, kb * 64, yrow + g.BN / 2);
            tma2_load_2d(a_dst + L.a_bytes + 2 * L.b_half, &tm_w, bf, kb * 64, nrow);     // X: this CTA's half of W_hi
          } else {
            tma2_load_2d(a_dst + L.a_bytes, &tm_w, bf, kb * 64, nrow);
            tma2_load_2d(a_dst + L.a_bytes + L.b_half, &tm_w, bf, kb * 64, g.w_lo + nrow);
          }
          if (++cc == cchunks) {
            cc = 0; cw += g.dw;
            if (++kj == g.kw) { kj = 0; cw = w0 - g.pw; ch += g.dh; }
          }
          a_dst += L.stage_bytes;
          if (++s == (uint32_t)g.stages) { s = 0; ph ^= 1u; a_dst = base + L.stages; }
        }
      }
    }
    __syncwarp();
  } else if (warp == TM_WARP_MMA) {
    if (lane == 0 && leader_cta) {
      const uint32_t idesc = umma_idesc(256, g.BN), idesc_w = umma_idesc(256, 2 * g.BN);
      const uint32_t desc_hi = (uint32_t)(1024 >> 4) | (1u << 14) | (2u << 29);
      const uint32_t a_lo0 = ((base + L.stages) >> 4) & 0x3fffu, stage16 = L.stage_bytes >> 4, a16 = L.a_bytes >> 4;
      const uint32_t ah16 = L.a_half >> 4, bh16 = L.b_half >> 4;
      uint32_t s = 0, ph = 0, a_lo = a_lo0, ti_local = 0;
      for (long long tile = cluster_id; tile < num_tiles; tile += num_clusters, ++ti_local) {
        const uint32_t buf = ti_local & 1u, use = ti_local >> 1;
        mbar_wait(bar_tempty + 8 * buf, (use & 1u) ^ 1u);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + buf * acc_cols;
        uint32_t acc = 0u;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(bar_full + 8 * s, ph);
          tc_fence_after();
          const uint32_t b_lo = a_lo + a16;
#pragma unroll
          for (uint32_t k = 0; k < 4; ++k) {
            if (g.wide) {
              umma2_bf16_lohi(tmem_d, a_lo + 2 * k, b_lo + 2 * k, desc_hi, idesc_w, acc);                 // hi * [hi ; lo] -> columns [0, 2 BN)
              umma2_bf16_lohi(tmem_d, a_lo + ah16 + 2 * k, b_lo + 2 * bh16 + 2 * k, desc_hi, idesc, 1u);  // lo * hi      -> columns [0, BN)
            } else {
              umma2_bf16_lohi(tmem_d, a_lo + ah16 + 2 * k, b_lo + 2 * k, desc_hi, idesc, acc);
              umma2_bf16_lohi(tmem_d, a_lo + 2 * k, b_lo + bh16 + 2 * k, desc_hi, idesc, 1u);
              umma2_bf16_lohi(tmem_d, a_lo + 2 * k, b_lo + 2 * k, desc_hi, idesc, 1u);
            }
            acc = 1u;
          }
          umma2_commit_mc(bar_empty + 8 * s);
          a_lo += stage16;
          if (++s == (uint32_t)g.stages) { s = 0; ph ^= 1u; a_lo = a_lo0; }
        }
        umma2_commit_mc(bar_tfull + 8 * buf);
      }
    }
    __syncwarp();
  } else if (warp == TM_WARP_RES) {
    // residual slab pairs of this CTA's own m-tile: a CTA-local producer / consumer pair, exactly as in the 1-CTA kernel
    if (lane == 0 && g.has_res) {
      uint32_t ti_local = 0;
      const int slabs = g.BN / 64;
      for (long long tile = cluster_id; tile < num_tiles; tile += num_clusters, ++ti_local) {
        const int nt = (int)(tile % g.n_tiles);
        const long long mt = 2 * (tile / g.n_tiles) + rank;
        const int w0 = (int)(mt % g.tiles_w) * g.bw;
        const int h0 = (int)((mt / g.tiles_w) % g.tiles_h) * g.bh;
        const int i0 = (int)(mt / ((long long)g.tiles_w * g.tiles_h)) * g.bn;
        const uint32_t rb = ti_local & 1u, ruse = ti_local >> 1;
        mbar_wait(bar_rempty + 8 * rb, (ruse & 1u) ^ 1u);
        const uint32_t lo_off = g.res_inplace ? (uint32_t)TM_SLAB_BYTES : L.res_slab;
        const uint32_t pdst = g.res_inplace ? base + L.out + rb * (uint32_t)slabs * 2u * TM_SLAB_BYTES
                                            : base + L.res + rb * (uint32_t)slabs * 2u * L.res_slab;
        mbar_arrive_expect_tx(bar_rfull + 8 * rb, (g.res_up2 ? box_bytes / 4 : box_bytes) * (uint32_t)slabs * 2u);
        for (int sl = 0; sl < slabs; ++sl) {
          const int c = pair_chan(nt * g.BN + sl * 64, g.pg);
          tma_load_4d(pdst + (uint32_t)sl * 2u * lo_off, &tm_r, bar_rfull + 8 * rb, c, w0 >> g.res_up2, h0 >> g.res_up2, i0);
          tma_load_4d(pdst + (uint32_t)sl * 2u * lo_off + lo_off, &tm_r, bar_rfull + 8 * rb, c + g.pg, w0 >> g.res_up2,
                      h0 >> g.res_up2, i0);
        }
      }
    }
    __syncwarp();
  } else if (warp < TM_EPI_WARPS) {
    const int q = warp & 3, half = warp >> 2;
    const int row = q * 32 + lane;
    const uint32_t sw_row = (uint32_t)row * 128u;
    const uint32_t rx = (uint32_t)(row & 7);
    int rrow = row;
    if (g.res_up2) {
      const int w = row % g.bw, h = (row / g.bw) % g.bh, n = row / (g.bw * g.bh);
      rrow = (w >> 1) + (g.bw >> 1) * ((h >> 1) + (g.bh >> 1) * n);
    }
    const uint32_t rs_row = (uint32_t)rrow * 128u, rrx = (uint32_t)(rrow & 7);
    const uint32_t tempty_leader = mapa_rank(bar_tempty, 0);
    uint32_t ti_local = 0, oc = 0;
    for (long long tile = cluster_id; tile < num_tiles; tile += num_clusters, ++ti_local) {
      const int nt = (int)(tile % g.n_tiles);
      const long long mt = 2 * (tile / g.n_tiles) + rank;
      const int w0 = (int)(mt % g.tiles_w) * g.bw;
      const int h0 = (int)((mt / g.tiles_w) % g.tiles_h) * g.bh;
      const int i0 = (int)(mt / ((long long)g.tiles_w * g.tiles_h)) * g.bn;
      const int n0 = nt * g.BN;
      const uint32_t buf = ti_local & 1u, use = ti_local >> 1;
      mbar_wait(bar_tfull + 8 * buf, use & 1u);
      tc_fence_after();
      if (g.has_res) mbar_wait(bar_rfull + 8 * buf, use & 1u);
      const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16) + buf * acc_cols;
      if (g.direct) {
        const int wq = row % g.bw, hq = (row / g.bw) % g.bh, nq = row / (g.bw * g.bh);
        const int wo = w0 + wq, ho = h0 + hq, ni = i0 + nq;
        const bool ok = nq < g.bn && wo < g.Wo && ho < g.Ho && ni < g.N;
        const size_t HoWo = (size_t)g.Ho * g.Wo;
        const size_t pix = ((size_t)ni * g.Ho + ho) * g.Wo + wo;
        const int cbeg = half * (g.BN / 2), cend = cbeg + g.BN / 2;
        for (int cb = cbeg; cb < cend; cb += 16) {
          uint32_t v[16];
          if (g.wide) {       // hi*hi + lo*hi in column cb, hi*lo in column BN + cb
            uint32_t v2[16];
            tmem_ld16_issue(trow + (uint32_t)cb, v);
            tmem_ld16_issue(trow + (uint32_t)(g.BN + cb), v2);
            tmem_ld_wait();
#pragma unroll
            for (int e = 0; e < 16; ++e) v[e] = __float_as_uint(__uint_as_float(v[e]) + __uint_as_float(v2[e]));
          } else {
            tmem_ld16(trow + (uint32_t)cb, v);
          }
          const int co0 = n0 + cb;
          if (!ok || co0 >= g.Cout) continue;
#pragma unroll
          for (int e = 0; e < 16; ++e) {
            const int co = co0 + e;
            if (co >= g.Cout) break;
            float o = __uint_as_float(v[e]);
            if (g.bias) o += __ldg(g.bias + co);
            if (g.relu) o = fmaxf(o, 0.f);
            if (g.sig_from >= 0 && co >= g.sig_from) o = 1.f / (1.f + expf(-o));
            const size_t oi = g.out_nhwc ? pix * g.Cout + co : ((size_t)ni * g.Cout + co) * HoWo + (size_t)ho * g.Wo + wo;
            reinterpret_cast<float*>(g.y)[oi] = o;
          }
        }
      } else {
        const int nslab = g.BN / 64;
        const uint32_t jb = (uint32_t)half * 4u;
        const bool lead = warp == 0 && lane == 0;
        for (int sl = 0; sl < nslab; ++sl, ++oc) {
          const int u = 2 * sl + half;
          uint32_t v0[16], v1[16];
          tmem_ld16_issue(trow + (uint32_t)(u * 32), v0);
          tmem_ld16_issue(trow + (uint32_t)(u * 32 + 16), v1);
          uint32_t ob;
          if (g.res_inplace) {
            ob = base + L.out + (buf * (uint32_t)nslab + (uint32_t)sl) * 2u * TM_SLAB_BYTES;   // holds this slab's residual
          } else {
            ob = base + L.out + (g.opairs == 2 ? (oc & 1u) : 0u) * 2u * TM_SLAB_BYTES;
            if (lead) {
              if (g.opairs == 2) bulk_wait_read1(); else bulk_wait_read0();
            }
            named_bar_sync(1, 256);
          }
          tmem_ld_wait();
          float o[32];
#pragma unroll
          for (int e = 0; e < 16; ++e) { o[e] = __uint_as_float(v0[e]); o[16 + e] = __uint_as_float(v1[e]); }
          if (g.wide) {       // + the hi*lo column group
            tmem_ld16_issue(trow + (uint32_t)(g.BN + u * 32), v0);
            tmem_ld16_issue(trow + (uint32_t)(g.BN + u * 32 + 16), v1);
            tmem_ld_wait();
#pragma unroll
            for (int e = 0; e < 16; ++e) { o[e] += __uint_as_float(v0[e]); o[16 + e] += __uint_as_float(v1[e]); }
          }
          if (g.bias) {
            const float4* bp = reinterpret_cast<const float4*>(g.bias + n0 + u * 32);
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const float4 b4 = __ldg(bp + e);
              o[4 * e] += b4.x; o[4 * e + 1] += b4.y; o[4 * e + 2] += b4.z; o[4 * e + 3] += b4.w;
            }
          }
          if (g.has_res) {
            uint32_t rh, rl, xr;
            if (g.res_inplace) { rh = ob + sw_row; rl = rh + TM_SLAB_BYTES; xr = rx; }
            else { rh = base + L.res + (buf * (uint32_t)nslab + (uint32_t)sl) * 2u * L.res_slab + rs_row; rl = rh + L.res_slab; xr = rrx; }
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              const uint4 hv = lds128(rh + (((jb + c) ^ xr) << 4)), lv = lds128(rl + (((jb + c) ^ xr) << 4));
              const uint32_t hw[4] = {hv.x, hv.y, hv.z, hv.w}, lw[4] = {lv.x, lv.y, lv.z, lv.w};
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                o[c * 8 + 2 * e] += __uint_as_float(hw[e] << 16) + __uint_as_float(lw[e] << 16);
                o[c * 8 + 2 * e + 1] += __uint_as_float(hw[e] & 0xffff0000u) + __uint_as_float(lw[e] & 0xffff0000u);
              }
            }
          }
          if (g.relu) {
#pragma unroll
            for (int e = 0; e < 32; ++e) o[e] = fmaxf(o[e], 0.f);
          }
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            uint32_t hw[4], lw[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float a = o[c * 8 + 2 * e], b = o[c * 8 + 2 * e + 1];
              hw[e] = pack_bf16x2(a, b);
              lw[e] = pack_bf16x2(a - __uint_as_float(hw[e] << 16), b - __uint_as_float(hw[e] & 0xffff0000u));
            }
            sts128(ob + sw_row + (((jb + c) ^ rx) << 4), make_uint4(hw[0], hw[1], hw[2], hw[3]));
            sts128(ob + TM_SLAB_BYTES + sw_row + (((jb + c) ^ rx) << 4), make_uint4(lw[0], lw[1], lw[2], lw[3]));
          }
          fence_proxy_async();
          named_bar_sync(1, 256);
          if (lead) {
            const int c = pair_chan(n0 + sl * 64, g.pg);
            tma_store_4d(&tm_y, ob, c, w0, h0, i0);
            tma_store_4d(&tm_y, ob + TM_SLAB_BYTES, c + g.pg, w0, h0, i0);
            bulk_commit();
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive_cluster(tempty_leader + 8 * buf);     // leader's barrier: 16 arrivals = both CTAs drained
        if (g.has_res && !g.res_inplace) mbar_arrive(bar_rempty + 8 * buf);
      }
      if (g.res_inplace && warp == 0 && lane == 0) {      // slab pairs may be refilled once their stores have been read out
        bulk_wait_read0();
        mbar_arrive(bar_rempty + 8 * buf);
      }
    }
    if (warp == 0 && lane == 0) bulk_wait0();
  }
  tc_fence_before();
  cluster_sync_all();          // no CTA may exit (or free TMEM) while its peer still reads its shared memory / signals its barriers
  if (warp == TM_WARP_MMA) {
    tc_fence_after();
    tmem_dealloc2(tmem_base, tmem_cols);
  }
}

// ----------------------------------------------------------------------------------------------
// host side: tensor maps, tile geometry, launch
// ----------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn tma_encoder() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qr = cudaDriverEntryPointSymbolNotFound;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr) == cudaSuccess &&
        qr == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
    else
      (void)cudaGetLastError();
    tried = true;
  }
  return fn;
}

// bf16 tensor, innermost dimension first; box[0] = 64 elements = one 128-byte swizzle span
static bool encode_bf16(EncodeTiledFn enc, CUtensorMap* tm, const void* ptr, int rank, const cuuint64_t* dims,
                        const cuuint32_t* box, const cuuint64_t* byte_strides = nullptr) {
  cuuint64_t strides[5];
  cuuint64_t acc = 2;
  for (int i = 0; i + 1 < rank; ++i) { acc *= dims[i]; strides[i] = byte_strides ? byte_strides[i] : acc; }
  const cuuint32_t es[5] = {1, 1, 1, 1, 1};
  return enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(ptr), dims, strides, box, es,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// Output-pixel box (bw, bh, bn) with bw*bh*bn <= 128: fewest tiles, then smallest input halo, then widest rows.
static void tma_pick_box(int N, int Ho, int Wo, int kh, int kw, int dh, int dw, bool even, int* bw_o, int* bh_o, int* bn_o) {
  long long best_tiles = -1, best_halo = 0;
  int bbw = 1, bbh = 1, bbn = 1;
  const int wmax = Wo < 128 ? Wo : 128;
  for (int bw = 1; bw <= wmax; ++bw) {
    if (Wo > 32 && (bw & (bw - 1))) continue;     // large maps: power-of-two widths only (keeps the search tiny)
    if (even && (bw & 1)) continue;
    const int hmax = (128 / bw) < Ho ? (128 / bw) : Ho;
    for (int bh = 1; bh <= hmax; ++bh) {
      if (even && (bh & 1)) continue;
      int bn = 128 / (bw * bh);
      if (bn > N) bn = N;
      if (bn < 1) continue;
      const long long tiles = (long long)((Wo + bw - 1) / bw) * ((Ho + bh - 1) / bh) * ((N + bn - 1) / bn);
      const long long halo = (long long)(bw + (kw - 1) * dw) * (bh + (kh - 1) * dh) * bn;
      if (best_tiles < 0 || tiles < best_tiles || (tiles == best_tiles && (halo < best_halo || (halo == best_halo && bw > bbw)))) {
        best_tiles = tiles; best_halo = halo; bbw = bw; bbh = bh; bbn = bn;
      }
    }
  }
  *bw_o = bbw; *bh_o = bbh; *bn_o = bbn;
}

int launch_igemm_tma(const TcParams& p, const void* packed, cudaStream_t stream) {
  if (p.no_tma || p.offset) return UPSNET_E_UNSUPPORTED;
  // precision bf16 runs on bf16 activations, precision bf16x3 on hi/lo bf16 pairs (same tile pipeline, three MMAs)
  const bool pair = p.x_pair != 0;
  if (pair != (p.x3 != 0) || (!pair && !p.x_bf16)) return UPSNET_E_UNSUPPORTED;
  if (p.y_pair && (!pair || !p.out_nhwc || (p.Cout % 64) != 0)) return UPSNET_E_UNSUPPORTED;
  if (pair && p.y_bf16) return UPSNET_E_UNSUPPORTED;     // pair in -> pair (slab epilogue) or fp32 (direct epilogue) out
  const int pg = p.y_pair ? (p.pair_group > 0 ? p.pair_group : p.Cout) : 0;
  if (p.y_pair && ((pg % 64) != 0 || (p.Cout % pg) != 0)) return UPSNET_E_UNSUPPORTED;
  // slab epilogue (TMA store): bf16 / pair NHWC output with Cout % 64 == 0; everything else without a residual goes
  // through the direct-store epilogue (small heads: Cout 9..45, fp32 planes)
  const bool direct = (!p.y_bf16 && !p.y_pair) || !p.out_nhwc || (p.Cout % 64) != 0;
  if (p.sig_from >= 0 && !direct) return UPSNET_E_UNSUPPORTED;
  if (direct && (p.residual || p.Cout > 256)) return UPSNET_E_UNSUPPORTED;
  if (p.res_up2 && (!p.residual || (p.Ho & 1) || (p.Wo & 1))) return UPSNET_E_UNSUPPORTED;
  // stride > 1 only for 1x1 / pad 0 (the ResNet down-sampling convs): the input is then addressed through a
  // strided VIEW (every sh-th row, sw-th pixel) and the layer is a stride-1 1x1 convolution of that view
  const bool strided = p.sh != 1 || p.sw != 1;
  if (strided && (p.kh != 1 || p.kw != 1 || p.ph != 0 || p.pw != 0)) return UPSNET_E_UNSUPPORTED;
  if ((p.Cin % 64) || (!direct && (p.Cout % 64)) || p.kh * p.kw > 49) return UPSNET_E_UNSUPPORTED;
  if ((((uintptr_t)p.x) & 15) || (((uintptr_t)p.y) & 15) || (((uintptr_t)packed) & 15) || (p.residual && (((uintptr_t)p.residual) & 15)))
    return UPSNET_E_UNSUPPORTED;
  if (p.bias && (((uintptr_t)p.bias) & 15)) return UPSNET_E_UNSUPPORTED;
  EncodeTiledFn enc = tma_encoder();
  if (!enc) return UPSNET_E_UNSUPPORTED;
  static int sms = 0;
  if (sms == 0) {
    int dev = 0, v = kNumSMs;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
    sms = v > 0 ? v : kNumSMs;
  }
  TmaGeom g{};
  g.bias = p.bias;
  g.N = p.N; g.Ho = p.Ho; g.Wo = p.Wo; g.Cout = p.Cout; g.Cin = p.Cin;
  g.kw = p.kw; g.KHW = p.kh * p.kw; g.ph = p.ph; g.pw = p.pw; g.dh = p.dh; g.dw = p.dw;
  g.relu = p.relu; g.has_res = p.residual ? 1 : 0; g.res_up2 = p.res_up2 ? 1 : 0; g.sig_from = p.sig_from;
  tma_pick_box(p.N, p.Ho, p.Wo, p.kh, p.kw, p.dh, p.dw, g.res_up2 != 0, &g.bw, &g.bh, &g.bn);
  // halo mode for k x k filters (UPSNET_TMA_HALO=0 disables it, =2 also enables it for BN = 256): 8 x 16-pixel tiles,
  // 16-pixel-wide patch rows; the patch must cover 8 + (kw-1)*dw <= 16 pixels per row
  static int halo_env = -1;
  if (halo_env < 0) { const char* e = getenv("UPSNET_TMA_HALO"); halo_env = e ? atoi(e) : 1; }
  const int patch_h = 16 + (p.kh - 1) * p.dh;
  bool halo = halo_env > 0 && !strided && p.kh * p.kw > 1 && (p.kw - 1) * p.dw <= 8 && patch_h <= 48 && !g.res_up2 &&
              p.Wo >= 8 && p.Ho >= 8;
  // pair mode (measured, profiles/r2_mma_probe_pair.md): the patch slots of a hi/lo pair (2 x 36 KB per 3x3 slot) only fit next
  // to N tiles <= 64, and there the tap-shifted (not 1024-byte aligned) A descriptors make the short N <= 64 MMAs ~2x slower
  // than the per-tap boxes -- halo mode stays a bf16-stream optimisation (UPSNET_TMA_HALO=3 forces it for pairs)
  if (pair && halo_env < 3) halo = false;
  if (halo) { g.bw = 8; g.bh = 16; g.bn = 1; }
  g.tiles_w = (p.Wo + g.bw - 1) / g.bw;
  g.tiles_h = (p.Ho + g.bh - 1) / g.bh;
  g.tiles_n = (p.N + g.bn - 1) / g.bn;
  long long m_tiles = (long long)g.tiles_w * g.tiles_h * g.tiles_n;
  const int Cout_pad = p.Cout <= 32 ? 32 : (p.Cout + 63) / 64 * 64;      // rows of the packed weight planes (tc_cout_pad)
  int BN = (Cout_pad % 256 == 0 && !g.has_res) ? 256 : ((Cout_pad % 128 == 0) ? 128 : (Cout_pad % 64 == 0 ? 64 : 32));
  // pair mode: operand tiles are twice as large -- N tile <= 128, and 64 next to residual slab pairs (shared memory)
  g.x3 = pair ? 1 : 0; g.x_lo = p.Cin; g.w_lo = Cout_pad; g.pg = pg;
  g.res_inplace = (pair && g.has_res && !g.res_up2) ? 1 : 0;
  g.opairs = 2;
  if (pair && BN > 128) BN = 128;
  if (pair && g.has_res && BN > 64) BN = 64;
  // N tile: as wide as possible (operand bytes per flop fall with BN) while ~2/3 of the SMs still get a tile; measured
  // on B200 (profiles/r1_bn_sweep.md): 64 m-tiles x Cout 256 -> BN 128 (128 CTAs) beats BN 64 (256 tiles) by 38 %
  // and BN 256 (64 CTAs) by 11 %; 16 m-tiles x Cout 512 -> BN 64 (128 CTAs) stays best.
  while (BN > 64 && m_tiles * (Cout_pad / BN) < 96) BN /= 2;
  (void)sms;
  if (const char* fb = getenv("UPSNET_TMA_FORCE_BN")) {     // tuning hook (scripts/bn_sweep.py): force the N tile
    const int v = atoi(fb);
    if ((v == 64 || v == 128 || v == 256) && Cout_pad % v == 0 && !(g.has_res && v > 128) && !(pair && (v > 128 || (g.has_res && v > 64)))) BN = v;
  }
  if (halo && BN == 256 && halo_env < 2) {     // wide-N layers are MMA-bound: keep the fewest-tiles box for them
    halo = false;
    tma_pick_box(p.N, p.Ho, p.Wo, p.kh, p.kw, p.dh, p.dw, false, &g.bw, &g.bh, &g.bn);
    g.tiles_w = (p.Wo + g.bw - 1) / g.bw;
    g.tiles_h = (p.Ho + g.bh - 1) / g.bh;
    g.tiles_n = (p.N + g.bn - 1) / g.bn;
    m_tiles = (long long)g.tiles_w * g.tiles_h * g.tiles_n;
  }
  g.BN = BN;
  g.n_tiles = Cout_pad / BN;
  g.direct = direct ? 1 : 0; g.y_bf16 = p.y_bf16; g.out_nhwc = p.out_nhwc; g.y = p.y;
  { const char* e = getenv("UPSNET_TMA_DEBUG"); g.dbg = e ? atoi(e) : 0; }
  { static int rot_env = -1; if (rot_env < 0) { const char* e = getenv("UPSNET_TMA_ROTATE"); rot_env = e ? atoi(e) : 0; } g.rotate = rot_env; }
  g.halo = halo ? 1 : 0; g.patch_rows = halo ? 16 * patch_h : 0; g.pstages = halo ? 3 : 0; g.kh = p.kh;
  int stages = TM_MAX_STAGES;
  TmaSmem L;
  // resident weights: the n-tile's KHW * Cin/64 weight tiles stay in smem for the life of the (persistent) CTA when they
  // fit next to >= 2 patch slots -- the 64->64 3x3 bottleneck convs and the 18-channel offset convs of the semantic head
  g.wres = 0;
  if (halo && Cout_pad == BN) {
    const int wtiles = g.KHW * (p.Cin / 64);
    for (int ps = 4; ps >= 2 && !g.wres; --ps) {
      L = tma_smem_layout(BN, wtiles, g.has_res != 0, g.patch_rows, ps, direct, pair, g.res_up2 != 0, g.res_inplace != 0, g.opairs);
      if (L.total + 1024 <= 227 * 1024) { g.wres = 1; g.wtiles = wtiles; g.pstages = ps; stages = 1; }
    }
  }
  if (!g.wres) {
    const uint32_t cap = 227 * 1024 - 1024;
    auto lay = [&](int st) { return tma_smem_layout(BN, st, g.has_res != 0, g.patch_rows, g.pstages, direct, pair, g.res_up2 != 0, g.res_inplace != 0, g.opairs); };
    // deepest ring that fits (<= TM_MAX_STAGES, >= 2); pair mode: if two alternating output slab pairs leave fewer than
    // four stages, one pair buys another stage -- the main loop needs the depth more than the epilogue does
    auto fit = [&]() {
      g.opairs = 2;
      stages = TM_MAX_STAGES;
      while (stages > 2 && lay(stages).total > cap) --stages;
      if (pair && !direct && !g.res_inplace && stages < 4) {
        g.opairs = 1;
        int st1 = 4;
        while (st1 > 2 && lay(st1).total > cap) --st1;
        if (st1 > stages || lay(stages).total > cap) stages = st1; else g.opairs = 2;
      }
      L = lay(stages);
      return L.total <= cap;
    };
    if (pair && halo) g.pstages = 2;       // pair patches are 2 x 36 KB (3x3): two slots, the rest goes to the weight ring
    bool ok = fit();
    if (!ok && g.pstages > 2) { g.pstages = 2; ok = fit(); }
    if (!ok && pair && halo) {
      // pair mode: the patch slots do not fit next to a weight ring for this N tile -- plain per-tap boxes instead
      halo = false;
      g.halo = 0; g.patch_rows = 0; g.pstages = 0;
      tma_pick_box(p.N, p.Ho, p.Wo, p.kh, p.kw, p.dh, p.dw, g.res_up2 != 0, &g.bw, &g.bh, &g.bn);
      g.tiles_w = (p.Wo + g.bw - 1) / g.bw;
      g.tiles_h = (p.Ho + g.bh - 1) / g.bh;
      g.tiles_n = (p.N + g.bn - 1) / g.bn;
      m_tiles = (long long)g.tiles_w * g.tiles_h * g.tiles_n;
      ok = fit();
    }
    if (!ok) return UPSNET_E_UNSUPPORTED;
  }
  g.stages = stages;

  const int Kp = g.KHW * p.Cin;
  CUtensorMap tm_x, tm_w, tm_y, tm_r;
  {
    const cuuint64_t xm = pair ? 2 : 1;      // pair tensors carry 2*C channels per pixel (hi plane, lo plane)
    const cuuint64_t dx[4] = {(cuuint64_t)p.Cin * xm, (cuuint64_t)(strided ? p.Wo : p.W), (cuuint64_t)(strided ? p.Ho : p.H),
                              (cuuint64_t)p.N};
    const cuuint64_t sx[3] = {(cuuint64_t)p.sw * p.Cin * 2 * xm, (cuuint64_t)p.sh * p.W * p.Cin * 2 * xm,
                              (cuuint64_t)p.H * p.W * p.Cin * 2 * xm};
    const cuuint64_t dy[4] = {(cuuint64_t)p.Cout * xm, (cuuint64_t)p.Wo, (cuuint64_t)p.Ho, (cuuint64_t)p.N};
    const cuuint64_t dwt[2] = {(cuuint64_t)Kp, (cuuint64_t)Cout_pad * xm};       // the packed weights ARE [hi plane][lo plane]
    const cuuint32_t box[4] = {64, (cuuint32_t)g.bw, (cuuint32_t)g.bh, (cuuint32_t)g.bn};
    const cuuint32_t boxp[4] = {64, 16, (cuuint32_t)patch_h, 1};            // halo mode: the input patch of a tile
    const cuuint32_t boxw[2] = {64, (cuuint32_t)BN};
    if (!encode_bf16(enc, &tm_x, p.x, 4, dx, halo ? boxp : box, sx)) return UPSNET_E_UNSUPPORTED;
    if (!encode_bf16(enc, &tm_w, packed, 2, dwt, boxw)) return UPSNET_E_UNSUPPORTED;
    if (direct) {   // no TMA store / residual in the direct-store epilogue: the two maps are placeholders
      tm_y = tm_x;
      tm_r = tm_x;
    } else {
    if (!encode_bf16(enc, &tm_y, p.y, 4, dy, box)) return UPSNET_E_UNSUPPORTED;
    if (g.res_up2) {
      const cuuint64_t dr[4] = {(cuuint64_t)p.Cout * xm, (cuuint64_t)(p.Wo / 2), (cuuint64_t)(p.Ho / 2), (cuuint64_t)p.N};
      const cuuint32_t boxr[4] = {64, (cuuint32_t)(g.bw / 2), (cuuint32_t)(g.bh / 2), (cuuint32_t)g.bn};
      if (!encode_bf16(enc, &tm_r, p.residual, 4, dr, boxr)) return UPSNET_E_UNSUPPORTED;
    } else if (!encode_bf16(enc, &tm_r, p.residual ? p.residual : p.y, 4, dy, box)) {
      return UPSNET_E_UNSUPPORTED;
    }
    }
  }
  const long long num_tiles = m_tiles * g.n_tiles;
  if (num_tiles <= 0) return 0;
  // ---- 2-CTA variant (cta_group::2): pair stream, per-tap boxes, no residual; each CTA stages half of the weight tile ----
  static int two_env = -1;
  if (two_env < 0) { const char* e = getenv("UPSNET_TMA_2CTA"); two_env = e ? atoi(e) : 1; }
  // measured (profiles/r2_2cta_ab.md): -10..-20 % on the long-K tiles (3x3, FC, 1x1 with Cin >= 512), but the cross-CTA
  // accumulator hand-shake costs more than the deeper ring buys when a tile has only 1-4 k-blocks (the HBM-bound 1x1 layers
  // of res2 / res3, +res or not): those keep the 1-CTA kernel.  UPSNET_TMA_2CTA=2 forces the pair kernel everywhere.
  const int num_kb_h = g.KHW * (p.Cin / 64);
  if (two_env > 0 && (two_env > 1 || num_kb_h >= 8) && pair && !g.halo && !g.stem && BN >= 32 && m_tiles >= 2 && sms >= 2 &&
      !(direct && g.has_res)) {
    const int opairs_1cta = g.opairs;
    int st2 = TM_MAX_STAGES;
    g.opairs = 2;
    static int wide_env = -1;
    if (wide_env < 0) { const char* e = getenv("UPSNET_TMA_WIDE"); wide_env = e ? atoi(e) : 1; }
    g.wide = (wide_env > 0 && BN <= 64 && !g.has_res) ? 1 : 0;
    auto lay2 = [&](int st, int op) { return tma2_smem_layout(BN, st, direct, op, g.has_res != 0, g.res_up2 != 0, g.res_inplace != 0, g.wide != 0); };
    Tma2Smem L2 = lay2(st2, g.opairs);
    while (st2 > 2 && L2.total + 1024 > 227 * 1024) { --st2; L2 = lay2(st2, g.opairs); }
    if (!direct && !g.res_inplace && st2 < 5) {           // one output slab pair buys a stage
      int st1 = st2;
      while (st1 < TM_MAX_STAGES && lay2(st1 + 1, 1).total + 1024 <= 227 * 1024) ++st1;
      if (st1 > st2) { st2 = st1; g.opairs = 1; L2 = lay2(st2, 1); }
    }
    if (L2.total + 1024 <= 227 * 1024) {
      g.stages = st2;
      CUtensorMap tm_w2;
      const cuuint64_t dwt2[2] = {(cuuint64_t)Kp, (cuuint64_t)Cout_pad * 2};
      const cuuint32_t boxw2[2] = {64, (cuuint32_t)(BN / 2)};
      if (encode_bf16(enc, &tm_w2, packed, 2, dwt2, boxw2)) {
        static ups::PerDeviceOnce configured2;
        if (configured2.need()) {
          UPS_CUDA(cudaFuncSetAttribute(igemm_tma2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        }
        const long long ctiles = ((m_tiles + 1) / 2) * g.n_tiles;
        const long long clusters = ctiles < sms / 2 ? ctiles : sms / 2;
        igemm_tma2_kernel<<<dim3((unsigned)(2 * clusters)), TM_THREADS, L2.total + 1024, stream>>>(tm_x, tm_w2, tm_y, tm_r, g);
        UPS_CHECK_LAUNCH();
        return 0;
      }
    }
    g.stages = stages; g.opairs = opairs_1cta; g.wide = 0;     // fall through to the 1-CTA kernel with its own geometry
  }
  {   // 1-CTA kernel, pair stream, N tile <= 64, per-tap boxes (incl. the in-place +res layers): two MMAs per K slice
    static int wide1_env = -1;
    if (wide1_env < 0) { const char* e = getenv("UPSNET_TMA_WIDE"); wide1_env = e ? atoi(e) : 1; }
    g.wide = (wide1_env > 0 && pair && BN <= 64 && !g.halo && !g.wres) ? 1 : 0;
  }
  static ups::PerDeviceOnce configured;
  if (configured.need()) {
    UPS_CUDA(cudaFuncSetAttribute(igemm_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  }
  dim3 grid((unsigned)(num_tiles < sms ? num_tiles : sms));
  igemm_tma_kernel<<<grid, TM_THREADS, L.total + 1024, stream>>>(tm_x, tm_w, tm_y, tm_r, g);
  UPS_CHECK_LAUNCH();
  return 0;
}


// ----------------------------------------------------------------------------------------------
// RGB stem (models/resnet.py:155-162 conv1: 7x7 / stride 2 / pad 3, Cin = 3) on the TMA kernel.
// The fp32 NCHW image is first packed to a zero-padded bf16 NHWC8 image (stem_pack_image_kernel); the eight input
// pixels x 8 channels an output pixel needs from filter row ky are then 128 contiguous bytes, consecutive output
// pixels start 32 bytes apart, and even / odd input rows are split by a parity dimension -- a 5-D tensor map
// (64 el, Wo @32 B, 2 @pitch, Hp/2 @2*pitch, N) whose boxes ARE the im2col tiles, one k-block per filter row.
// Weights are packed [Cout][ky][8 px][8 ch] with zeros for kx >= kw and c >= Cin (K = 64*kh).
// ----------------------------------------------------------------------------------------------
__global__ void stem_pack_image_kernel(const float* __restrict__ x, uint4* __restrict__ xp, uint4* __restrict__ xp_lo, int N, int C,
                                       int H, int W, int pad, int Hp, int Wp) {
  const long long total = (long long)N * Hp * Wp;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(t % Wp);
    const long long rest = t / Wp;
    const int r = (int)(rest % Hp), n = (int)(rest / Hp);
    const int hi = r - pad, wi = c - pad;
    float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (hi >= 0 && hi < H && wi >= 0 && wi < W)
      for (int ch = 0; ch < C; ++ch) v[ch] = __ldg(x + (((size_t)n * C + ch) * H + hi) * W + wi);
    uint4 o;
    o.x = pack_bf16x2(v[0], v[1]); o.y = pack_bf16x2(v[2], v[3]); o.z = pack_bf16x2(v[4], v[5]); o.w = pack_bf16x2(v[6], v[7]);
    xp[t] = o;
    if (xp_lo) {     // pair stem: second plane = bf16(v - hi)
      uint4 l;
      l.x = pack_bf16x2(v[0] - __uint_as_float(o.x << 16), v[1] - __uint_as_float(o.x & 0xffff0000u));
      l.y = pack_bf16x2(v[2] - __uint_as_float(o.y << 16), v[3] - __uint_as_float(o.y & 0xffff0000u));
      l.z = pack_bf16x2(v[4] - __uint_as_float(o.z << 16), v[5] - __uint_as_float(o.z & 0xffff0000u));
      l.w = pack_bf16x2(v[6] - __uint_as_float(o.w << 16), v[7] - __uint_as_float(o.w & 0xffff0000u));
      xp_lo[t] = l;
    }
  }
}

__global__ void stem_pack_weight_kernel(const float* __restrict__ w, int Cout, int Cin, int kh, int kw,
                                        __nv_bfloat16* __restrict__ packed) {
  const int total = Cout * kh * 64;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int c = i & 7, kx = (i >> 3) & 7, ky = (i >> 6) % kh, co = i / (64 * kh);
    const float v = (c < Cin && kx < kw) ? w[(((size_t)co * Cin + c) * kh + ky) * kw + kx] : 0.f;
    const __nv_bfloat16 h = __float2bfloat16_rn(v);
    packed[i] = h;
    packed[total + i] = __float2bfloat16_rn(v - __bfloat162float(h));      // lo plane (pair stem)
  }
}

static void stem_geometry(int H, int W, int kh, int kw, int pad, int* Ho, int* Wo, int* Hp, int* Wp) {
  *Ho = (H + 2 * pad - kh) / 2 + 1;
  *Wo = (W + 2 * pad - kw) / 2 + 1;
  *Hp = 2 * (*Ho + (kh >> 1));
  *Wp = 2 * *Wo + 8;
}

}  // namespace ups

extern "C" int upsnet_stem_workspace_bytes(int N, int H, int W, int kh, int kw, int pad, size_t* bytes) {
  if (!bytes || N <= 0 || H <= 0 || W <= 0 || kh <= 0 || kw <= 0 || kw > 8 || pad < 0) return UPSNET_E_BADARG;
  int Ho, Wo, Hp, Wp;
  ups::stem_geometry(H, W, kh, kw, pad, &Ho, &Wo, &Hp, &Wp);
  if (Ho <= 0 || Wo <= 0) return UPSNET_E_BADARG;
  *bytes = (size_t)N * Hp * Wp * 16 * 2;     // hi plane + lo plane (the lo plane is only written / read by the pair stem)
  return 0;
}

extern "C" int upsnet_stem_packed_weight_bytes(int Cout, int kh, size_t* bytes) {
  if (!bytes || Cout <= 0 || kh <= 0) return UPSNET_E_BADARG;
  *bytes = (size_t)Cout * kh * 64 * 2 * 2;   // bf16 hi plane + lo plane
  return 0;
}

extern "C" int upsnet_stem_pack_weight(const float* weight, int Cout, int Cin, int kh, int kw, void* packed, void* stream) {
  if (!weight || !packed || Cout <= 0 || Cin <= 0 || Cin > 8 || kh <= 0 || kw <= 0 || kw > 8) return UPSNET_E_BADARG;
  const int total = Cout * kh * 64;
  ups::stem_pack_weight_kernel<<<(total + 255) / 256, 256, 0, (cudaStream_t)stream>>>(weight, Cout, Cin, kh, kw,
                                                                                         (__nv_bfloat16*)packed);
  UPS_CHECK_LAUNCH();
  return 0;
}

extern "C" int upsnet_stem_forward(const float* x, const void* packed_w, const float* bias, void* y, int N, int Cin, int H,
                                   int W, int Cout, int kh, int kw, int pad, int epi_flags, void* workspace,
                                   size_t workspace_bytes, void* stream) {
  using namespace ups;
  if (!x || !packed_w || !y || !workspace) return UPSNET_E_BADARG;
  if (N <= 0 || Cin <= 0 || Cin > 8 || H <= 0 || W <= 0 || kh <= 0 || kw <= 0 || kw > 8 || pad < 0) return UPSNET_E_BADARG;
  if ((Cout % 64) || Cout > 256 || kh > 16) return UPSNET_E_UNSUPPORTED;
  if ((((uintptr_t)y) & 15) || (((uintptr_t)packed_w) & 15) || (((uintptr_t)workspace) & 15)) return UPSNET_E_BADARG;
  if (bias && (((uintptr_t)bias) & 15)) return UPSNET_E_UNSUPPORTED;
  int Ho, Wo, Hp, Wp;
  stem_geometry(H, W, kh, kw, pad, &Ho, &Wo, &Hp, &Wp);
  if (Ho <= 0 || Wo <= 0) return UPSNET_E_BADARG;
  const bool pair = (epi_flags & UPSNET_EPI_STEM_PAIR) != 0;
  const size_t plane_bytes = (size_t)N * Hp * Wp * 16;
  if (workspace_bytes < plane_bytes * (pair ? 2 : 1)) return UPSNET_E_WORKSPACE;
  EncodeTiledFn enc = tma_encoder();
  if (!enc) return UPSNET_E_UNSUPPORTED;
  cudaStream_t st = (cudaStream_t)stream;
  TmaGeom g{};
  g.bias = bias;
  g.N = N; g.Ho = Ho; g.Wo = Wo; g.Cout = Cout; g.Cin = 64;
  g.kw = 1; g.KHW = kh; g.ph = 0; g.pw = 0; g.dh = 1; g.dw = 1;
  g.relu = (epi_flags & UPSNET_EPI_RELU) ? 1 : 0;
  g.sig_from = -1;
  g.stem = 1; g.y = y; g.y_bf16 = 1; g.out_nhwc = 1;
  g.x3 = pair ? 1 : 0; g.w_lo = Cout; g.pg = Cout; g.opairs = 2; g.x_lo = 0;
  tma_pick_box(N, Ho, Wo, 1, 1, 1, 1, false, &g.bw, &g.bh, &g.bn);
  g.tiles_w = (Wo + g.bw - 1) / g.bw;
  g.tiles_h = (Ho + g.bh - 1) / g.bh;
  g.tiles_n = (N + g.bn - 1) / g.bn;
  g.BN = Cout % 256 == 0 ? 256 : (Cout % 128 == 0 ? 128 : 64);
  if (pair && g.BN > 128) g.BN = 128;
  g.n_tiles = Cout / g.BN;
  {   // pair stem, N tile 64: two MMAs per K slice over [W_hi ; W_lo] (see TmaGeom::wide)
    const char* e = getenv("UPSNET_TMA_WIDE");
    g.wide = (pair && g.BN <= 64 && (!e || atoi(e) > 0)) ? 1 : 0;
  }
  int stages = TM_MAX_STAGES;
  TmaSmem L = tma_smem_layout(g.BN, stages, false, 0, 0, false, pair, false, false, g.opairs);
  while (stages > 2 && L.total + 1024 > 227 * 1024) { --stages; L = tma_smem_layout(g.BN, stages, false, 0, 0, false, pair, false, false, g.opairs); }
  g.stages = stages;
  CUtensorMap tm_x, tm_w, tm_y, tm_lo;
  {
    const cuuint64_t pitch = (cuuint64_t)Wp * 16;
    const cuuint64_t dx[5] = {64, (cuuint64_t)Wo, 2, (cuuint64_t)(Hp / 2), (cuuint64_t)N};
    const cuuint64_t sx[4] = {32, pitch, 2 * pitch, (cuuint64_t)Hp * pitch};
    const cuuint32_t bx[5] = {64, (cuuint32_t)g.bw, 1, (cuuint32_t)g.bh, (cuuint32_t)g.bn};
    const cuuint64_t dwt[2] = {(cuuint64_t)kh * 64, (cuuint64_t)Cout * 2};       // hi plane rows, then lo plane rows
    const cuuint32_t bw2[2] = {64, (cuuint32_t)g.BN};
    const cuuint64_t dy[4] = {(cuuint64_t)Cout * (pair ? 2 : 1), (cuuint64_t)Wo, (cuuint64_t)Ho, (cuuint64_t)N};
    const cuuint32_t by[4] = {64, (cuuint32_t)g.bw, (cuuint32_t)g.bh, (cuuint32_t)g.bn};
    if (!encode_bf16(enc, &tm_x, workspace, 5, dx, bx, sx)) return UPSNET_E_UNSUPPORTED;
    tm_lo = tm_x;
    if (pair && !encode_bf16(enc, &tm_lo, (char*)workspace + plane_bytes, 5, dx, bx, sx)) return UPSNET_E_UNSUPPORTED;
    if (!encode_bf16(enc, &tm_w, packed_w, 2, dwt, bw2)) return UPSNET_E_UNSUPPORTED;
    if (!encode_bf16(enc, &tm_y, y, 4, dy, by)) return UPSNET_E_UNSUPPORTED;
  }
  {
    const long long total = (long long)N * Hp * Wp;
    long long blocks = (total + 255) / 256;
    if (blocks > kNumSMs * 32) blocks = kNumSMs * 32;
    stem_pack_image_kernel<<<(unsigned)blocks, 256, 0, st>>>(x, (uint4*)workspace, pair ? (uint4*)((char*)workspace + plane_bytes) : nullptr,
                                                             N, Cin, H, W, pad, Hp, Wp);
    UPS_CHECK_LAUNCH();
  }
  static ups::PerDeviceOnce configured;
  if (configured.need()) {
    UPS_CUDA(cudaFuncSetAttribute(igemm_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  }
  static int sms = 0;
  if (sms == 0) {
    int dev = 0, v = kNumSMs;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
    sms = v > 0 ? v : kNumSMs;
  }
  const long long num_tiles = (long long)g.tiles_w * g.tiles_h * g.tiles_n * g.n_tiles;
  dim3 grid((unsigned)(num_tiles < sms ? num_tiles : sms));
  igemm_tma_kernel<<<grid, TM_THREADS, L.total + 1024, st>>>(tm_x, tm_w, tm_y, pair ? tm_lo : tm_y, g);
  UPS_CHECK_LAUNCH();
  return 0;
}
