// dcn_win.cu -- fused deformable convolution (v1 / v2) on hi/lo bf16 PAIR activations with the bilinear corners
// gathered from a SHARED-MEMORY WINDOW instead of global memory (sm_100a).
//
// Reference semantics: operators/src/deform_conv_kernel.cu:89-118 (bilinear corner rule), :194-242 (im2col, the
// h > -1 && w > -1 && h < H && w < W test at :229), operators/functions/deform_conv.py:44-57 (im2col + torch.mm);
// v2 mask: operators/src/mod_deform_conv_kernel.cu.  Same arithmetic contract as igemm_tc_kernel<1,2> (igemm_tc.cu),
// which this kernel replaces for 3x3 / stride-1 layers on the pair stream: round 2's profile of that kernel
// (profiles/r2_pair_full.md) shows the gather ISSUE-bound -- ~300 instructions per (pixel, tap, 8 channels), a third of
// them 64-bit address arithmetic and unpacking around eight long-latency LDG.128 -- with the tensor pipe at 18 %.
//
// Here the K axis is ordered (16-channel sub-chunk, tap, channel): for one sub-chunk the nine taps x four corners of a
// 16 x 8-pixel tile touch one small window of the input -- (16 + 3 + offset range) x (8 + 3 + offset range) pixels x
// 16 channels x (hi, lo) -- which ONE TMA box pair stages in shared memory (30 x 20 pixels, 2 x 19.2 KB, double buffered).
// The 16 gather warps then read corners with LDS.128 at immediate offsets (+32 B = x+1, +960 B = y+1), blend (hi plane:
// packed fp32x2 FMAs, lo plane: packed bf16x2 HFMA2), re-split and store the A tile in the SWIZZLE_128B K-major layout
// tcgen05 consumes; weights (packed in the same K order by dcn_win_pack_weight) arrive by TMA.  Samples that fall
// outside the window (large offsets) are flagged in the per-tile sample table and gathered from global memory by the
// same thread, so the result never depends on the window size -- only the speed does.
//
// Warp roles (23 warps): 0-3 epilogue (TMEM -> bias / ReLU -> hi/lo split -> NHWC pair stores), 4 MMA issuer (three
// tcgen05.mma per K=16 slice: lo*hi, hi*lo, hi*hi; two TMEM accumulator buffers), 5 weight TMA, 6 window TMA,
// 7-22 gather producers in two groups that fill alternate k-blocks.
// Roofline: tensor pipe (2*P*Cout*Cin*9 flop x 3 passes); per k-block and SM the gather costs ~4 K warp instructions
// and 128 KB of shared-memory reads against 12 MMAs of 128x128x16 -- see DESIGN.md section 4.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cstdlib>

#include "common.cuh"
#include "tc_ptx.cuh"
#include "tc_params.cuh"

namespace ups {

constexpr int DW_BM = 128;            // pixels per M tile
constexpr int DW_BK = 64;             // K elements per k-block = 4 slices of 16 channels
constexpr int DW_STAGES = 3;          // operand ring depth: A stages live in TMEM (gather warps), B stages in smem (TMA)
constexpr int DW_WW = 32, DW_WH = 24; // window box in pixels; the row pitch (32 px = 1024 B) keeps bank = f(x) only
constexpr int DW_PLANE = DW_WW * DW_WH * 32;      // one plane (hi or lo) of a window: 16 channels x 2 B per pixel
constexpr int DW_WIN_BYTES = 2 * DW_PLANE;
constexpr int DW_KHW = 9;
constexpr int DW_GROUP = 256, DW_GROUPS = 2, DW_PRODUCERS = DW_GROUP * DW_GROUPS;
constexpr int DW_WARP_MMA = 4, DW_WARP_TMAB = 5, DW_WARP_TMAW = 6, DW_WARP_PROD0 = 7;
constexpr int DW_THREADS = (DW_WARP_PROD0 + DW_PRODUCERS / 32) * 32;   // 736
constexpr uint32_t DW_TMEM_A0 = 256;  // TMEM columns [0,256): two accumulators; [256 + 64 s, +64): A stage s (hi 32 cols, lo 32 cols)

// shared-memory map (byte offsets from the 1024-aligned base)
constexpr uint32_t DW_OFF_BARS = 0;        // 18 mbarriers
constexpr uint32_t DW_OFF_TMEM = 160;
constexpr uint32_t DW_OFF_STATS = 192;     // 2 x int[8]: min w, min h, max w, max h, sum w, sum h, count, -
constexpr uint32_t DW_OFF_ORG = 256;       // 2 x int4: window origin (w, h), image, -
constexpr uint32_t DW_OFF_TW = 512;        // float4 [9][128] corner weights
constexpr uint32_t DW_OFF_TP = DW_OFF_TW + DW_KHW * DW_BM * 16;   // int [9][128] window byte offset / outlier code
constexpr uint32_t DW_OFF_STAGES = DW_OFF_TP + DW_KHW * DW_BM * 4;   // 23552 = 23 * 1024
static_assert(DW_OFF_STAGES % 1024 == 0, "stage buffers need 1024-byte alignment (SWIZZLE_128B)");

struct DwParams {
  const void* x;          // pair NHWC [N,H,W,2*Cin] (outlier gathers)
  const float* offset;    // NCHW fp32 [N,18,Ho,Wo]
  const float* mask;      // NCHW fp32 [N,9,Ho,Wo] or null
  const float* bias;
  void* y;                // pair NHWC [N,Ho,Wo,2*Cout]
  int N, H, W, Cin, Cout, Cout_pad, Ho, Wo, ph, pw, dh, dw, relu, BN, tile_w, tile_h;
  // DENSE mode (offset == null): plain 3x3 / stride-1 convolution through the same pipeline -- the window of a tile is its
  // receptive field (origin = tile origin - pad, known without a sample table), a "gather" is one LDS.128 per plane copied
  // to the TMEM A operand, no blend.  Serves the small-N 3x3 layers of the pair stream (18-channel offset convs, 64->64
  // bottleneck convs), which the per-tap TMA boxes of igemm_tma.cu make L2->SM-bandwidth-bound (A re-fetched per tap).
  int dense;
  int out_nchw;           // y = fp32 NCHW [N,Cout,Ho,Wo] (offset maps) instead of the pair NHWC tensor
  int win_bytes;          // bytes one window fill delivers (both planes): the dense box is only tile + halo rows high
  int win_h;              // rows of the window box (<= DW_WH)
  // DENSE mode window geometry: one fill = a 64-channel chunk (four 16-channel sub-chunks), 128-byte pixels, SWIZZLE_128B --
  // the 32-byte box rows of the deformable window would make the plain copy loop TMA-request-bound (~4.5 cycles per box row)
  int win_pitch;          // pixels per window row (dense: 24)
  int win_plane;          // bytes of one plane of a window buffer
  int win_buf;            // bytes of one window buffer (both planes)
  int stages;             // operand ring depth (2 or 3)
};

__device__ __forceinline__ void dw_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void dw_tma_4d(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void dw_tma_2d(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void dw_prefetch_tmap(const CUtensorMap* tm) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(tm)) : "memory");
}
__device__ __forceinline__ uint4 dw_lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void dw_producer_bar() { asm volatile("bar.sync 1, %0;" ::"n"(DW_PRODUCERS) : "memory"); }
// registers -> TMEM: lane i of the warp writes four consecutive 32-bit columns of TMEM lane (quadrant base + i)
__device__ __forceinline__ void dw_tmem_st4(uint32_t taddr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(taddr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void dw_tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// D[tmem] (+)= A[tmem] * B[smem desc]^T, kind::f16: A = 128 TMEM lanes x 8 columns (16 bf16 of K, two per column)
__device__ __forceinline__ void dw_umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint32_t b_lo, uint32_t desc_hi, uint32_t idesc,
                                           uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 db;\n\t"
      "mov.b64 db, {%2, %5};\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "r"(b_lo), "r"(idesc), "r"(accumulate), "r"(desc_hi) : "memory");
}

__global__ void __launch_bounds__(DW_THREADS, 1)
dcn_win_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_w, const DwParams p) {
  extern __shared__ __align__(1024) uint8_t smem_dyn[];
  const uint32_t raw = smem_u32(smem_dyn);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* sm = smem_dyn + (base - raw);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t b_bytes = (uint32_t)p.BN * 128;
  const uint32_t stage_bytes = 2 * b_bytes;                                 // weight tile: hi plane, lo plane
  const uint32_t NST = (uint32_t)p.stages;
  const uint32_t win_base = base + DW_OFF_STAGES + NST * stage_bytes;       // 1024-aligned (stage_bytes % 1024 == 0)
  // barriers
  const uint32_t bar_fa = base + DW_OFF_BARS;            // full_a[3]: 8 producer warps each (A stage written to TMEM)
  const uint32_t bar_fb = bar_fa + 24;                   // full_b[3]: weight TMA (tx)
  const uint32_t bar_em = bar_fa + 48;                   // empty[3]: tcgen05.commit (A stage in TMEM + B stage in smem consumed)
  const uint32_t bar_tf = bar_fa + 72;                   // tmem_full[2]
  const uint32_t bar_te = bar_fa + 88;                   // tmem_empty[2]: 4 epilogue warps
  const uint32_t bar_wf = bar_fa + 104;                  // win_full[2]: window TMA (tx)
  const uint32_t bar_we = bar_fa + 120;                  // win_empty[2]: 16 producer warps
  const uint32_t bar_og = bar_fa + 136;                  // org_full[2]: window origin of a tile published
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(sm + DW_OFF_TMEM);
  int* stats = reinterpret_cast<int*>(sm + DW_OFF_STATS);
  int4* org = reinterpret_cast<int4*>(sm + DW_OFF_ORG);
  float4* tw = reinterpret_cast<float4*>(sm + DW_OFF_TW);
  int* tp = reinterpret_cast<int*>(sm + DW_OFF_TP);

  const int HoWo = p.Ho * p.Wo;
  const int nsc = p.Cin / 16;                         // 16-channel sub-chunks
  const int spf = p.dense ? 4 : 1;                    // sub-chunks per window fill
  const int nfill = nsc / spf;                        // window fills per tile
  const int num_kb = nsc * DW_KHW / 4;                // Cin % 64 == 0 -> integral
  const int n_tiles = p.Cout_pad / p.BN;
  const int TW = p.tile_w, TH = p.tile_h;
  const int tw_shift = TW == 16 ? 4 : 3;
  const int tiles_w = (p.Wo + TW - 1) / TW, tiles_h = (p.Ho + TH - 1) / TH;
  const long long num_tiles = (long long)p.N * tiles_w * tiles_h * n_tiles;
  constexpr uint32_t tmem_cols = 512;

  if (warp == DW_WARP_MMA) {
    if (lane == 0) {
      for (int s = 0; s < DW_STAGES; ++s) {
        mbar_init(bar_fa + 8 * s, DW_GROUP / 32);
        mbar_init(bar_fb + 8 * s, 1);
        mbar_init(bar_em + 8 * s, 1);
      }
      for (int s = 0; s < 2; ++s) {
        mbar_init(bar_tf + 8 * s, 1);
        mbar_init(bar_te + 8 * s, 4);
        mbar_init(bar_wf + 8 * s, 1);
        mbar_init(bar_we + 8 * s, DW_PRODUCERS / 32);
        mbar_init(bar_og + 8 * s, 1);
      }
      fence_mbar_init();
      for (int i = 0; i < 2; ++i) {
        stats[i * 8 + 0] = 0x7fffffff; stats[i * 8 + 1] = 0x7fffffff;
        stats[i * 8 + 2] = -0x7fffffff; stats[i * 8 + 3] = -0x7fffffff;
        stats[i * 8 + 4] = 0; stats[i * 8 + 5] = 0; stats[i * 8 + 6] = 0; stats[i * 8 + 7] = 0;
      }
    }
    __syncwarp();
    tmem_alloc(smem_u32(tmem_ptr_smem), tmem_cols);
  }
  if (warp == DW_WARP_TMAB && lane == 0) dw_prefetch_tmap(&tm_w);
  if (warp == DW_WARP_TMAW && lane == 0) dw_prefetch_tmap(&tm_x);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp >= DW_WARP_PROD0) {
    // =============================== GATHER PRODUCERS ===============================
    // A warp owns the 32 rows of ITS TMEM lane quadrant (hardware rule: warp w reaches lanes 32 (w % 4) .. +31), lane = row:
    // the 32 lanes of an LDS.128 read one (slice, 8-channel half) of 32 consecutive tile pixels.  The window is stored by
    // TMA with SWIZZLE_32B (16-byte chunk ^= address bit 7, i.e. pixel bit 2), so eight horizontally consecutive pixels of
    // one half occupy eight different 16-byte bank groups: a quarter-warp wavefront is conflict-free whenever its samples
    // stay on consecutive pixels of any rows (row pitch 1024 B).
    const int pt = tid - DW_WARP_PROD0 * 32;     // 0..511: sample-table work is spread over all producer threads
    const int pw = warp - DW_WARP_PROD0;         // 0..15
    const int group = pw >> 3;                   // alternate k-blocks
    const int u = (pw >> 2) & 1;                 // which two of the k-block's four slices this warp gathers
    const int quad = warp & 3;                   // TMEM lane quadrant
    const int r = quad * 32 + lane;              // A row = tile pixel
    const __nv_bfloat16* xh = reinterpret_cast<const __nv_bfloat16*>(p.x);
    uint32_t g0 = 0, wf0 = 0;                    // running k-block / window-fill counters at the start of the tile
    uint32_t wf_ready = 0;                       // window fills [0, wf_ready) have been observed complete by this thread
    uint32_t tile_it = 0;
    const bool dense = p.dense != 0;
    if (dense) {
      // tile-independent table: window offset of (tap, tile pixel) relative to the tile's receptive-field origin
      for (int e = pt; e < DW_KHW * DW_BM; e += DW_PRODUCERS) {
        const int tap = e >> 7, rr = e & 127;
        const int ki = tap / 3, kj = tap - ki * 3;
        const int ry = min(rr >> tw_shift, TH - 1), rx = rr & (TW - 1);      // rows past the block read a valid (unused) pixel
        tp[e] = (ry + ki * p.dh) * p.win_pitch + rx + kj * p.dw;      // window PIXEL index
      }
      dw_producer_bar();
    }
    for (long long tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++tile_it) {
      const long long mt = tile / n_tiles;
      const int tx = (int)(mt % tiles_w), ty = (int)((mt / tiles_w) % tiles_h), n = (int)(mt / ((long long)tiles_w * tiles_h));
      const int par = (int)(tile_it & 1u);
      if (!dense) {
      // ---- sample table, phase 1: every thread computes up to three (tap, pixel) entries in registers ----
      float4 ewv[3];
      int ehl[3], ewl[3];
      bool evalid[3];
      int mnw = 0x7fffffff, mnh = 0x7fffffff, mxw = -0x7fffffff, mxh = -0x7fffffff, sw_ = 0, sh_ = 0, cnt = 0;
#pragma unroll
      for (int it = 0; it < 3; ++it) {
        const int e = pt + it * DW_PRODUCERS;
        ewv[it] = make_float4(0.f, 0.f, 0.f, 0.f);
        ehl[it] = 0; ewl[it] = 0; evalid[it] = false;
        if (e < DW_KHW * DW_BM) {
          const int tap = e >> 7, rr = e & 127;
          const int ry = rr >> tw_shift;
          const int wo = tx * TW + (rr & (TW - 1)), ho = ty * TH + ry;
          if (ry < TH && wo < p.Wo && ho < p.Ho) {
            const int pp = ho * p.Wo + wo;
            const int ki = tap / 3, kj = tap - ki * 3;
            const float* offp = p.offset + ((size_t)n * 2 * DW_KHW + 2 * tap) * HoWo + pp;
            const float oh = __ldg(offp), ow = __ldg(offp + HoWo);
            const float h = (float)(ho - p.ph + ki * p.dh) + oh;
            const float w = (float)(wo - p.pw + kj * p.dw) + ow;
            if (h > -1.f && w > -1.f && h < (float)p.H && w < (float)p.W) {      // deform_conv_kernel.cu:229
              const int hl = (int)floorf(h), wl = (int)floorf(w), hh = hl + 1, wh = wl + 1;
              const float lh = h - hl, lw = w - wl, ch = 1.f - lh, cw = 1.f - lw;
              const bool t_ok = hl >= 0, b_ok = hh <= p.H - 1, l_ok = wl >= 0, r_ok = wh <= p.W - 1;
              float m = 1.f;
              if (p.mask) m = __ldg(p.mask + ((size_t)n * DW_KHW + tap) * HoWo + pp);
              ewv[it].x = (t_ok && l_ok) ? ch * cw * m : 0.f;
              ewv[it].y = (t_ok && r_ok) ? ch * lw * m : 0.f;
              ewv[it].z = (b_ok && l_ok) ? lh * cw * m : 0.f;
              ewv[it].w = (b_ok && r_ok) ? lh * lw * m : 0.f;
              ehl[it] = hl; ewl[it] = wl; evalid[it] = true;
              mnw = min(mnw, wl); mxw = max(mxw, wl + 1); mnh = min(mnh, hl); mxh = max(mxh, hl + 1);
              sw_ += wl; sh_ += hl; ++cnt;
            }
          }
        }
      }
      mnw = __reduce_min_sync(0xffffffffu, mnw); mnh = __reduce_min_sync(0xffffffffu, mnh);
      mxw = __reduce_max_sync(0xffffffffu, mxw); mxh = __reduce_max_sync(0xffffffffu, mxh);
      sw_ = __reduce_add_sync(0xffffffffu, sw_); sh_ = __reduce_add_sync(0xffffffffu, sh_);
      cnt = __reduce_add_sync(0xffffffffu, cnt);
      if (lane == 0 && cnt > 0) {
        int* st = stats + par * 8;
        atomicMin(st + 0, mnw); atomicMin(st + 1, mnh); atomicMax(st + 2, mxw); atomicMax(st + 3, mxh);
        atomicAdd(st + 4, sw_); atomicAdd(st + 5, sh_); atomicAdd(st + 6, cnt);
      }
      dw_producer_bar();      // (B) statistics complete; every producer has also finished the previous tile's gather
      // ---- window origin (same integer arithmetic in every thread), table phase 2 ----
      int ox = 0, oy = 0;
      {
        const int* st = stats + par * 8;
        const int c = st[6];
        if (c > 0) {
          const int a = st[0], b = st[1], cc = st[2], d = st[3];
          // the bounding box of all corners fits: start the window there; otherwise centre it on the mean sample
          ox = (cc - a + 1 <= DW_WW) ? a : (int)floorf((float)st[4] / (float)c + 1.0f) - DW_WW / 2;
          oy = (d - b + 1 <= p.win_h) ? b : (int)floorf((float)st[5] / (float)c + 1.0f) - p.win_h / 2;
        }
      }
#pragma unroll
      for (int it = 0; it < 3; ++it) {
        const int e = pt + it * DW_PRODUCERS;
        if (e < DW_KHW * DW_BM) {
          int code = 0;
          if (evalid[it]) {
            const int dx = ewl[it] - ox, dy = ehl[it] - oy;
            if (dx >= 0 && dx + 1 < DW_WW && dy >= 0 && dy + 1 < p.win_h) code = (dy * DW_WW + dx) * 32;
            else code = (int)(0x80000000u | ((uint32_t)(ehl[it] + 1) << 15) | (uint32_t)(ewl[it] + 1));   // outlier: global gather
          }
          tw[e] = ewv[it];
          tp[e] = code;
        }
      }
      if (pt == 0) {
        org[par] = make_int4(ox, oy, n, 0);
        int* so = stats + (par ^ 1) * 8;      // reset the other slot for the next tile (last read before barrier B of this tile)
        so[0] = 0x7fffffff; so[1] = 0x7fffffff; so[2] = -0x7fffffff; so[3] = -0x7fffffff; so[4] = 0; so[5] = 0; so[6] = 0;
        mbar_arrive(bar_og + 8 * par);        // release: the window TMA thread may read the origin
      }
      dw_producer_bar();      // (C) table visible
      }                       // !dense
      const __nv_bfloat16* ximg = xh + (size_t)n * p.H * p.W * (size_t)(2 * p.Cin);

      int rel = 0;            // sub-chunks of this tile this WARP has released
      for (int kb = (int)((uint32_t)(group - (int)g0) & 1u); kb < num_kb; kb += DW_GROUPS) {
        const uint32_t g = g0 + (uint32_t)kb;
        const uint32_t s = g % NST, it = g / NST;
        mbar_wait(bar_em + 8 * s, (it & 1u) ^ 1u);
        tc_fence_after();
        const uint32_t a_col = tmem_base + ((uint32_t)(quad * 32) << 16) + DW_TMEM_A0 + s * 64u;
#pragma unroll 1
        for (int pass = 0; pass < 4; ++pass) {
          const int sl = 2 * u + (pass >> 1), half = pass & 1;
          const int q = kb * 4 + sl;                    // slice index: (sub-chunk, tap)
          const int sc = q / DW_KHW, tap = q - sc * DW_KHW;
          const uint32_t wf = wf0 + (uint32_t)(dense ? (sc >> 2) : sc);      // dense: one fill per 64-channel chunk
          if (wf >= wf_ready) {                         // first touch of this window fill
            mbar_wait(bar_wf + 8 * (wf & 1u), (wf >> 1) & 1u);
            wf_ready = wf + 1;
          }
          const uint32_t wbuf = win_base + (wf & 1u) * (uint32_t)p.win_buf;
          const int code = tp[tap * DW_BM + r];
          if (dense) {        // plain copy of the tap's pixel: window -> TMEM A operand
            // SWIZZLE_128B window: a pixel is one 128-byte row (64 channels), 16-byte chunk c sits at c ^ (pixel & 7)
            const uint32_t pix = (uint32_t)code, ch = (uint32_t)((sc & 3) * 2 + half);
            const uint32_t al = wbuf + pix * 128u + ((ch ^ (pix & 7u)) << 4);
            const uint4 h4 = dw_lds128(al), l4 = dw_lds128(al + (uint32_t)p.win_plane);
            const uint32_t col = a_col + (uint32_t)(sl * 8 + half * 4);
            dw_tmem_st4(col, h4.x, h4.y, h4.z, h4.w);
            dw_tmem_st4(col + 32u, l4.x, l4.y, l4.z, l4.w);
            continue;
          }
          const float4 wv = tw[tap * DW_BM + r];
          uint4 hc[4], lc[4];
          if (code >= 0) {
            // SWIZZLE_32B: the 16-byte chunk of a pixel sits at (half ^ bit 7 of the pixel's byte offset)
            const uint32_t cl = (uint32_t)code, cr = cl + 32u;
            const uint32_t al = wbuf + cl + ((((cl >> 7) & 1u) ^ (uint32_t)half) << 4);
            const uint32_t ar = wbuf + cr + ((((cr >> 7) & 1u) ^ (uint32_t)half) << 4);
            hc[0] = dw_lds128(al); hc[1] = dw_lds128(ar); hc[2] = dw_lds128(al + DW_WW * 32); hc[3] = dw_lds128(ar + DW_WW * 32);
            lc[0] = dw_lds128(al + DW_PLANE); lc[1] = dw_lds128(ar + DW_PLANE);
            lc[2] = dw_lds128(al + DW_PLANE + DW_WW * 32); lc[3] = dw_lds128(ar + DW_PLANE + DW_WW * 32);
          } else {
            // outlier sample: the four corners come from global memory (clamped addresses; invalid corners carry weight 0)
            const int hl = (int)(((uint32_t)code >> 15) & 0xffffu) - 1, wl = (int)((uint32_t)code & 0x7fffu) - 1;
            const int h0 = max(hl, 0), h1 = min(hl + 1, p.H - 1), w0 = max(wl, 0), w1 = min(wl + 1, p.W - 1);
            const size_t pc = (size_t)(2 * p.Cin);
            const int cg = sc * 16 + half * 8;            // first channel of this pass's 8-channel vector
            const __nv_bfloat16* b00 = ximg + ((size_t)h0 * p.W + w0) * pc + cg;
            const __nv_bfloat16* b01 = ximg + ((size_t)h0 * p.W + w1) * pc + cg;
            const __nv_bfloat16* b10 = ximg + ((size_t)h1 * p.W + w0) * pc + cg;
            const __nv_bfloat16* b11 = ximg + ((size_t)h1 * p.W + w1) * pc + cg;
            hc[0] = __ldg(reinterpret_cast<const uint4*>(b00)); lc[0] = __ldg(reinterpret_cast<const uint4*>(b00 + p.Cin));
            hc[1] = __ldg(reinterpret_cast<const uint4*>(b01)); lc[1] = __ldg(reinterpret_cast<const uint4*>(b01 + p.Cin));
            hc[2] = __ldg(reinterpret_cast<const uint4*>(b10)); lc[2] = __ldg(reinterpret_cast<const uint4*>(b10 + p.Cin));
            hc[3] = __ldg(reinterpret_cast<const uint4*>(b11)); lc[3] = __ldg(reinterpret_cast<const uint4*>(b11 + p.Cin));
          }
          // blend: hi plane in packed fp32x2 (exact products of bf16 values), lo plane in packed bf16x2 (2^-9 of the
          // value: its blend needs 2^-9 relative accuracy only), summed in fp32 and split again
          const float wf4[4] = {wv.x, wv.y, wv.z, wv.w};
          unsigned long long acc[4] = {0ull, 0ull, 0ull, 0ull};
          __nv_bfloat162 lacc[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const uint32_t hw_[4] = {hc[i].x, hc[i].y, hc[i].z, hc[i].w};
            const uint32_t lw_[4] = {lc[i].x, lc[i].y, lc[i].z, lc[i].w};
            const __nv_bfloat162 wb = __float2bfloat162_rn(wf4[i]);
            unsigned long long wp;
            asm("mov.b64 %0, {%1, %1};" : "=l"(wp) : "r"(__float_as_uint(wf4[i])));
#pragma unroll
            for (int qq = 0; qq < 4; ++qq) {
              unsigned long long hp;
              asm("mov.b64 %0, {%1, %2};" : "=l"(hp) : "r"(hw_[qq] << 16), "r"(hw_[qq] & 0xffff0000u));
              asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(acc[qq]) : "l"(wp), "l"(hp), "l"(acc[qq]));
              const __nv_bfloat162 lv = *reinterpret_cast<const __nv_bfloat162*>(&lw_[qq]);
              lacc[qq] = i == 0 ? __hmul2(wb, lv) : __hfma2(wb, lv, lacc[qq]);
            }
          }
          uint32_t ohi[4], olo[4];
#pragma unroll
          for (int qq = 0; qq < 4; ++qq) {
            const uint32_t lw = *reinterpret_cast<const uint32_t*>(&lacc[qq]);
            unsigned long long lp;
            asm("mov.b64 %0, {%1, %2};" : "=l"(lp) : "r"(lw << 16), "r"(lw & 0xffff0000u));
            asm("add.rn.f32x2 %0, %1, %2;" : "=l"(acc[qq]) : "l"(acc[qq]), "l"(lp));
            uint32_t a0, a1;
            asm("mov.b64 {%0, %1}, %2;" : "=r"(a0), "=r"(a1) : "l"(acc[qq]));
            const float v0 = __uint_as_float(a0), v1 = __uint_as_float(a1);
            ohi[qq] = pack_bf16x2(v0, v1);
            olo[qq] = pack_bf16x2(v0 - __uint_as_float(ohi[qq] << 16), v1 - __uint_as_float(ohi[qq] & 0xffff0000u));
          }
          // A operand in TMEM: K element k of the k-block = 16-bit slot k of the row's 32 columns (slice sl = columns 8 sl .. +7)
          const uint32_t col = a_col + (uint32_t)(sl * 8 + half * 4);
          dw_tmem_st4(col, ohi[0], ohi[1], ohi[2], ohi[3]);
          dw_tmem_st4(col + 32u, olo[0], olo[1], olo[2], olo[3]);
        }
        dw_tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(bar_fa + 8 * s);
          // window buffers this warp will not read again: its next k-block (kb + 2) starts at slice 4 * (kb + 2)
          while (rel < nfill && DW_KHW * spf * (rel + 1) <= 4 * (kb + DW_GROUPS)) {
            mbar_arrive(bar_we + 8 * ((wf0 + (uint32_t)rel) & 1u));
            ++rel;
          }
        }
        rel = __shfl_sync(0xffffffffu, rel, 0);
      }
      if (lane == 0) {
        while (rel < nfill) { mbar_arrive(bar_we + 8 * ((wf0 + (uint32_t)rel) & 1u)); ++rel; }
      }
      __syncwarp();
      g0 += (uint32_t)num_kb;
      wf0 += (uint32_t)nfill;
    }
  } else if (warp == DW_WARP_TMAW) {
    // =============================== WINDOW TMA ===============================
    if (lane == 0) {
      uint32_t wf = 0, ti = 0;
      for (long long tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++ti) {
        const uint32_t par = ti & 1u;
        int4 o;
        if (p.dense) {       // receptive-field origin of the tile: no sample table, no hand-shake
          const long long mt = tile / n_tiles;
          o = make_int4((int)(mt % tiles_w) * TW - p.pw, (int)((mt / tiles_w) % tiles_h) * TH - p.ph,
                        (int)(mt / ((long long)tiles_w * tiles_h)), 0);
        } else {
          mbar_wait(bar_og + 8 * par, (ti >> 1) & 1u);
          const volatile int* ov = reinterpret_cast<const volatile int*>(&org[par]);
          o = make_int4(ov[0], ov[1], ov[2], 0);
        }
        const int cstep = 16 * spf;
        for (int f = 0; f < nfill; ++f, ++wf) {
          const uint32_t b = wf & 1u;
          mbar_wait(bar_we + 8 * b, ((wf >> 1) & 1u) ^ 1u);
          dw_expect_tx(bar_wf + 8 * b, (uint32_t)p.win_bytes);
          const uint32_t dst = win_base + b * (uint32_t)p.win_buf;
          dw_tma_4d(dst, &tm_x, bar_wf + 8 * b, f * cstep, o.x, o.y, o.z);
          dw_tma_4d(dst + (uint32_t)p.win_plane, &tm_x, bar_wf + 8 * b, p.Cin + f * cstep, o.x, o.y, o.z);
        }
      }
    }
    __syncwarp();
  } else if (warp == DW_WARP_TMAB) {
    // =============================== WEIGHT TMA ===============================
    if (lane == 0) {
      uint32_t g = 0;
      for (long long tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int n0 = (int)(tile % n_tiles) * p.BN;
        for (int kb = 0; kb < num_kb; ++kb, ++g) {
          const uint32_t s = g % NST, it = g / NST;
          mbar_wait(bar_em + 8 * s, (it & 1u) ^ 1u);
          const uint32_t stage = base + DW_OFF_STAGES + s * stage_bytes;
          dw_expect_tx(bar_fb + 8 * s, 2 * b_bytes);
          dw_tma_2d(stage, &tm_w, bar_fb + 8 * s, kb * DW_BK, n0);
          dw_tma_2d(stage + b_bytes, &tm_w, bar_fb + 8 * s, kb * DW_BK, p.Cout_pad + n0);
        }
      }
    }
    __syncwarp();
  } else if (warp == DW_WARP_MMA) {
    // =============================== MMA ISSUER ===============================
    if (lane == 0) {
      const uint32_t idesc = umma_idesc(DW_BM, p.BN);
      const uint32_t desc_hi = (uint32_t)(1024 >> 4) | (1u << 14) | (2u << 29);
      const uint32_t b0 = ((base + DW_OFF_STAGES) >> 4) & 0x3fffu, stage16 = stage_bytes >> 4, b16 = b_bytes >> 4;
      uint32_t s = 0, ph = 0, b_hi = b0, ti = 0;
      for (long long tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++ti) {
        const uint32_t buf = ti & 1u, use = ti >> 1;
        mbar_wait(bar_te + 8 * buf, (use & 1u) ^ 1u);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + buf * (uint32_t)p.BN;
        uint32_t acc = 0u;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(bar_fb + 8 * s, ph);
          mbar_wait(bar_fa + 8 * s, ph);
          tc_fence_after();
          const uint32_t a_hi = tmem_base + DW_TMEM_A0 + s * 64u, a_lo = a_hi + 32u, b_lo = b_hi + b16;
#pragma unroll
          for (uint32_t k = 0; k < DW_BK / 16; ++k) {
            dw_umma_ts(tmem_d, a_lo + 8 * k, b_hi + 2 * k, desc_hi, idesc, acc);
            dw_umma_ts(tmem_d, a_hi + 8 * k, b_lo + 2 * k, desc_hi, idesc, 1u);
            dw_umma_ts(tmem_d, a_hi + 8 * k, b_hi + 2 * k, desc_hi, idesc, 1u);
            acc = 1u;
          }
          umma_commit(bar_em + 8 * s);
          b_hi += stage16;
          if (++s == NST) { s = 0; ph ^= 1u; b_hi = b0; }
        }
        umma_commit(bar_tf + 8 * buf);
      }
    }
    __syncwarp();
  } else {
    // =============================== EPILOGUE (warps 0-3) ===============================
    const int q = warp & 3;
    uint32_t ti = 0;
    __nv_bfloat16* yb = reinterpret_cast<__nv_bfloat16*>(p.y);
    for (long long tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++ti) {
      const long long mt = tile / n_tiles;
      const int n0 = (int)(tile % n_tiles) * p.BN;
      const int tx = (int)(mt % tiles_w), ty = (int)((mt / tiles_w) % tiles_h), n = (int)(mt / ((long long)tiles_w * tiles_h));
      const uint32_t buf = ti & 1u, use = ti >> 1;
      mbar_wait(bar_tf + 8 * buf, use & 1u);
      tc_fence_after();
      const int m = q * 32 + lane;
      const int ry = m >> tw_shift;
      const int wo = tx * TW + (m & (TW - 1)), ho = ty * TH + ry;
      const bool row_ok = ry < TH && wo < p.Wo && ho < p.Ho;
      __nv_bfloat16* yp = yb + (((size_t)n * p.Ho + ho) * p.Wo + wo) * (size_t)(2 * p.Cout);
      const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16) + buf * (uint32_t)p.BN;
      for (int cb = 0; cb < p.BN; cb += 16) {
        if (n0 + cb >= p.Cout) break;              // zero-padded weight rows (warp-uniform)
        uint32_t rr[16];
        tmem_ld16(trow + (uint32_t)cb, rr);
        if (!row_ok) continue;
        const int co = n0 + cb;
        float o[16];
#pragma unroll
        for (int e = 0; e < 16; ++e) o[e] = __uint_as_float(rr[e]);
        if (p.bias) {
          if (co + 16 <= p.Cout) {
#pragma unroll
            for (int e4 = 0; e4 < 4; ++e4) {
              const float4 bv = __ldg(reinterpret_cast<const float4*>(p.bias + co) + e4);
              o[4 * e4] += bv.x; o[4 * e4 + 1] += bv.y; o[4 * e4 + 2] += bv.z; o[4 * e4 + 3] += bv.w;
            }
          } else {
#pragma unroll
            for (int e = 0; e < 16; ++e)
              if (co + e < p.Cout) o[e] += __ldg(p.bias + co + e);
          }
        }
        if (p.relu) {
#pragma unroll
          for (int e = 0; e < 16; ++e) o[e] = fmaxf(o[e], 0.f);
        }
        if (p.out_nchw) {     // plane-wise fp32 output (offset maps): for a fixed channel the lanes write consecutive pixels
          float* yf = reinterpret_cast<float*>(p.y) + ((size_t)n * p.Cout + co) * HoWo + (size_t)ho * p.Wo + wo;
#pragma unroll
          for (int e = 0; e < 16; ++e)
            if (co + e < p.Cout) yf[(size_t)e * HoWo] = o[e];
          continue;
        }
        uint32_t hw[8], lw[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          hw[e] = pack_bf16x2(o[2 * e], o[2 * e + 1]);
          lw[e] = pack_bf16x2(o[2 * e] - __uint_as_float(hw[e] << 16), o[2 * e + 1] - __uint_as_float(hw[e] & 0xffff0000u));
        }
        uint4* dh_ = reinterpret_cast<uint4*>(yp + co);
        uint4* dl_ = reinterpret_cast<uint4*>(yp + p.Cout + co);
        dh_[0] = make_uint4(hw[0], hw[1], hw[2], hw[3]); dh_[1] = make_uint4(hw[4], hw[5], hw[6], hw[7]);
        dl_[0] = make_uint4(lw[0], lw[1], lw[2], lw[3]); dl_[1] = make_uint4(lw[4], lw[5], lw[6], lw[7]);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_te + 8 * buf);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == DW_WARP_MMA) {
    tc_fence_after();
    tmem_dealloc(tmem_base, tmem_cols);
  }
}

// ----------------------------------------------------------------------------------------------
// weight pre-pack: fp32 [Cout,Cin,3,3] -> bf16 hi / lo planes [Cout_pad][K], k = (c / 16) * 144 + tap * 16 + c % 16
// ----------------------------------------------------------------------------------------------
__global__ void dcn_win_pack_kernel(const float* __restrict__ w, int Cout, int Cin, int Cout_pad, int K,
                                    uint16_t* __restrict__ hi, uint16_t* __restrict__ lo) {
  const size_t total = (size_t)Cout_pad * K;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int kk = (int)(i % (size_t)K), co = (int)(i / (size_t)K);
    const int sc = kk / (16 * DW_KHW), rem = kk - sc * 16 * DW_KHW, tap = rem >> 4, c = sc * 16 + (rem & 15);
    const float v = co < Cout ? w[((size_t)co * Cin + c) * DW_KHW + tap] : 0.f;
    const __nv_bfloat16 h = __float2bfloat16_rn(v);
    const __nv_bfloat16 l = __float2bfloat16_rn(v - __bfloat162float(h));
    hi[i] = *reinterpret_cast<const uint16_t*>(&h);
    lo[i] = *reinterpret_cast<const uint16_t*>(&l);
  }
}

static int dw_cout_pad(int Cout) { return Cout <= 32 ? 32 : (Cout + 63) / 64 * 64; }

// packing / kernel: any Cout (rows are zero-padded); the pair NHWC epilogue additionally needs Cout % 16 == 0
static bool dw_supported(int Cin, int Cout, int kh, int kw) {
  return kh == 3 && kw == 3 && Cin % 64 == 0 && Cin >= 64 && Cout >= 1;
}

typedef CUresult (*DwEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static DwEncodeFn dw_encoder() {
  static DwEncodeFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    void* q = nullptr;
    cudaDriverEntryPointQueryResult qr = cudaDriverEntryPointSymbolNotFound;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &q, cudaEnableDefault, &qr) == cudaSuccess && qr == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<DwEncodeFn>(q);
    else
      (void)cudaGetLastError();
    tried = true;
  }
  return fn;
}

}  // namespace ups

extern "C" int upsnet_dcn_packed_weight_bytes(int Cout, int Cin, int kh, int kw, size_t* bytes) {
  if (!bytes || Cout <= 0 || Cin <= 0) return UPSNET_E_BADARG;
  if (!ups::dw_supported(Cin, Cout, kh, kw)) return UPSNET_E_UNSUPPORTED;
  *bytes = (size_t)2 * ups::dw_cout_pad(Cout) * (size_t)(9 * Cin) * sizeof(uint16_t);
  return 0;
}

extern "C" int upsnet_dcn_pack_weight(const float* weight, int Cout, int Cin, int kh, int kw, void* packed, void* stream) {
  if (!weight || !packed || Cout <= 0 || Cin <= 0) return UPSNET_E_BADARG;
  if (!ups::dw_supported(Cin, Cout, kh, kw)) return UPSNET_E_UNSUPPORTED;
  const int Cout_pad = ups::dw_cout_pad(Cout), K = 9 * Cin;
  uint16_t* hi = reinterpret_cast<uint16_t*>(packed);
  uint16_t* lo = hi + (size_t)Cout_pad * K;
  const size_t total = (size_t)Cout_pad * K;
  int blocks = (int)((total + 255) / 256);
  if (blocks > ups::kNumSMs * 16) blocks = ups::kNumSMs * 16;
  ups::dcn_win_pack_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(weight, Cout, Cin, Cout_pad, K, hi, lo);
  UPS_CHECK_LAUNCH();
  return 0;
}

static int dw_launch(const void* x_pair, const float* offset, const float* mask, const void* packed,
                     const float* bias, void* y_pair, int N, int H, int W, int Cin, int Cout, int kh,
                     int kw, int pad_h, int pad_w, int dil_h, int dil_w, int epi_flags, bool dense, bool out_nchw, void* stream) {
  using namespace ups;
  if (!x_pair || (!dense && !offset) || !packed || !y_pair) return UPSNET_E_BADARG;
  if (N <= 0 || H <= 0 || W <= 0 || Cin <= 0 || Cout <= 0 || pad_h < 0 || pad_w < 0 || dil_h <= 0 || dil_w <= 0) return UPSNET_E_BADARG;
  if (!dw_supported(Cin, Cout, kh, kw)) return UPSNET_E_UNSUPPORTED;
  if (!out_nchw && (Cout % 16) != 0) return UPSNET_E_UNSUPPORTED;
  if (H >= 32767 || W >= 32767) return UPSNET_E_UNSUPPORTED;                 // outlier code packs (h, w) into 16 + 15 bits
  if ((((uintptr_t)x_pair) & 15) || (((uintptr_t)packed) & 15) || (((uintptr_t)y_pair) & 15) || (bias && (((uintptr_t)bias) & 15)))
    return UPSNET_E_UNSUPPORTED;
  DwParams p{};
  p.x = x_pair; p.offset = offset; p.mask = mask; p.bias = bias; p.y = y_pair;
  p.N = N; p.H = H; p.W = W; p.Cin = Cin; p.Cout = Cout; p.Cout_pad = dw_cout_pad(Cout);
  p.ph = pad_h; p.pw = pad_w; p.dh = dil_h; p.dw = dil_w;
  p.Ho = conv_out_size(H, pad_h, dil_h, 3, 1);
  p.Wo = conv_out_size(W, pad_w, dil_w, 3, 1);
  if (p.Ho <= 0 || p.Wo <= 0) return UPSNET_E_BADARG;
  p.relu = (epi_flags & UPSNET_EPI_RELU) ? 1 : 0;
  p.BN = p.Cout_pad % 128 == 0 ? 128 : (p.Cout_pad % 64 == 0 ? 64 : 32);
  p.dense = dense ? 1 : 0; p.out_nchw = out_nchw ? 1 : 0;
  static int sms = 0;
  if (sms == 0) {
    int dev = 0, v = kNumSMs;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
    sms = v > 0 ? v : kNumSMs;
  }
  p.tile_w = 16; p.tile_h = 8;
  auto dtiles = [&]() { return (long long)p.N * ((p.Wo + p.tile_w - 1) / p.tile_w) * ((p.Ho + p.tile_h - 1) / p.tile_h) * (p.Cout_pad / p.BN); };
  // few tiles (coarse pyramid levels): smaller pixel blocks -> more CTAs share the serial k-block chain
  if (dtiles() < sms / 2) { p.tile_w = 8; p.tile_h = 8; }
  if (dtiles() < sms / 2) { p.tile_h = 4; }
  {
    static int tile_env = -1;     // tuning hook shared with igemm_tc.cu
    if (tile_env < 0) { const char* e = getenv("UPSNET_DCN_TILE"); tile_env = e ? atoi(e) : 0; }
    if (tile_env == 168) { p.tile_w = 16; p.tile_h = 8; }
    if (tile_env == 88) { p.tile_w = 8; p.tile_h = 8; }
    if (tile_env == 84) { p.tile_w = 8; p.tile_h = 4; }
  }
  const long long num_tiles = dtiles();
  if (num_tiles <= 0) return 0;
  // dense: the window box is the tile's receptive field: 24 pixels x (tile + halo) rows x 64 channels (128-byte pixels)
  int win_h = DW_WH;
  p.win_pitch = DW_WW; p.win_plane = DW_PLANE; p.win_buf = DW_WIN_BYTES; p.stages = DW_STAGES;
  if (dense) {
    win_h = p.tile_h + 2 * dil_h;
    p.win_pitch = 24;
    if (win_h > 16 || p.tile_w + 2 * dil_w > p.win_pitch) return UPSNET_E_UNSUPPORTED;
    p.win_plane = p.win_pitch * win_h * 128;
    p.win_plane = (p.win_plane + 1023) / 1024 * 1024;     // SWIZZLE_128B destinations: 1024-byte aligned planes
    p.win_buf = 2 * p.win_plane;
  }
  if (!dense) {     // tuning hook: UPSNET_DCN_WIN_H = rows of the deformable window box (12..24)
    static int wh_env = -1;
    if (wh_env < 0) { const char* e = getenv("UPSNET_DCN_WIN_H"); wh_env = e ? atoi(e) : 0; }
    if (wh_env >= 12 && wh_env <= DW_WH) win_h = wh_env;
  }
  p.win_bytes = dense ? 2 * p.win_pitch * win_h * 128 : 2 * DW_WW * win_h * 32;
  p.win_h = win_h;
  auto smem_need = [&]() { return (size_t)DW_OFF_STAGES + (size_t)p.stages * 2 * (p.BN * 128) + 2 * (size_t)p.win_buf + 1024; };
  if (smem_need() > 227 * 1024 && p.stages > 2) p.stages = 2;
  if (smem_need() > 227 * 1024) return UPSNET_E_UNSUPPORTED;
  DwEncodeFn enc = dw_encoder();
  if (!enc) return UPSNET_E_UNSUPPORTED;
  CUtensorMap tm_x, tm_w;
  {
    const cuuint64_t dx[4] = {(cuuint64_t)(2 * Cin), (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
    const cuuint64_t sx[3] = {(cuuint64_t)(2 * Cin) * 2, (cuuint64_t)W * (2 * Cin) * 2, (cuuint64_t)H * W * (2 * Cin) * 2};
    const cuuint32_t bx[4] = {(cuuint32_t)(dense ? 64 : 16), (cuuint32_t)p.win_pitch, (cuuint32_t)win_h, 1};
    const cuuint32_t es[4] = {1, 1, 1, 1};
    if (enc(&tm_x, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(x_pair), dx, sx, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
            dense ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return UPSNET_E_UNSUPPORTED;
    const cuuint64_t dwt[2] = {(cuuint64_t)(9 * Cin), (cuuint64_t)(2 * p.Cout_pad)};
    const cuuint64_t sw[1] = {(cuuint64_t)(9 * Cin) * 2};
    const cuuint32_t bw[2] = {64, (cuuint32_t)p.BN};
    if (enc(&tm_w, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(packed), dwt, sw, bw, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return UPSNET_E_UNSUPPORTED;
  }
  const size_t smem = smem_need();
  static PerDeviceOnce configured;
  if (configured.need()) {
    UPS_CUDA(cudaFuncSetAttribute(dcn_win_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  }
  dim3 grid((unsigned)(num_tiles < sms ? num_tiles : sms));
  dcn_win_kernel<<<grid, DW_THREADS, smem, (cudaStream_t)stream>>>(tm_x, tm_w, p);
  UPS_CHECK_LAUNCH();
  return 0;
}

extern "C" int upsnet_dcn_pair_forward(const void* x_pair, const float* offset, const float* mask, const void* packed,
                                       const float* bias, void* y_pair, int N, int H, int W, int Cin, int Cout, int kh,
                                       int kw, int pad_h, int pad_w, int dil_h, int dil_w, int epi_flags, void* stream) {
  if (!offset) return UPSNET_E_BADARG;
  return dw_launch(x_pair, offset, mask, packed, bias, y_pair, N, H, W, Cin, Cout, kh, kw, pad_h, pad_w, dil_h, dil_w, epi_flags,
                   false, false, stream);
}

extern "C" int upsnet_conv3x3_pair_forward(const void* x_pair, const void* packed, const float* bias, void* y, int N, int H, int W,
                                           int Cin, int Cout, int pad_h, int pad_w, int dil_h, int dil_w, int out_layout,
                                           int epi_flags, void* stream) {
  if (out_layout != UPSNET_LAYOUT_NCHW && out_layout != UPSNET_LAYOUT_NHWC) return UPSNET_E_BADARG;
  return dw_launch(x_pair, nullptr, nullptr, packed, bias, y, N, H, W, Cin, Cout, 3, 3, pad_h, pad_w, dil_h, dil_w, epi_flags,
                   true, out_layout == UPSNET_LAYOUT_NCHW, stream);
}
