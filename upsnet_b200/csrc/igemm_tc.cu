// igemm_tc.cu -- tcgen05 implicit-GEMM convolution / deformable convolution for sm_100a.
//
// One warp-specialised kernel covers the dense k x k convolutions of the backbone / FPN / RPN /
// heads, the fully connected layers and the FUSED deformable conv v1/v2 (im2col never leaves the
// SM; reference: deformable_im2col -> 1.2 GB col buffer -> torch.mm, operators/functions/
// deform_conv.py:44-57 + operators/src/deform_conv_kernel.cu:194-242).
//
//   D[128 pixels x BN couts] (fp32, TMEM)  +=  A[128 x 64] (bf16, smem)  *  B[BN x 64]^T (bf16, smem)
//
// * A (activations, NHWC fp32 in HBM) is GATHERED by 8 producer warps: for k-block (tap, 64
//   channels) every (pixel,8-channel) item is one or -- when deformable -- four 32-byte reads
//   (the four bilinear corners; weights/offsets come from a per-tile sample table computed once
//   per (tap,pixel), reused by all channels), blended in fp32, converted to bf16 and stored with
//   one 16-byte st.shared into the K-major SWIZZLE_128B layout tcgen05 consumes.
// * B (weights) is pre-packed once to bf16 [Cout_pad][tap][Cin] and copied by the same warps.
// * One elected thread issues tcgen05.mma (kind::f16, M=128, N=BN, K=16) per 16-column slice;
//   tcgen05.commit releases the smem stage to the producers through an mbarrier; accumulators
//   live in TMEM (two buffers) and are drained with tcgen05.ld (32x32b.x16) by 4 epilogue warps
//   while the next tile's main loop is already running (persistent CTAs, static tile schedule).
// * Precision modes: BF16 (one pass) and BF16X3 (x = hi + lo split of both operands, three
//   MMAs: hi*hi + lo*hi + hi*lo; error ~2^-16 relative, i.e. fp32-grade results for the
//   "fp32 logits within 1e-3" contract at 3x the tensor work).
// Roofline: tensor pipe (flops = 2*P*Cout*Cin*kh*kw); the deformable variant is bounded by the
// LSU gather rate of the producers (4 corner reads per element) -- see DESIGN.md.
#include <cuda_bf16.h>
#include <cstdio>
#include <cstdlib>

#include "common.cuh"
#include "tc_ptx.cuh"
#include "tc_params.cuh"

namespace ups {

constexpr int TC_BM = 128;         // pixels per tile (UMMA M)
constexpr int TC_BK = 64;          // bf16 elements per k-block row (= 128 bytes, one swizzle span)
constexpr int TC_GROUP = 256;       // producer threads that fill one smem stage together (8 warps)
constexpr int TC_GROUPS = 2;        // producer groups work on alternate k-blocks (two stages in flight)
constexpr int TC_PRODUCERS = TC_GROUP * TC_GROUPS;  // 16 warps
constexpr int TC_EPILOGUE = 128;      // 4 warps: TMEM lane quadrant = warp & 3 (x TC_EPI_SPLIT column slices)
constexpr int TC_EPI_SPLIT = TC_EPILOGUE / 128;
constexpr int TC_EPI_PITCH = 36;      // floats per staged row: 32 columns + 4 pad (16-byte aligned, conflict-free)
constexpr int TC_THREADS = TC_EPILOGUE + 32 + TC_PRODUCERS;  // 21 warps
constexpr int TC_MAX_STAGES = 6;



// shared-memory carve-up (offsets from the 1024-aligned base)
struct TcSmem {
  uint32_t bars;      // full[6], empty[6], tmem_full[2], tmem_empty[2] (16 x 8 B) then tmem ptr
  uint32_t rowbase;   // long long [128]
  uint32_t epi;       // epilogue staging: 4 warps x 32 rows x 36 floats (coalesced NHWC stores)
  uint32_t table;     // deform: float4 [KHW][128] + int4 [KHW][128]; dense: int [KHW][128]
  uint32_t stages;    // 1024-aligned
  uint32_t a_bytes, b_bytes, stage_bytes, total;
};
__host__ __device__ inline TcSmem tc_smem_layout(bool deform, int KHW, int BN, int stages, bool x3) {
  TcSmem s;
  s.bars = 0;
  s.rowbase = 256;
  s.epi = s.rowbase + 2 * TC_BM * 8;     // two row-info buffers, then the epilogue staging area
  s.table = s.epi + (TC_EPILOGUE / 32) * 32 * TC_EPI_PITCH * 4;
  const uint32_t tbytes = deform ? KHW * TC_BM * 32 : KHW * TC_BM * 4;
  s.stages = (uint32_t)((s.table + tbytes + 1023) / 1024 * 1024);
  s.a_bytes = TC_BM * 128;
  s.b_bytes = BN * 128;
  s.stage_bytes = (s.a_bytes + s.b_bytes) * (x3 ? 2 : 1);
  s.total = s.stages + s.stage_bytes * stages;
  return s;
}

// Named barrier among the producer warps only (ids 1.. ; id 0 is __syncthreads)
__device__ __forceinline__ void producer_bar_sync() {
  asm volatile("bar.sync 1, %0;" ::"n"(TC_PRODUCERS) : "memory");
}

// Persistent, warp-specialised kernel.  Roles (21 warps):
//   warps 0-3   epilogue: TMEM -> registers -> bias/residual/ReLU -> global (TMEM lane quadrant = warp)
//   warp  4     MMA issuer (one elected lane), owns TMEM alloc/dealloc and barrier init
//   warps 5-20  producers, two groups of 8 warps filling alternate k-blocks (two smem stages in flight):
//               sample table, B via cp.async, A gather (loads issued first, then bf16 conversion)
// Pipelines: smem ring full[s]/empty[s] (producers <-> MMA) runs across tiles; two TMEM accumulator
// buffers tmem_full[b]/tmem_empty[b] (MMA <-> epilogue) overlap tile i's epilogue with tile i+1's
// main loop.  Tiles: id = blockIdx.x + it*gridDim.x, n-tile fastest (concurrent CTAs share the A rows in L2).
// MODE: 0 = dense (Cin % 64 == 0, NHWC), 1 = deformable, 2 = tiny Cin (stem: NCHW fp32 image, K = kh*kw*Cin
// flattened and zero-padded to a multiple of 64, element-wise gather through a per-k table)
// XM: activation storage of x -- 0 fp32, 1 bf16, 2 hi/lo bf16 pairs (NHWC with 2*Cin channels; always the 3-MMA split)
template <int MODE, int XM>
__global__ void __launch_bounds__(TC_THREADS, 1)
igemm_tc_kernel(const TcParams p) {
  constexpr bool XBF16 = XM == 1;
  constexpr bool XPAIR = XM == 2;
  constexpr bool DEFORM = MODE == 1;
  constexpr bool SMALLC = MODE == 2;
  extern __shared__ __align__(1024) uint8_t smem_dyn[];
  const uint32_t raw = smem_u32(smem_dyn);
  const uint32_t base = (raw + 1023u) & ~1023u;   // SWIZZLE_128B stage buffers need 1024-byte alignment
  uint8_t* sm = smem_dyn + (base - raw);

  const int KHW = p.kh * p.kw;
  const bool x3 = p.x3 != 0;
  const TcSmem L = tc_smem_layout(DEFORM, SMALLC ? 2 : KHW, p.BN, p.stages, x3);
  const uint32_t bar_full = base + L.bars, bar_empty = bar_full + 8 * TC_MAX_STAGES;
  const uint32_t bar_tfull = bar_empty + 8 * TC_MAX_STAGES, bar_tempty = bar_tfull + 16;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(sm + L.bars + 8 * (2 * TC_MAX_STAGES + 4));
  long long* rowbase = reinterpret_cast<long long*>(sm + L.rowbase);
  float4* tw = reinterpret_cast<float4*>(sm + L.table);
  int4* to = reinterpret_cast<int4*>(sm + L.table + (DEFORM ? KHW * TC_BM * 16 : 0));
  int* ti = reinterpret_cast<int*>(sm + L.table);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int HoWo = p.Ho * p.Wo;
  const long long Ptot = (long long)p.N * HoWo;
  const int cchunks = SMALLC ? 1 : p.Cin / TC_BK;
  const int Kreal = KHW * p.Cin;
  const int Kp = (Kreal + TC_BK - 1) / TC_BK * TC_BK;      // == Kreal unless SMALLC
  const int num_kb = Kp / TC_BK;
  const int n_tiles = p.Cout_pad / p.BN;
  // M-tile = 128 output pixels.  Dense / stem modes: 128 consecutive pixels of the flattened (n, ho, wo) index.
  // Deformable mode: a 16 x 8 pixel BLOCK of one image -- its nine taps x four corners then revisit ~(16+3) x (8+3)
  // input pixels per channel chunk (27 KB: L1-resident) instead of four 131-pixel row segments.
  // (small maps use 8x8 or 8x4 blocks -- rows past the block stay empty and cost no gather work -- so that more CTAs
  // share the serial k-block chain; the launcher picks the block)
  const int TW = DEFORM ? p.tile_w : 16, TH = DEFORM ? p.tile_h : 8;
  const int tw_shift = TW == 16 ? 4 : 3;
  const int tiles_w = (p.Wo + TW - 1) / TW, tiles_h = (p.Ho + TH - 1) / TH;
  const long long m_tiles = DEFORM ? (long long)p.N * tiles_w * tiles_h : (Ptot + TC_BM - 1) / TC_BM;
  const long long num_tiles = m_tiles * n_tiles;
  // flattened output pixel of row r of M-tile mt, or -1 when the row lies outside the tensor
  auto tile_pixel = [&](long long mt, int r) -> long long {
    if (DEFORM) {
      const int tx = (int)(mt % tiles_w), ty = (int)((mt / tiles_w) % tiles_h), n = (int)(mt / ((long long)tiles_w * tiles_h));
      const int ry = r >> tw_shift;
      const int wo = tx * TW + (r & (TW - 1)), ho = ty * TH + ry;
      return (ry < TH && wo < p.Wo && ho < p.Ho) ? ((long long)n * p.Ho + ho) * p.Wo + wo : -1ll;
    }
    const long long pg = mt * TC_BM + r;
    return pg < Ptot ? pg : -1ll;
  };
  uint32_t tmem_cols = 32;
  while ((int)tmem_cols < 2 * p.BN) tmem_cols <<= 1;

  // ---------------- one-time setup ----------------
  if (warp == TC_EPILOGUE / 32) {
    if (lane == 0) {
      for (int s = 0; s < p.stages; ++s) {
        mbar_init(bar_full + 8 * s, TC_GROUP / 32);
        mbar_init(bar_empty + 8 * s, 1);
      }
      for (int b = 0; b < 2; ++b) {
        mbar_init(bar_tfull + 8 * b, 1);
        mbar_init(bar_tempty + 8 * b, TC_EPILOGUE / 32);   // one arrive per epilogue warp
      }
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc(smem_u32(tmem_ptr_smem), tmem_cols);
  }
  if (SMALLC) {   // per-k table: k -> (dy, dx, channel) packed, -1 for the zero padding of K
    for (int k = tid; k < Kp; k += TC_THREADS) {
      int v = -1;
      if (k < Kreal) {
        const int tap = k / p.Cin, c = k - tap * p.Cin;
        const int ki = tap / p.kw, kj = tap - ki * p.kw;
        v = ((ki * p.dh) << 16) | ((kj * p.dw) << 8) | c;
      }
      ti[k] = v;
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp > TC_EPILOGUE / 32) {
    // =============================== PRODUCERS ===============================
    const int pt = tid - (TC_EPILOGUE + 32);  // 0..511
    const int group = pt / TC_GROUP;      // which alternate k-blocks this thread fills
    const int gt = pt - group * TC_GROUP; // 0..255 inside the group
    const int j = gt & 7;                 // 16-byte chunk (8 channels) inside the 128-byte row
    const int r_first = gt >> 3;          // 32 rows per pass
    uint32_t g0 = 0;                      // ring position of this tile's first k-block
    uint32_t tile_it = 0;
    // bf16 dense mode: both operands travel by cp.async, so the producer keeps ONE k-block outstanding and
    // signals the previous one only after issuing the next (two k-blocks of loads in flight per group).
    const bool deferred = (MODE == 0) && (XBF16 || XPAIR) && p.stages >= 3;
    int pend_s = -1;
    for (long long tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const long long mt = tile / n_tiles;
      const int n0 = (int)(tile % n_tiles) * p.BN;
      // Per-tile row info (image, top-left input coordinate of the receptive field) is double-buffered, so ONE
      // producer barrier per tile suffices for the dense / stem modes; the deformable sample table is a single
      // buffer and needs the extra barrier before it is overwritten.
      long long* rowinfo = rowbase + (tile_it & 1) * TC_BM;
      if (DEFORM) producer_bar_sync();   // every producer is done with the previous tile's sample table
      for (int r = pt; r < TC_BM; r += TC_PRODUCERS) {
        const long long pg = tile_pixel(mt, r);
        long long v = -1;
        if (pg >= 0) {
          const int n = (int)(pg / HoWo), pp = (int)(pg - (long long)n * HoWo);
          const int ho = pp / p.Wo, wo = pp - ho * p.Wo;
          if (DEFORM) v = (long long)n * p.H * p.W * (long long)p.Cin * (XPAIR ? 2 : 1);
          else v = ((long long)n << 40) | ((long long)(ho * p.sh - p.ph + (1 << 19)) << 20) | (long long)(wo * p.sw - p.pw + (1 << 19));
        }
        rowinfo[r] = v;
      }
      // deformable: per-tile sample table (channel independent), one entry per (tap, pixel)
      for (int e = pt; e < (DEFORM ? KHW * TC_BM : 0); e += TC_PRODUCERS) {
        const int tap = e / TC_BM, r = e - tap * TC_BM;
        const long long pg = tile_pixel(mt, r);
        const int ki = tap / p.kw, kj = tap - ki * p.kw;
        {
          float4 wv = make_float4(0.f, 0.f, 0.f, 0.f);
          int4 ov = make_int4(0, 0, 0, 0);
          if (pg >= 0) {
            const int n = (int)(pg / HoWo), pp = (int)(pg - (long long)n * HoWo);
            const int ho = pp / p.Wo, wo = pp - ho * p.Wo;
            const float* offp = p.offset + ((size_t)n * 2 * KHW + 2 * tap) * HoWo + pp;
            const float oh = __ldg(offp), ow = __ldg(offp + HoWo);
            const float h = (float)(ho * p.sh - p.ph + ki * p.dh) + oh;
            const float w = (float)(wo * p.sw - p.pw + kj * p.dw) + ow;
            if (h > -1.f && w > -1.f && h < (float)p.H && w < (float)p.W) {  // deform_conv_kernel.cu:229
              const int hl = (int)floorf(h), wl = (int)floorf(w), hh = hl + 1, wh = wl + 1;
              const float lh = h - hl, lw = w - wl, ch = 1.f - lh, cw = 1.f - lw;
              const bool t_ok = hl >= 0, b_ok = hh <= p.H - 1, l_ok = wl >= 0, r_ok = wh <= p.W - 1;
              float m = 1.f;
              if (p.mask) m = __ldg(p.mask + ((size_t)n * KHW + tap) * HoWo + pp);
              wv.x = (t_ok && l_ok) ? ch * cw * m : 0.f;
              wv.y = (t_ok && r_ok) ? ch * lw * m : 0.f;
              wv.z = (b_ok && l_ok) ? lh * cw * m : 0.f;
              wv.w = (b_ok && r_ok) ? lh * lw * m : 0.f;
              ov.x = (t_ok && l_ok) ? hl * p.W + wl : 0;
              ov.y = (t_ok && r_ok) ? hl * p.W + wh : 0;
              ov.z = (b_ok && l_ok) ? hh * p.W + wl : 0;
              ov.w = (b_ok && r_ok) ? hh * p.W + wh : 0;
            }
          }
          if (XPAIR) {   // pair gather: element offsets in the 2*Cin-channel tensor, fp32 weights
            ov.x *= 2 * p.Cin; ov.y *= 2 * p.Cin; ov.z *= 2 * p.Cin; ov.w *= 2 * p.Cin;
            tw[e] = wv;
          } else if (XBF16) {   // bf16 gather blends in packed bf16x2: weights replicated into both halves, offsets in elements
            ov.x *= p.Cin; ov.y *= p.Cin; ov.z *= p.Cin; ov.w *= p.Cin;
            uint4 wp;
            wp.x = pack_bf16x2(wv.x, wv.x); wp.y = pack_bf16x2(wv.y, wv.y);
            wp.z = pack_bf16x2(wv.z, wv.z); wp.w = pack_bf16x2(wv.w, wv.w);
            reinterpret_cast<uint4*>(tw)[e] = wp;
          } else {
            tw[e] = wv;
          }
          to[e] = ov;
        }
      }
      producer_bar_sync();   // row info (and sample table) visible to all producers
      ++tile_it;

      for (int kb = (int)((uint32_t)(group - (int)g0) & 1u); kb < num_kb; kb += TC_GROUPS) {  // ring parity == group
        const uint32_t g = g0 + (uint32_t)kb;
        const uint32_t s = g % (uint32_t)p.stages, it = g / (uint32_t)p.stages;
        mbar_wait(bar_empty + 8 * s, (it & 1u) ^ 1u);
        uint8_t* stage = sm + L.stages + (size_t)s * L.stage_bytes;
        uint8_t* a_hi = stage;
        uint8_t* b_hi = stage + L.a_bytes;
        uint8_t* a_lo = stage + L.a_bytes + L.b_bytes;
        uint8_t* b_lo = a_lo + L.a_bytes;
        // k-block order: dense = tap-major (matches the packed weight columns); deformable = CHANNEL-CHUNK-major, so
        // that the nine taps x four bilinear corners of one 64-channel chunk -- which revisit the same few hundred
        // 128-byte lines of the input -- run back to back and hit in L1 instead of going to L2 36 times.
        const int tap = DEFORM ? kb % KHW : kb / cchunks;
        const int cck = DEFORM ? kb / KHW : kb - tap * cchunks;
        const int c0 = cck * TC_BK + j * 8;
        const size_t kcol = (size_t)tap * p.Cin + (size_t)cck * TC_BK;   // == kb * 64 for the dense modes
        const int tki = tap / p.kw, tdy = tki * p.dh, tdx = (tap - tki * p.kw) * p.dw;   // dense: tap displacement
        // ---- B: BN rows x 8 chunks of packed bf16 weights, cp.async straight into the swizzled stage
        //      (no registers, overlaps the A gather below) ----
        for (int r = r_first; r < p.BN; r += 32) {
          const size_t gi = (size_t)(n0 + r) * Kp + (SMALLC ? (size_t)kb * TC_BK : kcol) + j * 8;
          const uint32_t soff = (uint32_t)r * 128u + (uint32_t)((j ^ (r & 7)) << 4);
          cp_async16(smem_u32(b_hi + soff), p.w_hi + gi);
          if (x3) cp_async16(smem_u32(b_lo + soff), p.w_lo + gi);
        }
        cp_async_commit();
        // ---- A: gather 128 rows x 8 chunks ----
        if (SMALLC) {
          // tiny Cin (stem): every k of the flattened (ky,kx,c) axis is an independent scalar read of the NCHW image
          const float* xf = reinterpret_cast<const float*>(p.x);
#pragma unroll 1
          for (int pass = 0; pass < TC_BM / 32; ++pass) {
            const int r = r_first + pass * 32;
            const long long rb = rowinfo[r];
            float v[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) v[q] = 0.f;
            if (rb >= 0) {
              const int n = (int)(rb >> 40), h0 = (int)((rb >> 20) & 0xfffff) - (1 << 19), w0 = (int)(rb & 0xfffff) - (1 << 19);
              const float* xn = xf + (size_t)n * p.Cin * p.H * p.W;
#pragma unroll
              for (int q = 0; q < 8; ++q) {
                const int e = ti[kb * TC_BK + j * 8 + q];
                if (e >= 0) {
                  const int hi = h0 + (e >> 16), wi = w0 + ((e >> 8) & 0xff), c = e & 0xff;
                  if (hi >= 0 && hi < p.H && wi >= 0 && wi < p.W) v[q] = __ldg(xn + ((size_t)c * p.H + hi) * p.W + wi);
                }
              }
            }
            const uint32_t soff = (uint32_t)r * 128u + (uint32_t)((j ^ (r & 7)) << 4);
            uint4 hi4;
            hi4.x = pack_bf16x2(v[0], v[1]); hi4.y = pack_bf16x2(v[2], v[3]);
            hi4.z = pack_bf16x2(v[4], v[5]); hi4.w = pack_bf16x2(v[6], v[7]);
            *reinterpret_cast<uint4*>(a_hi + soff) = hi4;
            if (x3) {
              uint4 lo;
              lo.x = pack_bf16x2(v[0] - bf16_round(v[0]), v[1] - bf16_round(v[1]));
              lo.y = pack_bf16x2(v[2] - bf16_round(v[2]), v[3] - bf16_round(v[3]));
              lo.z = pack_bf16x2(v[4] - bf16_round(v[4]), v[5] - bf16_round(v[5]));
              lo.w = pack_bf16x2(v[6] - bf16_round(v[6]), v[7] - bf16_round(v[7]));
              *reinterpret_cast<uint4*>(a_lo + soff) = lo;
            }
          }
        } else if (!DEFORM && XPAIR) {
          // dense, hi/lo pair activations: the hi and lo 128-byte rows ARE the smem rows of the two A tiles -> two cp.async
          // of 16 B per (row, chunk) straight into the swizzled stage (zero-fill for padding / out-of-range rows)
          const __nv_bfloat16* xh = reinterpret_cast<const __nv_bfloat16*>(p.x);
#pragma unroll
          for (int pass = 0; pass < TC_BM / 32; ++pass) {
            const int r = r_first + pass * 32;
            const long long rb = rowinfo[r];
            const int hi = (int)((rb >> 20) & 0xfffff) - (1 << 19) + tdy, wi = (int)(rb & 0xfffff) - (1 << 19) + tdx;
            const bool ok = rb >= 0 && hi >= 0 && hi < p.H && wi >= 0 && wi < p.W;
            const __nv_bfloat16* src = ok ? xh + (((size_t)(rb >> 40) * p.H + hi) * p.W + wi) * (size_t)(2 * p.Cin) + c0 : xh;
            const uint32_t soff = (uint32_t)r * 128u + (uint32_t)((j ^ (r & 7)) << 4);
            cp_async16_zfill(smem_u32(a_hi + soff), src, ok ? 16u : 0u);
            cp_async16_zfill(smem_u32(a_lo + soff), ok ? src + p.Cin : xh, ok ? 16u : 0u);
          }
        } else if (DEFORM && XPAIR) {
          // deformable, hi/lo pair activations: 8 lanes x 16 B cover a row's 64 channels of one plane; per corner one hi and
          // one lo load.  The gather is ISSUE-bound (ncu: 2.4 IPC, tensor pipe 18 %), so the blend is written for instruction
          // count: hi plane = packed fp32x2 FMAs (FFMA2: two channels per instruction, exact fp32 products of the bf16 values),
          // lo plane = packed bf16x2 HFMA2 with bf16-rounded weights (the lo plane is 2^-9 of the value, its blend only needs
          // 2^-9 relative accuracy -> 2^-18 overall); the two sums are added in fp32 and the result is split again.
          const __nv_bfloat16* xh = reinterpret_cast<const __nv_bfloat16*>(p.x);
#pragma unroll 1
          for (int pass = 0; pass < TC_BM / 32; ++pass) {
            const int r = r_first + pass * 32;
            const long long rb = rowinfo[r];
            unsigned long long acc[4] = {0ull, 0ull, 0ull, 0ull};      // fp32x2: channels (2q, 2q+1)
            if (rb >= 0) {
              const __nv_bfloat16* xb = xh + rb + c0;
              const float4 wv = tw[tap * TC_BM + r];
              const int4 ov = to[tap * TC_BM + r];
              const uint4 ha = __ldg(reinterpret_cast<const uint4*>(xb + ov.x)), la = __ldg(reinterpret_cast<const uint4*>(xb + ov.x + p.Cin));
              const uint4 hb = __ldg(reinterpret_cast<const uint4*>(xb + ov.y)), lb = __ldg(reinterpret_cast<const uint4*>(xb + ov.y + p.Cin));
              const uint4 hd = __ldg(reinterpret_cast<const uint4*>(xb + ov.z)), ld = __ldg(reinterpret_cast<const uint4*>(xb + ov.z + p.Cin));
              const uint4 he = __ldg(reinterpret_cast<const uint4*>(xb + ov.w)), le = __ldg(reinterpret_cast<const uint4*>(xb + ov.w + p.Cin));
              const uint32_t H[4][4] = {{ha.x, ha.y, ha.z, ha.w}, {hb.x, hb.y, hb.z, hb.w}, {hd.x, hd.y, hd.z, hd.w}, {he.x, he.y, he.z, he.w}};
              const uint32_t Lo[4][4] = {{la.x, la.y, la.z, la.w}, {lb.x, lb.y, lb.z, lb.w}, {ld.x, ld.y, ld.z, ld.w}, {le.x, le.y, le.z, le.w}};
              const float wf[4] = {wv.x, wv.y, wv.z, wv.w};
              __nv_bfloat162 lacc[4];
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const __nv_bfloat162 wb = __float2bfloat162_rn(wf[i]);
                unsigned long long wp;
                asm("mov.b64 %0, {%1, %1};" : "=l"(wp) : "r"(__float_as_uint(wf[i])));
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                  unsigned long long hp;
                  asm("mov.b64 %0, {%1, %2};" : "=l"(hp) : "r"(H[i][q] << 16), "r"(H[i][q] & 0xffff0000u));
                  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(acc[q]) : "l"(wp), "l"(hp), "l"(acc[q]));
                  const __nv_bfloat162 lv = *reinterpret_cast<const __nv_bfloat162*>(&Lo[i][q]);
                  lacc[q] = i == 0 ? __hmul2(wb, lv) : __hfma2(wb, lv, lacc[q]);
                }
              }
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                const uint32_t lw = *reinterpret_cast<const uint32_t*>(&lacc[q]);
                unsigned long long lp;
                asm("mov.b64 %0, {%1, %2};" : "=l"(lp) : "r"(lw << 16), "r"(lw & 0xffff0000u));
                asm("add.rn.f32x2 %0, %1, %2;" : "=l"(acc[q]) : "l"(acc[q]), "l"(lp));
              }
            }
            const uint32_t soff = (uint32_t)r * 128u + (uint32_t)((j ^ (r & 7)) << 4);
            uint32_t hw[4], lw4[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              uint32_t a0, a1;
              asm("mov.b64 {%0, %1}, %2;" : "=r"(a0), "=r"(a1) : "l"(acc[q]));
              const float v0 = __uint_as_float(a0), v1 = __uint_as_float(a1);
              hw[q] = pack_bf16x2(v0, v1);
              lw4[q] = pack_bf16x2(v0 - __uint_as_float(hw[q] << 16), v1 - __uint_as_float(hw[q] & 0xffff0000u));
            }
            *reinterpret_cast<uint4*>(a_hi + soff) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
            *reinterpret_cast<uint4*>(a_lo + soff) = make_uint4(lw4[0], lw4[1], lw4[2], lw4[3]);
          }
        } else if (!DEFORM && XBF16) {
          // dense, bf16 activations: the 128-byte row IS the smem row -> cp.async 16 B per (row, chunk)
          // straight into the swizzled stage (zero-fill for padding / out-of-range rows), no registers.
          const __nv_bfloat16* xh = reinterpret_cast<const __nv_bfloat16*>(p.x);
#pragma unroll
          for (int pass = 0; pass < TC_BM / 32; ++pass) {
            const int r = r_first + pass * 32;
            const long long rb = rowinfo[r];
            const int hi = (int)((rb >> 20) & 0xfffff) - (1 << 19) + tdy, wi = (int)(rb & 0xfffff) - (1 << 19) + tdx;
            const bool ok = rb >= 0 && hi >= 0 && hi < p.H && wi >= 0 && wi < p.W;
            const __nv_bfloat16* src = ok ? xh + (((size_t)(rb >> 40) * p.H + hi) * p.W + wi) * p.Cin + c0 : xh;
            const uint32_t soff = (uint32_t)r * 128u + (uint32_t)((j ^ (r & 7)) << 4);
            cp_async16_zfill(smem_u32(a_hi + soff), src, ok ? 16u : 0u);
          }
        } else if (!DEFORM) {
          const float* xf = reinterpret_cast<const float*>(p.x);
          // dense: a row's 64 fp32 channels (256 B) are read by 16 consecutive lanes, 16 B each, so every
          // warp-wide LDG.128 covers two fully used 256-byte spans; all eight loads of a thread are
          // issued before the first conversion (memory-level parallelism); each lane then stores 4 bf16
          // (8 B) into its half of the swizzled 16-byte chunk.
          const int l16 = gt & 15;            // 4-channel group inside the 64-channel row
          const int rr0 = gt >> 4;            // 16 rows per pass
          const int cg = c0 - j * 8 + l16 * 4;  // first channel of this lane's group
          float4 qv[TC_BM / 16];
#pragma unroll
          for (int pass = 0; pass < TC_BM / 16; ++pass) {
            const int r = rr0 + pass * 16;
            const long long rb = rowinfo[r];
            const int hi = (int)((rb >> 20) & 0xfffff) - (1 << 19) + tdy, wi = (int)(rb & 0xfffff) - (1 << 19) + tdx;
            qv[pass] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (rb >= 0 && hi >= 0 && hi < p.H && wi >= 0 && wi < p.W)
              qv[pass] = __ldg(reinterpret_cast<const float4*>(xf + (((size_t)(rb >> 40) * p.H + hi) * p.W + wi) * p.Cin + cg));
          }
#pragma unroll
          for (int pass = 0; pass < TC_BM / 16; ++pass) {
            const int r = rr0 + pass * 16;
            const uint32_t soff = (uint32_t)r * 128u + (uint32_t)(((l16 >> 1) ^ (r & 7)) << 4) + (uint32_t)((l16 & 1) << 3);
            uint2 hi;
            hi.x = pack_bf16x2(qv[pass].x, qv[pass].y); hi.y = pack_bf16x2(qv[pass].z, qv[pass].w);
            *reinterpret_cast<uint2*>(a_hi + soff) = hi;
            if (x3) {
              uint2 lo;
              lo.x = pack_bf16x2(qv[pass].x - bf16_round(qv[pass].x), qv[pass].y - bf16_round(qv[pass].y));
              lo.y = pack_bf16x2(qv[pass].z - bf16_round(qv[pass].z), qv[pass].w - bf16_round(qv[pass].w));
              *reinterpret_cast<uint2*>(a_lo + soff) = lo;
            }
          }
        } else if (XBF16) {
          // deformable, bf16 activations: 4 lanes x 32 B cover a row's 64 channels (two 16-byte loads per corner,
          // one address computation); the four corners are blended in packed bf16x2 (HFMA2.BF16: 4 ops per 8
          // channels and corner pair instead of 8 unpack + 8 FMA + pack) with the tile's sample table, whose
          // weights are stored as replicated bf16 pairs.  This kernel is issue-bound (ncu: 2.3 IPC, L2 16 %), so
          // instructions per gathered element are what matters.
          const __nv_bfloat16* xh = reinterpret_cast<const __nv_bfloat16*>(p.x);
          const int j2 = gt & 3, rr0 = gt >> 2;              // 64 rows per pass
          const int cp0 = c0 - j * 8 + j2 * 16;              // first channel of this lane's 16-channel slice
          const uint4* twp = reinterpret_cast<const uint4*>(tw);
#pragma unroll 1
          for (int pass = 0; pass < TC_BM / 64; ++pass) {
            const int r = rr0 + pass * 64;
            const long long rb = rowinfo[r];
            uint4 o0 = make_uint4(0u, 0u, 0u, 0u), o1 = o0;
            if (rb >= 0) {
              const __nv_bfloat16* xb = xh + rb + cp0;
              const uint4 wv = twp[tap * TC_BM + r];
              const int4 ov = to[tap * TC_BM + r];
              // one 256-bit load per corner (LDG.E.256): 4 lanes cover a row's full 128-byte line, so the L1
              // wavefront count stays that of the 8-lane x 16-byte mapping while the address work is halved
              uint4 a0, a1, b0, b1, d0, d1, e0, e1;
              ldg256(xb + ov.x, a0, a1);
              ldg256(xb + ov.y, b0, b1);
              ldg256(xb + ov.z, d0, d1);
              ldg256(xb + ov.w, e0, e1);
              o0.x = bf2_blend(wv, a0.x, b0.x, d0.x, e0.x); o0.y = bf2_blend(wv, a0.y, b0.y, d0.y, e0.y);
              o0.z = bf2_blend(wv, a0.z, b0.z, d0.z, e0.z); o0.w = bf2_blend(wv, a0.w, b0.w, d0.w, e0.w);
              o1.x = bf2_blend(wv, a1.x, b1.x, d1.x, e1.x); o1.y = bf2_blend(wv, a1.y, b1.y, d1.y, e1.y);
              o1.z = bf2_blend(wv, a1.z, b1.z, d1.z, e1.z); o1.w = bf2_blend(wv, a1.w, b1.w, d1.w, e1.w);
            }
            const uint32_t rbase = (uint32_t)r * 128u, rx = (uint32_t)(r & 7);
            *reinterpret_cast<uint4*>(a_hi + rbase + ((((uint32_t)j2 * 2u) ^ rx) << 4)) = o0;
            *reinterpret_cast<uint4*>(a_hi + rbase + ((((uint32_t)j2 * 2u + 1u) ^ rx) << 4)) = o1;
          }
        } else {
          const float* xf = reinterpret_cast<const float*>(p.x);
          // deformable: same coalesced lane mapping; per (row, 4-channel group) four 16-byte corner reads,
          // blended in fp32 with the tile's sample table (weights already carry validity and the v2 mask).
          const int l16 = gt & 15;
          const int rr0 = gt >> 4;
          const int cg = c0 - j * 8 + l16 * 4;
#pragma unroll 2
          for (int pass = 0; pass < TC_BM / 16; ++pass) {
            const int r = rr0 + pass * 16;
            const long long rb = rowinfo[r];
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (rb >= 0) {
              const float* xb = xf + rb + cg;
              const float4 wv = tw[tap * TC_BM + r];
              const int4 ov = to[tap * TC_BM + r];
              const float4 a0 = __ldg(reinterpret_cast<const float4*>(xb + (size_t)ov.x * p.Cin));
              const float4 b0 = __ldg(reinterpret_cast<const float4*>(xb + (size_t)ov.y * p.Cin));
              const float4 d0 = __ldg(reinterpret_cast<const float4*>(xb + (size_t)ov.z * p.Cin));
              const float4 e0 = __ldg(reinterpret_cast<const float4*>(xb + (size_t)ov.w * p.Cin));
              v.x = wv.x * a0.x + wv.y * b0.x + wv.z * d0.x + wv.w * e0.x;
              v.y = wv.x * a0.y + wv.y * b0.y + wv.z * d0.y + wv.w * e0.y;
              v.z = wv.x * a0.z + wv.y * b0.z + wv.z * d0.z + wv.w * e0.z;
              v.w = wv.x * a0.w + wv.y * b0.w + wv.z * d0.w + wv.w * e0.w;
            }
            const uint32_t soff = (uint32_t)r * 128u + (uint32_t)(((l16 >> 1) ^ (r & 7)) << 4) + (uint32_t)((l16 & 1) << 3);
            uint2 hi;
            hi.x = pack_bf16x2(v.x, v.y); hi.y = pack_bf16x2(v.z, v.w);
            *reinterpret_cast<uint2*>(a_hi + soff) = hi;
            if (x3) {
              uint2 lo;
              lo.x = pack_bf16x2(v.x - bf16_round(v.x), v.y - bf16_round(v.y));
              lo.y = pack_bf16x2(v.z - bf16_round(v.z), v.w - bf16_round(v.w));
              *reinterpret_cast<uint2*>(a_lo + soff) = lo;
            }
          }
        }
        cp_async_commit();
        if (deferred) {
          if (pend_s >= 0) {
            cp_async_wait_but2();   // everything except this k-block's two commit groups has landed
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_full + 8 * pend_s);
          }
          pend_s = (int)s;
        } else {
          cp_async_wait_all();
          fence_proxy_async();  // generic-proxy stores -> visible to the tensor core (async proxy)
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_full + 8 * s);
        }
      }
      g0 += (uint32_t)num_kb;
    }
    if (pend_s >= 0) {
      cp_async_wait_all();
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_full + 8 * pend_s);
    }
  } else if (warp == TC_EPILOGUE / 32) {
    // =============================== MMA ISSUER ===============================
    if (lane == 0) {
      const uint32_t idesc = umma_idesc(TC_BM, p.BN);
      // descriptors: constant high word (SBO 1024 B, version 1, SWIZZLE_128B) + running low word (address >> 4);
      // ring position / phase are running counters (no division on this single thread's per-k-block path)
      const uint32_t desc_hi = (uint32_t)(1024 >> 4) | (1u << 14) | (2u << 29);
      const uint32_t a_lo0 = ((base + L.stages) >> 4) & 0x3fffu, stage16 = L.stage_bytes >> 4;
      const uint32_t a16 = L.a_bytes >> 4, b16 = L.b_bytes >> 4;
      uint32_t s = 0, ph = 0, a_hi = a_lo0, ti_local = 0;
      for (long long tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++ti_local) {
        const uint32_t buf = ti_local & 1u, use = ti_local >> 1;
        mbar_wait(bar_tempty + 8 * buf, (use & 1u) ^ 1u);   // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + buf * (uint32_t)p.BN;
        uint32_t acc = 0u;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(bar_full + 8 * s, ph);
          tc_fence_after();
          const uint32_t b_hi = a_hi + a16, a_lo = b_hi + b16, b_lo = a_lo + a16;
#pragma unroll
          for (uint32_t k = 0; k < TC_BK / 16; ++k) {       // 16 bf16 = 32 bytes = 2 descriptor units
            if (x3) {
              umma_bf16_lohi(tmem_d, a_lo + 2 * k, b_hi + 2 * k, desc_hi, idesc, acc);
              umma_bf16_lohi(tmem_d, a_hi + 2 * k, b_lo + 2 * k, desc_hi, idesc, 1u);
              umma_bf16_lohi(tmem_d, a_hi + 2 * k, b_hi + 2 * k, desc_hi, idesc, 1u);
            } else {
              umma_bf16_lohi(tmem_d, a_hi + 2 * k, b_hi + 2 * k, desc_hi, idesc, acc);
            }
            acc = 1u;
          }
          umma_commit(bar_empty + 8 * s);   // frees the stage once the MMAs above have read it
          a_hi += stage16;
          if (++s == (uint32_t)p.stages) { s = 0; ph ^= 1u; a_hi = a_lo0; }
        }
        umma_commit(bar_tfull + 8 * buf);   // accumulator complete -> epilogue
      }
    }
    __syncwarp();
  } else {
    // =============================== EPILOGUE (warps 0-3) ===============================
    const int q = warp & 3;        // TMEM lane quadrant
    const int half = warp >> 2;    // which slice of the BN accumulator columns (TC_EPI_SPLIT slices)
    const bool vec_ptrs_ok = (((uintptr_t)p.y) & 15) == 0 && (!p.residual || (((uintptr_t)p.residual) & 15) == 0);
    uint32_t ti_local = 0;
    for (long long tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++ti_local) {
      const long long mt = tile / n_tiles;
      const int n0 = (int)(tile % n_tiles) * p.BN;
      const uint32_t buf = ti_local & 1u, use = ti_local >> 1;
      mbar_wait(bar_tfull + 8 * buf, use & 1u);
      tc_fence_after();
      const int m = q * 32 + lane;
      const long long pg = tile_pixel(mt, m);
      const bool row_ok = pg >= 0;
      const int n_img = row_ok ? (int)(pg / HoWo) : 0;
      const int pp = row_ok ? (int)(pg - (long long)n_img * HoWo) : 0;
      const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16) + buf * (uint32_t)p.BN;
      if (p.out_nhwc && (p.Cout & 7) == 0 && vec_ptrs_ok) {
        // ---- NHWC: stage 32 rows x 32 columns (fp32, +bias) per warp in shared memory, then write them out
        //      row-wise: 4 (bf16) or 8 (fp32) lanes cover one row's 64 / 128 contiguous bytes, so every store
        //      -- and every residual read -- is made of fully used 32-byte sectors instead of one 16-byte
        //      fragment per 512-byte-strided row. ----
        float* st = reinterpret_cast<float*>(sm + L.epi) + (size_t)warp * 32 * TC_EPI_PITCH;
        const bool y16 = p.y_bf16 || p.y_pair;          // 16-bit storage: 8 columns (16 B) per lane
        const size_t ypitch = p.y_pair ? 2 * (size_t)p.Cout : (size_t)p.Cout;   // elements per stored pixel
        const int lpr = y16 ? 4 : 8;                    // lanes per row in the write-out phase
        const int rpi = 32 / lpr;                       // rows per iteration
        const int sub = lane % lpr, rsub = lane / lpr;
        for (int cb = 0; cb < p.BN; cb += 32) {
          if (n0 + cb >= p.Cout) break;                 // zero-padded weight rows beyond Cout
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            uint32_t rr[16];
            tmem_ld16(trow + (uint32_t)(cb + c * 16), rr);   // warp-collective
            const int co0 = n0 + cb + c * 16;
            float4* dst = reinterpret_cast<float4*>(st + lane * TC_EPI_PITCH + c * 16);
#pragma unroll
            for (int g4 = 0; g4 < 4; ++g4) {
              float4 o = make_float4(__uint_as_float(rr[g4 * 4]), __uint_as_float(rr[g4 * 4 + 1]),
                                     __uint_as_float(rr[g4 * 4 + 2]), __uint_as_float(rr[g4 * 4 + 3]));
              if (p.bias && co0 + g4 * 4 + 3 < p.Cout) {
                o.x += __ldg(p.bias + co0 + g4 * 4); o.y += __ldg(p.bias + co0 + g4 * 4 + 1);
                o.z += __ldg(p.bias + co0 + g4 * 4 + 2); o.w += __ldg(p.bias + co0 + g4 * 4 + 3);
              }
              dst[g4] = o;
            }
          }
          __syncwarp();
          const int ce = y16 ? sub * 8 : sub * 4;        // first staged column of this lane
          const int co = n0 + cb + ce;
          if (co < p.Cout) {
            for (int it = 0; it < lpr; ++it) {
              const int row = it * rpi + rsub;
              const long long pgr = tile_pixel(mt, q * 32 + row);
              if (pgr < 0) continue;
              size_t ridx = (size_t)pgr * ypitch + co;
              if (p.residual && p.res_up2) {
                const int ni = (int)(pgr / HoWo), ppr = (int)(pgr - (long long)ni * HoWo);
                const int ho = ppr / p.Wo, wo = ppr - ho * p.Wo;
                ridx = (((size_t)ni * (p.Ho >> 1) + (ho >> 1)) * (p.Wo >> 1) + (wo >> 1)) * ypitch + co;
              }
              const float* src = st + row * TC_EPI_PITCH + ce;
              if (p.y_pair) {
                // hi/lo pair output: residual = hi + lo (exact in fp32), result split into bf16(o) and bf16(o - hi)
                const float4 a = *reinterpret_cast<const float4*>(src), b = *reinterpret_cast<const float4*>(src + 4);
                float o[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
                if (p.residual) {
                  const __nv_bfloat16* rp = reinterpret_cast<const __nv_bfloat16*>(p.residual) + ridx;
                  const uint4 rh = __ldg(reinterpret_cast<const uint4*>(rp)), rl = __ldg(reinterpret_cast<const uint4*>(rp + p.Cout));
                  const uint32_t hw[4] = {rh.x, rh.y, rh.z, rh.w}, lw[4] = {rl.x, rl.y, rl.z, rl.w};
#pragma unroll
                  for (int e = 0; e < 4; ++e) {
                    o[2 * e] += __uint_as_float(hw[e] << 16) + __uint_as_float(lw[e] << 16);
                    o[2 * e + 1] += __uint_as_float(hw[e] & 0xffff0000u) + __uint_as_float(lw[e] & 0xffff0000u);
                  }
                }
                if (p.relu) {
#pragma unroll
                  for (int e = 0; e < 8; ++e) o[e] = fmaxf(o[e], 0.f);
                }
                uint32_t hw[4], lw[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  hw[e] = pack_bf16x2(o[2 * e], o[2 * e + 1]);
                  lw[e] = pack_bf16x2(o[2 * e] - __uint_as_float(hw[e] << 16), o[2 * e + 1] - __uint_as_float(hw[e] & 0xffff0000u));
                }
                __nv_bfloat16* yp = reinterpret_cast<__nv_bfloat16*>(p.y) + (size_t)pgr * ypitch + co;
                *reinterpret_cast<uint4*>(yp) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
                *reinterpret_cast<uint4*>(yp + p.Cout) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
              } else if (p.y_bf16) {
                const float4 a = *reinterpret_cast<const float4*>(src), b = *reinterpret_cast<const float4*>(src + 4);
                float o[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
                if (p.residual) {
                  const uint4 rv = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(p.residual) + ridx));
                  const uint32_t rw[4] = {rv.x, rv.y, rv.z, rv.w};
#pragma unroll
                  for (int e = 0; e < 4; ++e) {
                    o[2 * e] += __uint_as_float(rw[e] << 16);
                    o[2 * e + 1] += __uint_as_float(rw[e] & 0xffff0000u);
                  }
                }
                if (p.relu) {
#pragma unroll
                  for (int e = 0; e < 8; ++e) o[e] = fmaxf(o[e], 0.f);
                }
                uint4 w;
                w.x = pack_bf16x2(o[0], o[1]); w.y = pack_bf16x2(o[2], o[3]);
                w.z = pack_bf16x2(o[4], o[5]); w.w = pack_bf16x2(o[6], o[7]);
                *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.y) + (size_t)pgr * p.Cout + co) = w;
              } else {
                float4 o = *reinterpret_cast<const float4*>(src);
                if (p.residual) {
                  const float4 rv = __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p.residual) + ridx));
                  o.x += rv.x; o.y += rv.y; o.z += rv.z; o.w += rv.w;
                }
                if (p.relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
                *reinterpret_cast<float4*>(reinterpret_cast<float*>(p.y) + (size_t)pgr * p.Cout + co) = o;
              }
            }
          }
          __syncwarp();
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_tempty + 8 * buf);
        continue;
      }
      // ---- legacy per-lane path (NCHW outputs, odd channel counts) ----
      for (int col = 0; col < p.BN; col += 16) {
        uint32_t rr[16];
        tmem_ld16(trow + (uint32_t)col, rr);  // warp-collective
        if (!row_ok) continue;
        const int co0 = n0 + col;
        if (co0 >= p.Cout) continue;
        float o16[16];
#pragma unroll
        for (int e = 0; e < 16; ++e) o16[e] = __uint_as_float(rr[e]);
        if (p.bias) {
#pragma unroll
          for (int e = 0; e < 16; ++e) if (co0 + e < p.Cout) o16[e] += __ldg(p.bias + co0 + e);
        }
        if (p.out_nhwc) {
          const size_t oidx = (size_t)pg * p.Cout + co0;
          // FPN top-down path (models/fpn.py:88-93): lateral conv + nearest-2x-upsampled coarser map, fused:
          // the residual is indexed at (ho/2, wo/2) of the half-resolution tensor instead of being materialised.
          size_t ridx = oidx;
          if (p.res_up2) {
            const int ho = pp / p.Wo, wo = pp - ho * p.Wo;
            ridx = (((size_t)n_img * (p.Ho >> 1) + (ho >> 1)) * (p.Wo >> 1) + (wo >> 1)) * p.Cout + co0;
          }
          const bool full = (co0 + 15 < p.Cout) && ((p.Cout & 7) == 0) && vec_ptrs_ok;
          if (p.y_bf16) {
            __nv_bfloat16* yo = reinterpret_cast<__nv_bfloat16*>(p.y) + oidx;
            const __nv_bfloat16* ro = p.residual ? reinterpret_cast<const __nv_bfloat16*>(p.residual) + ridx : nullptr;
            if (full) {
              if (ro) {
                const uint4 r0 = __ldg(reinterpret_cast<const uint4*>(ro)), r1 = __ldg(reinterpret_cast<const uint4*>(ro) + 1);
                const uint32_t rw[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                  o16[2 * q] += __uint_as_float(rw[q] << 16);
                  o16[2 * q + 1] += __uint_as_float(rw[q] & 0xffff0000u);
                }
              }
              if (p.relu) {
#pragma unroll
                for (int e = 0; e < 16; ++e) o16[e] = fmaxf(o16[e], 0.f);
              }
              uint4 w0, w1;
              w0.x = pack_bf16x2(o16[0], o16[1]); w0.y = pack_bf16x2(o16[2], o16[3]);
              w0.z = pack_bf16x2(o16[4], o16[5]); w0.w = pack_bf16x2(o16[6], o16[7]);
              w1.x = pack_bf16x2(o16[8], o16[9]); w1.y = pack_bf16x2(o16[10], o16[11]);
              w1.z = pack_bf16x2(o16[12], o16[13]); w1.w = pack_bf16x2(o16[14], o16[15]);
              reinterpret_cast<uint4*>(yo)[0] = w0;
              reinterpret_cast<uint4*>(yo)[1] = w1;
            } else {
#pragma unroll
              for (int e = 0; e < 16; ++e) {
                if (co0 + e >= p.Cout) break;
                float o = o16[e];
                if (ro) o += __bfloat162float(ro[e]);
                if (p.relu) o = fmaxf(o, 0.f);
                yo[e] = __float2bfloat16_rn(o);
              }
            }
          } else {
            float* yo = reinterpret_cast<float*>(p.y) + oidx;
            const float* ro = p.residual ? reinterpret_cast<const float*>(p.residual) + ridx : nullptr;
            if (full) {
#pragma unroll
              for (int g4 = 0; g4 < 4; ++g4) {
                float4 o = make_float4(o16[g4 * 4], o16[g4 * 4 + 1], o16[g4 * 4 + 2], o16[g4 * 4 + 3]);
                if (ro) {
                  const float4 rv = __ldg(reinterpret_cast<const float4*>(ro + g4 * 4));
                  o.x += rv.x; o.y += rv.y; o.z += rv.z; o.w += rv.w;
                }
                if (p.relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
                *reinterpret_cast<float4*>(yo + g4 * 4) = o;
              }
            } else {
#pragma unroll
              for (int e = 0; e < 16; ++e) {
                if (co0 + e >= p.Cout) break;
                float o = o16[e];
                if (ro) o += __ldg(ro + e);
                if (p.relu) o = fmaxf(o, 0.f);
                yo[e] = o;
              }
            }
          }
        } else {  // NCHW: for a fixed cout the 32 lanes of a warp write 32 consecutive pixels
#pragma unroll
          for (int e = 0; e < 16; ++e) {
            const int co = co0 + e;
            if (co >= p.Cout) break;
            const size_t oidx = ((size_t)n_img * p.Cout + co) * HoWo + pp;
            float o = o16[e];
            if (p.y_bf16) {
              if (p.residual) o += __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p.residual)[oidx]);
              if (p.relu) o = fmaxf(o, 0.f);
              reinterpret_cast<__nv_bfloat16*>(p.y)[oidx] = __float2bfloat16_rn(o);
            } else {
              if (p.residual) o += __ldg(reinterpret_cast<const float*>(p.residual) + oidx);
              if (p.relu) o = fmaxf(o, 0.f);
              if (p.sig_from >= 0 && co >= p.sig_from) o = 1.f / (1.f + expf(-o));
              reinterpret_cast<float*>(p.y)[oidx] = o;
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_tempty + 8 * buf);   // accumulator buffer may be overwritten
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == TC_EPILOGUE / 32) {
    tc_fence_after();
    tmem_dealloc(tmem_base, tmem_cols);
  }
}

// ----------------------------------------------------------------------------------------------
// weight pre-pack: fp32 [Cout,Cin,kh,kw] -> bf16 hi / lo planes [Cout_pad][KHW*Cin], k = tap*Cin + c
// ----------------------------------------------------------------------------------------------
__global__ void pack_weight_kernel(const float* __restrict__ w, int Cout, int Cin, int KHW, int Cout_pad, int Kp,
                                   uint16_t* __restrict__ hi, uint16_t* __restrict__ lo) {
  const size_t total = (size_t)Cout_pad * Kp;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int kk = (int)(i % (size_t)Kp);
    const int co = (int)(i / (size_t)Kp);
    const int tap = kk / Cin, c = kk - tap * Cin;
    const float v = (co < Cout && kk < KHW * Cin) ? w[((size_t)co * Cin + c) * KHW + tap] : 0.f;
    const __nv_bfloat16 h = __float2bfloat16_rn(v);
    const __nv_bfloat16 l = __float2bfloat16_rn(v - __bfloat162float(h));
    hi[i] = *reinterpret_cast<const uint16_t*>(&h);
    lo[i] = *reinterpret_cast<const uint16_t*>(&l);
  }
}

static int tc_cout_pad(int Cout) { return Cout <= 32 ? 32 : (Cout + 63) / 64 * 64; }
static int tc_kp(int Cin, int KHW) { return (KHW * Cin + TC_BK - 1) / TC_BK * TC_BK; }

size_t tc_packed_weight_bytes(int Cout, int Cin, int kh, int kw) {
  return (size_t)2 * tc_cout_pad(Cout) * tc_kp(Cin, kh * kw) * sizeof(uint16_t);
}

int tc_pack_weight(const float* w, int Cout, int Cin, int kh, int kw, void* packed, cudaStream_t stream) {
  const int Cout_pad = tc_cout_pad(Cout), KHW = kh * kw, Kp = tc_kp(Cin, KHW);
  uint16_t* hi = reinterpret_cast<uint16_t*>(packed);
  uint16_t* lo = hi + (size_t)Cout_pad * Kp;
  const size_t total = (size_t)Cout_pad * Kp;
  int blocks = (int)((total + 255) / 256);
  if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
  pack_weight_kernel<<<blocks, 256, 0, stream>>>(w, Cout, Cin, KHW, Cout_pad, Kp, hi, lo);
  UPS_CHECK_LAUNCH();
  return 0;
}

// Cin % 64 == 0 (NHWC gather / cp.async), or a tiny Cin <= 8 (the RGB stem: NCHW fp32 input, flattened K)
bool tc_supported(int Cin, int kh, int kw, int dg) {
  return ((Cin % TC_BK) == 0 || Cin <= 8) && dg == 1 && kh * kw <= 49 && kh <= 15 && kw <= 15;
}

int launch_igemm_tc(TcParams p, const void* packed, cudaStream_t stream) {
  const int KHW = p.kh * p.kw;
  if (!tc_supported(p.Cin, p.kh, p.kw, 1)) return UPSNET_E_UNSUPPORTED;
  {  // stride-1 dense layers on bf16 NHWC activations: operands by TMA, no gather threads (igemm_tma.cu)
    const int rc = launch_igemm_tma(p, packed, stream);
    if (rc != UPSNET_E_UNSUPPORTED) return rc;
  }
  if ((((uintptr_t)p.x) & 15) || (((uintptr_t)packed) & 15) || (((uintptr_t)p.y) & 15)) return UPSNET_E_BADARG;
  p.Cout_pad = tc_cout_pad(p.Cout);
  p.w_hi = reinterpret_cast<const uint16_t*>(packed);
  p.w_lo = p.w_hi + (size_t)p.Cout_pad * tc_kp(p.Cin, KHW);
  const bool deform = p.offset != nullptr;
  const bool smallc = (p.Cin % TC_BK) != 0;
  if (deform && (p.x_bf16 || p.x_pair) && (long long)p.H * p.W * p.Cin * (p.x_pair ? 2 : 1) >= (1ll << 31)) return UPSNET_E_UNSUPPORTED;   // int32 element offsets
  if (smallc && (deform || p.x_bf16 || p.x_pair || p.dh * (p.kh - 1) > 255 || p.dw * (p.kw - 1) > 255)) return UPSNET_E_UNSUPPORTED;
  if (p.x_pair && !p.x3) return UPSNET_E_UNSUPPORTED;          // pairs are the storage format of precision bf16x3
  if (p.sig_from >= 0 && (p.out_nhwc || p.y_bf16)) return UPSNET_E_UNSUPPORTED;   // sigmoid lives in the fp32 NCHW per-lane epilogue
  if (p.y_pair) {
    // pair output: the staged NHWC epilogue only (16-byte vectors), plain [hi Cout][lo Cout] grouping
    if (!p.out_nhwc || (p.Cout & 7) || (p.pair_group && p.pair_group != p.Cout)) return UPSNET_E_UNSUPPORTED;
    if (p.residual && (((uintptr_t)p.residual) & 15)) return UPSNET_E_UNSUPPORTED;
  }
  // tile N: as wide as possible (each gathered A tile is reused by BN couts)
  int BN = p.Cout_pad;
  const int bn_cap = p.x3 ? 128 : 256;
  if (BN > bn_cap) BN = (p.Cout_pad % bn_cap == 0) ? bn_cap : ((p.Cout_pad % 128 == 0) ? 128 : 64);
  {  // few output tiles (FC layers, coarse pyramid levels): narrower N tiles so that every SM gets work
    const long long mt = ((long long)p.N * p.Ho * p.Wo + TC_BM - 1) / TC_BM;
    while (BN > 64 && (BN % 32) == 0 && mt * (p.Cout_pad / BN) < kNumSMs && p.Cout_pad % (BN / 2) == 0 && ((BN / 2) % 32) == 0) BN /= 2;
  }
  p.BN = BN;
  // One persistent CTA per SM: give the smem ring everything that is left after the sample table.
  const int khw_l = smallc ? 2 : KHW;   // table region: [KHW][128] entries, or the 1 KB per-k table of the stem mode
  // deformable: a short ring leaves ~100 KB of the SM's 256 KB as L1 for the corner gathers (see the producer)
  int stages = deform ? 3 : TC_MAX_STAGES;
  TcSmem L = tc_smem_layout(deform, khw_l, BN, stages, p.x3 != 0);
  while (stages > 2 && L.total + 1024 > 220 * 1024) { --stages; L = tc_smem_layout(deform, khw_l, BN, stages, p.x3 != 0); }
  if (L.total + 1024 > 227 * 1024) return UPSNET_E_UNSUPPORTED;
  p.stages = stages;
  const long long Ptot = (long long)p.N * p.Ho * p.Wo;
  if (Ptot <= 0) return 0;
  static int sms = 0;
  if (sms == 0) {
    int dev = 0, v = kNumSMs;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
    sms = v > 0 ? v : kNumSMs;
  }
  p.tile_w = 16; p.tile_h = 8;
  auto dtiles = [&]() { return (long long)p.N * ((p.Wo + p.tile_w - 1) / p.tile_w) * ((p.Ho + p.tile_h - 1) / p.tile_h) * (p.Cout_pad / BN); };
  if (deform) {   // few tiles (coarse pyramid levels): smaller pixel blocks -> more CTAs, proportionally less gather work each
    if (dtiles() < sms / 2) { p.tile_w = 8; p.tile_h = 8; }
    if (dtiles() < sms / 2) { p.tile_h = 4; }
    static int tile_env = -1;     // tuning hook: UPSNET_DCN_TILE=168 | 88 | 84 forces the pixel block (pair / bf16 experiments)
    if (tile_env < 0) { const char* e = getenv("UPSNET_DCN_TILE"); tile_env = e ? atoi(e) : 0; }
    if (tile_env == 168) { p.tile_w = 16; p.tile_h = 8; }
    if (tile_env == 88) { p.tile_w = 8; p.tile_h = 8; }
    if (tile_env == 84) { p.tile_w = 8; p.tile_h = 4; }
  }
  const long long num_tiles = (deform ? dtiles() : ((Ptot + TC_BM - 1) / TC_BM) * (p.Cout_pad / BN));
  dim3 grid((unsigned)(num_tiles < sms ? num_tiles : sms));
  size_t smem = L.total + 1024;
  if (deform) {     // experiment hook: extra (unused) dynamic shared memory shrinks L1 -- measures the gather's L1 sensitivity
    static int extra = -1;
    if (extra < 0) { const char* e = getenv("UPSNET_DCN_EXTRA_SMEM_KB"); extra = e ? atoi(e) : 0; }
    if (extra > 0 && smem + (size_t)extra * 1024 <= 227 * 1024) smem += (size_t)extra * 1024;
  }
  // opt in to the full 227 KB once per process (kept out of the per-launch path: CUDA-graph capture)
  // opt in to the full 227 KB once per device (kept out of the per-launch path: CUDA-graph capture)
  static PerDeviceOnce configured;
  if (configured.need()) {
    UPS_CUDA(cudaFuncSetAttribute(igemm_tc_kernel<1, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    UPS_CUDA(cudaFuncSetAttribute(igemm_tc_kernel<0, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    UPS_CUDA(cudaFuncSetAttribute(igemm_tc_kernel<1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    UPS_CUDA(cudaFuncSetAttribute(igemm_tc_kernel<0, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    UPS_CUDA(cudaFuncSetAttribute(igemm_tc_kernel<1, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    UPS_CUDA(cudaFuncSetAttribute(igemm_tc_kernel<0, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    UPS_CUDA(cudaFuncSetAttribute(igemm_tc_kernel<2, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    // deformable instantiations: prefer the smallest shared-memory carve-out that fits, the rest of the 256 KB is L1
    (void)cudaFuncSetAttribute(igemm_tc_kernel<1, 1>, cudaFuncAttributePreferredSharedMemoryCarveout, 60);
    (void)cudaFuncSetAttribute(igemm_tc_kernel<1, 0>, cudaFuncAttributePreferredSharedMemoryCarveout, 60);
    (void)cudaFuncSetAttribute(igemm_tc_kernel<1, 2>, cudaFuncAttributePreferredSharedMemoryCarveout, 60);
  }
  if (smallc) {
    igemm_tc_kernel<2, 0><<<grid, TC_THREADS, smem, stream>>>(p);
  } else if (p.x_pair) {
    if (deform) igemm_tc_kernel<1, 2><<<grid, TC_THREADS, smem, stream>>>(p);
    else igemm_tc_kernel<0, 2><<<grid, TC_THREADS, smem, stream>>>(p);
  } else if (p.x_bf16) {
    if (p.x3) return UPSNET_E_UNSUPPORTED;   // the hi/lo split needs fp32 or pair activations
    if (deform) igemm_tc_kernel<1, 1><<<grid, TC_THREADS, smem, stream>>>(p);
    else igemm_tc_kernel<0, 1><<<grid, TC_THREADS, smem, stream>>>(p);
  } else {
    if (deform) igemm_tc_kernel<1, 0><<<grid, TC_THREADS, smem, stream>>>(p);
    else igemm_tc_kernel<0, 0><<<grid, TC_THREADS, smem, stream>>>(p);
  }
  UPS_CHECK_LAUNCH();
  return 0;
}

}  // namespace ups
