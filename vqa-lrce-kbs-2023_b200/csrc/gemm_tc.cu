// gemm_tc.cu — the dense contraction of the LRCE hot path on 5th-gen tensor cores (tcgen05 + TMEM + TMA).
//
//   out[M, N] = epilogue( A[M, K] (bf16, row-major) x W[N, K]^T (bf16, row-major = nn.Linear weight) )
//
// Replaces every nn.Linear / Conv3d-as-GEMM on the path (reference call sites: video_swin_ori.py:52-55 Mlp,
// :165/:187 qkv/proj, :340 PatchMerging.reduction, :475 PatchEmbed3D.proj; fusionv3.py:185 projection_layer and
// the K/V in-projections of nn.TransformerDecoderLayer built at fusionv3.py:8-17).
//
// Design (one persistent CTA per SM, 384 threads):
//   warp 0      TMA producer  : A tile 128x64 and W tile BNx64 per k-block, 128B-swizzled, 4-6 stage mbarrier ring
//   warp 1      MMA issuer    : one thread issues tcgen05.mma (M=128, N=BN, K=16) x4 per k-block into TMEM
//   warp 2      TMEM allocator: 2 accumulator buffers of BN fp32 columns (double-buffered against the epilogue)
//   warps 4-11  epilogue      : tcgen05.ld (thread = one accumulator row), fused bias / GELU(erf) / residual /
//                               LayerNorm. Two warps share each 32-lane TMEM quarter and split the 64-column slabs.
//                               Every warp stages its 32 x 64 bf16 slab in a private 128B-swizzled 4 KB buffer and
//                               issues its own TMA bulk store (full-line writes, M tail clipped by the descriptor,
//                               no CTA-wide barrier in the epilogue).
//   LayerNorm is never a separate pass inside a Swin block: residual epilogues emit per-row (mean, M2) partials of
//   what they write, and the next GEMM folds the normalisation into its epilogue (see GemmParams::in_stats).
// Both operands are K-major, so no transposes exist anywhere; M and K tails are handled by TMA zero fill.
#include <stdlib.h>

#include "host_common.h"
#include "lrce_common.cuh"

namespace lrce {

enum { EPI_BIAS = 0, EPI_BIAS_GELU = 1, EPI_BIAS_RESIDUAL = 2, EPI_BIAS_LN = 3 };

struct GemmParams {
  int M, N, K;
  const float* bias;     // [N] or nullptr
  const bf16* residual;  // [M, ldr] (EPI_BIAS_RESIDUAL), may alias out
  int ldr;
  void* out;  // [M, ldo] bf16 or fp32
  int ldo;
  const float* ln_g;  // EPI_BIAS_LN: LayerNorm over the N == BN output features
  const float* ln_b;
  float ln_eps;
  // LayerNorm folded into the A operand (LNIN kernels): A holds the RAW rows x, W holds W * diag(gamma), and
  //   out[m, n] = rstd[m] * (acc[m, n] - mean[m] * colsum[n]) + bias'[n]
  // with (mean, rstd) of row m rebuilt from the per-64-column partials `in_stats` the producing GEMM emitted,
  // colsum[n] = sum_k W'[n, k] and bias' = bias + W beta (both prepared once at weight-pack time).
  const float* in_stats;   // float2 [M][K/in_chunk]: (mean, M2) of each in_chunk-column chunk of row m (4, 8 or 16 chunks)
  int in_chunk;            // 32 or 64
  const float* in_colsum;  // [N]
  float in_eps;
  float* out_stats;  // nullptr, or float2 [M][N/cw]: (mean, M2) of every cw-column chunk of the output rows, cw = 32 for
                     // the 128-wide tile kernels (N % 256 != 0) and 64 otherwise (taken before the bf16 rounding: the
                     // rounding noise shifts the mean by ~2^-9 rms / sqrt(cw), far below bf16)
};

// does this configuration stage its residual tile through TMA? (see GemmCfg)
constexpr bool gemm_res_tma(int BN, int EPI, int CG) { return EPI == 2 /*EPI_BIAS_RESIDUAL*/ && (CG == 2 || BN == 128); }

constexpr int GEMM_BM = 128;
constexpr int GEMM_BK = 64;
constexpr int GEMM_THREADS = 640;   // warps 0-3: TMA producer, MMA issuer, TMEM allocator, row-statistics; warps 4-19: epilogue
constexpr int GEMM_EPI_WARPS = 16;  // four per TMEM lane quarter, each owning a quarter of the tile's columns

// RES: the residual tile of the EPI_BIAS_RESIDUAL epilogue (this CTA's 128 rows x BN columns) is staged in shared memory by
// TMA a tile ahead of its use (CTA pairs and 128-wide tiles; the 128 x 256 single-CTA tiles have no room for it)
template <int BN, int CG, bool RES = false>
struct GemmCfg {
  static constexpr int STAGES = (BN == 256 && CG == 1) ? 4 : (RES ? 4 : 6);
  static constexpr int A_BYTES = GEMM_BM * GEMM_BK * 2;
  static constexpr int B_BYTES = (BN / CG) * GEMM_BK * 2;  // a CTA pair splits the B tile between its two CTAs
  static constexpr int TMEM_COLS = 2 * BN;  // 256 or 512: power of two
  static constexpr int SLAB_BYTES = 32 * 32 * 2;  // one epilogue warp's 32-row x 32-column bf16 output slab (64B-swizzled)
  static constexpr int RN_BYTES = 2 * GEMM_BM * 8;  // (rstd, -rstd*mean) of the rows of two tiles in flight
  static constexpr int RES_BYTES = RES ? GEMM_BM * BN * 2 : 0;  // BN / 32 column pieces of 128 rows x 64 B, 64B-swizzled
  static constexpr int SMEM_BYTES = STAGES * (A_BYTES + B_BYTES) + GEMM_EPI_WARPS * SLAB_BYTES + RES_BYTES + RN_BYTES + 256 /*barriers*/;
  static_assert(SMEM_BYTES <= 232448, "GEMM shared-memory budget");
};

// erf-GELU (nn.GELU default, video_swin_ori.py:42) as 0.5 x (1 + tanh(u)), u = x (a + b x^2 + c x^4) fitted to the erf form
// (0.5 (1 + tanh u) == sigmoid(2u) ~= Phi(x)): max |error| 2.5e-5 over all x before the approximate tanh, whose 2^-11
// relative error stays below half a bf16 ulp of the result. Two values at a time: 6 packed FMA-pipe ops + ONE MUFU per
// value — the fc1 epilogues of stages 1-2 are bound by the 16 MUFU / clk / SM, so the sigmoid form (ex2 + rcp) cost twice.
__device__ __forceinline__ float2 gelu_tanh2(float2 x) {
  float2 x2 = fmul2(x, x);
  x2.x = fminf(x2.x, 64.0f);
  x2.y = fminf(x2.y, 64.0f);
  float2 p = ffma2(x2, make_float2(-0.00035151677f, -0.00035151677f), make_float2(0.03700564325f, 0.03700564325f));
  p = ffma2(p, x2, make_float2(0.79750783595f, 0.79750783595f));
  const float2 u = fmul2(p, x);
  float2 t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t.x) : "f"(u.x));
  asm("tanh.approx.f32 %0, %1;" : "=f"(t.y) : "f"(u.y));
  const float2 hx = fmul2(x, make_float2(0.5f, 0.5f));
  return ffma2(hx, t, hx);
}

// (mean, M2) of 32 values: four independent packed accumulators (a serial chain of 32 dependent adds costs ~130 cycles of
// latency per piece in an epilogue that is latency-bound on the short-K shapes)
__device__ __forceinline__ float2 stats32(const float (&v)[32]) {
  float2 s0 = make_float2(v[0], v[1]), s1 = make_float2(v[2], v[3]), s2 = make_float2(v[4], v[5]), s3 = make_float2(v[6], v[7]);
#pragma unroll
  for (int i = 8; i < 32; i += 8) {
    s0 = fadd2(s0, make_float2(v[i + 0], v[i + 1]));
    s1 = fadd2(s1, make_float2(v[i + 2], v[i + 3]));
    s2 = fadd2(s2, make_float2(v[i + 4], v[i + 5]));
    s3 = fadd2(s3, make_float2(v[i + 6], v[i + 7]));
  }
  s0 = fadd2(fadd2(s0, s1), fadd2(s2, s3));
  const float mean = (s0.x + s0.y) * (1.0f / 32);
  const float2 nm = make_float2(-mean, -mean);
  float2 m0 = make_float2(0.f, 0.f), m1 = m0, m2 = m0, m3 = m0;
#pragma unroll
  for (int i = 0; i < 32; i += 8) {
    const float2 d0 = fadd2(make_float2(v[i + 0], v[i + 1]), nm), d1 = fadd2(make_float2(v[i + 2], v[i + 3]), nm);
    const float2 d2 = fadd2(make_float2(v[i + 4], v[i + 5]), nm), d3 = fadd2(make_float2(v[i + 6], v[i + 7]), nm);
    m0 = ffma2(d0, d0, m0);
    m1 = ffma2(d1, d1, m1);
    m2 = ffma2(d2, d2, m2);
    m3 = ffma2(d3, d3, m3);
  }
  m0 = fadd2(fadd2(m0, m1), fadd2(m2, m3));
  return make_float2(mean, m0.x + m0.y);
}
// Chan's combination of two equally sized groups of n values each
__device__ __forceinline__ float2 stats_merge(float2 a, float2 b, float n) {
  const float d = a.x - b.x;
  return make_float2(0.5f * (a.x + b.x), a.y + b.y + 0.5f * n * d * d);
}
// Row statistics of the K-wide LayerNorm input from its NC = K/cw chunk partials -> (rstd, -rstd * mean), for the
// ROWS rows lane + 32 * (i0 + i) of one 128-row tile. `stats_fetch` issues every load (they are independent: one L2
// round trip for the whole batch), `stats_reduce` combines the partials with Chan's formula.
template <int NC, int ROWS>
__device__ __forceinline__ void stats_fetch(float2 (&t)[ROWS][NC], const float* stats, int M, int m0, int lane, int i0) {
#pragma unroll
  for (int i = 0; i < ROWS; ++i) {
    const int row = min(m0 + lane + 32 * (i0 + i), M - 1);  // tail rows of the last tile read a valid row, never stored
    const float4* st = reinterpret_cast<const float4*>(stats) + static_cast<size_t>(row) * (NC / 2);  // NC float2 per row
#pragma unroll
    for (int c = 0; c < NC; c += 2) {
      const float4 v = __ldg(st + c / 2);
      t[i][c] = make_float2(v.x, v.y);
      t[i][c + 1] = make_float2(v.z, v.w);
    }
  }
}
template <int NC, int ROWS>
__device__ __forceinline__ void stats_reduce(const float2 (&t)[ROWS][NC], float2* s_rn, int lane, int i0, float cw, int K,
                                             float eps) {
#pragma unroll
  for (int i = 0; i < ROWS; ++i) {
    float sm = 0.f;
#pragma unroll
    for (int c = 0; c < NC; ++c) sm += t[i][c].x;
    const float mean = sm * (1.0f / NC);
    float m2 = 0.f;
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      const float d = t[i][c].x - mean;
      m2 += fmaf(cw * d, d, t[i][c].y);
    }
    const float rstd = rsqrtf(m2 / K + eps);
    s_rn[lane + 32 * (i0 + i)] = make_float2(rstd, -rstd * mean);
  }
}

// bias (or folded LayerNorm) / GELU / residual on 32 accumulator columns of one row -> v[32] fp32
template <int EPI, bool LNIN>
__device__ __forceinline__ void epilogue_values(const uint32_t (&acc)[32], int col0, const GemmParams& p, float2 rn,
                                                const uint4 (&res)[4], float (&v)[32]) {
  if (LNIN) {
    const float4* c4 = reinterpret_cast<const float4*>(p.in_colsum + col0);
    const float4* b4 = reinterpret_cast<const float4*>(p.bias + col0);
    const float2 rx = make_float2(rn.x, rn.x), ry = make_float2(rn.y, rn.y);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float4 c = __ldg(c4 + i), b = __ldg(b4 + i);
      const float2 lo = ffma2(rx, make_float2(__uint_as_float(acc[4 * i + 0]), __uint_as_float(acc[4 * i + 1])),
                              ffma2(ry, make_float2(c.x, c.y), make_float2(b.x, b.y)));
      const float2 hi = ffma2(rx, make_float2(__uint_as_float(acc[4 * i + 2]), __uint_as_float(acc[4 * i + 3])),
                              ffma2(ry, make_float2(c.z, c.w), make_float2(b.z, b.w)));
      v[4 * i + 0] = lo.x; v[4 * i + 1] = lo.y; v[4 * i + 2] = hi.x; v[4 * i + 3] = hi.y;
    }
  } else {
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(acc[i]);
    if (p.bias != nullptr) {
      const float4* b4 = reinterpret_cast<const float4*>(p.bias + col0);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float4 b = __ldg(b4 + i);
        const float2 lo = fadd2(make_float2(v[4 * i + 0], v[4 * i + 1]), make_float2(b.x, b.y));
        const float2 hi = fadd2(make_float2(v[4 * i + 2], v[4 * i + 3]), make_float2(b.z, b.w));
        v[4 * i + 0] = lo.x; v[4 * i + 1] = lo.y; v[4 * i + 2] = hi.x; v[4 * i + 3] = hi.y;
      }
    }
  }
  if (EPI == EPI_BIAS_GELU) {
#pragma unroll
    for (int i = 0; i < 32; i += 2) {
      const float2 g = gelu_tanh2(make_float2(v[i], v[i + 1]));
      v[i] = g.x; v[i + 1] = g.y;
    }
  }
  if (EPI == EPI_BIAS_RESIDUAL) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const uint32_t w[4] = {res[i].x, res[i].y, res[i].z, res[i].w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float2 f = fadd2(make_float2(v[8 * i + 2 * k], v[8 * i + 2 * k + 1]), bf16x2_to_f32x2(w[k]));
        v[8 * i + 2 * k] = f.x; v[8 * i + 2 * k + 1] = f.y;
      }
    }
  }
}

// round v[32] to bf16, packed into 4 x 16 B
__device__ __forceinline__ void pack32(const float (&v)[32], uint4 (&o)[4]) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    o[i].x = pack_bf16x2(v[8 * i + 0], v[8 * i + 1]);
    o[i].y = pack_bf16x2(v[8 * i + 2], v[8 * i + 3]);
    o[i].z = pack_bf16x2(v[8 * i + 4], v[8 * i + 5]);
    o[i].w = pack_bf16x2(v[8 * i + 6], v[8 * i + 7]);
  }
}

// fp32 output path (bias epilogue only): direct 16-byte global stores of 32 accumulator columns of one row
template <int EPI, typename OutT>
__device__ __forceinline__ void epilogue_chunk(const uint32_t (&acc)[32], int row, int col0, const GemmParams& p) {
  static_assert(sizeof(OutT) == 4 && EPI == EPI_BIAS, "direct stores are only used for fp32 output");
  if (row >= p.M) return;
  float v[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(acc[i]);
  if (p.bias != nullptr) {
    const float4* b4 = reinterpret_cast<const float4*>(p.bias + col0);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float4 b = __ldg(b4 + i);
      v[4 * i + 0] += b.x;
      v[4 * i + 1] += b.y;
      v[4 * i + 2] += b.z;
      v[4 * i + 3] += b.w;
    }
  }
  float4* o4 = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + static_cast<size_t>(row) * p.ldo + col0);
#pragma unroll
  for (int i = 0; i < 8; ++i) o4[i] = make_float4(v[4 * i + 0], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
}

// CG = 1: one CTA per SM computes 128 x BN tiles. CG = 2: the two CTAs of a 2-cluster (one TPC) compute a 256 x BN tile
// with ONE tcgen05.mma.cta_group::2 stream issued by the leader: each CTA stages its own 128 rows of A and half of the
// B tile (a third less L2 -> shared-memory traffic per FLOP, half the B reads per MMA) and drains its own 128
// accumulator rows.
template <int BN, int EPI, typename OutT, bool LNIN, int CG>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmR, const GemmParams p) {
  constexpr bool TMA_STORE = (sizeof(OutT) == 2) && (EPI != EPI_BIAS_LN);
  static_assert(CG == 1 || (CG == 2 && TMA_STORE), "CTA pairs are used with the TMA-store epilogues only");
  constexpr bool RES_TMA = gemm_res_tma(BN, EPI, CG);
  using Cfg = GemmCfg<BN, CG, RES_TMA>;
  constexpr int STAGES = Cfg::STAGES;
  constexpr int WCOLS = BN / 4;      // columns of the tile owned by one epilogue warp
  constexpr int PIECES = WCOLS / 32;  // 32-column pieces per warp and tile
  // 1024-byte alignment is required by the 128B swizzle atoms; the pointer stays derived from the __shared__ array so
  // that the epilogue's slab writes compile to STS (an integer round trip would demote them to generic stores)
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * Cfg::A_BYTES;
  uint8_t* sC = sB + STAGES * Cfg::B_BYTES;  // 16 per-warp output slabs (stage sizes are multiples of 1024)
  uint8_t* sRes = sC + GEMM_EPI_WARPS * Cfg::SLAB_BYTES;  // residual tile (RES_TMA), 512-byte aligned pieces
  float2* s_rn = reinterpret_cast<float2*>(sRes + Cfg::RES_BYTES);
  uint64_t* bar_full = reinterpret_cast<uint64_t*>(sRes + Cfg::RES_BYTES + Cfg::RN_BYTES);
  uint64_t* bar_empty = bar_full + STAGES;
  uint64_t* bar_tfull = bar_empty + STAGES;
  uint64_t* bar_tempty = bar_tfull + 2;
  uint64_t* bar_rnfull = bar_tempty + 2;
  uint64_t* bar_rnempty = bar_rnfull + 2;
  uint64_t* bar_resfull = bar_rnempty + 2;
  uint64_t* bar_resempty = bar_resfull + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_resempty + 1);
  // LayerNorm epilogue scratch [tile parity][column quarter][sum | sumsq][row] aliases the (then unused) output slabs
  float* ln_part = reinterpret_cast<float*>(sC);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n_tiles_n = p.N / BN;
  const int n_tiles_m = (p.M + GEMM_BM * CG - 1) / (GEMM_BM * CG);
  const int n_tiles = n_tiles_m * n_tiles_n;
  const int n_kb = (p.K + GEMM_BK - 1) / GEMM_BK;
  // work unit = CTA (CG = 1) or CTA pair (CG = 2); `row_off` = this CTA's rows inside the unit's M tile
  const int cta_rank = (CG == 2) ? static_cast<int>(cluster_ctarank()) : 0;
  const int unit = (CG == 2) ? static_cast<int>(blockIdx.x >> 1) : static_cast<int>(blockIdx.x);
  const int n_units = (CG == 2) ? static_cast<int>(gridDim.x >> 1) : static_cast<int>(gridDim.x);
  const int row_off = cta_rank * GEMM_BM;
  const bool leader = cta_rank == 0;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (TMA_STORE) tma_prefetch_desc(&tmC);
    if (RES_TMA) tma_prefetch_desc(&tmR);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&bar_full[s], 1);
      mbar_init(&bar_empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&bar_tfull[a], 1);
      mbar_init(&bar_tempty[a], GEMM_EPI_WARPS * CG);  // CG = 2: the epilogue warps of both CTAs arrive on the leader's
      mbar_init(&bar_rnfull[a], 1);
      mbar_init(&bar_rnempty[a], GEMM_EPI_WARPS);
    }
    mbar_init(bar_resfull, 1);
    mbar_init(bar_resempty, GEMM_EPI_WARPS);
    fence_barrier_init();
  }
  if (warp == 2) {
    if (CG == 2) tmem_alloc_pair(tmem_slot, Cfg::TMEM_COLS);
    else tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
  }
  tcgen05_fence_before();
  if (CG == 2) cluster_sync_all();  // the peer's barriers must be initialised before any remote arrival / TMA credit
  else __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // everything above overlaps the tail of the previous kernel of the stream (programmatic dependent launch); global memory is
  // touched only below. The next kernel may be scheduled as soon as every CTA got here (it holds its TMEM by then).
  griddep_wait();
  griddep_launch();

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0, rit = 0;
      uint32_t phase = 0;
      for (int t = unit; t < n_tiles; t += n_units) {
        const int m0 = (t / n_tiles_n) * (GEMM_BM * CG) + row_off;
        const int n0 = (t % n_tiles_n) * BN;
        for (int kb = 0; kb < n_kb; ++kb) {
          mbar_wait_parked(&bar_empty[stage], phase ^ 1);
          if (CG == 2) {
            // both CTAs' bytes are credited to the leader's barrier, which alone expects them
            if (leader) mbar_expect_tx(&bar_full[stage], 2 * (Cfg::A_BYTES + Cfg::B_BYTES));
            tma_load_2d_pair(sA + stage * Cfg::A_BYTES, &tmA, &bar_full[stage], kb * GEMM_BK, m0);
            tma_load_2d_pair(sB + stage * Cfg::B_BYTES, &tmB, &bar_full[stage], kb * GEMM_BK, n0 + cta_rank * (BN / 2));
          } else {
            mbar_expect_tx(&bar_full[stage], Cfg::A_BYTES + Cfg::B_BYTES);
            tma_load_2d(sA + stage * Cfg::A_BYTES, &tmA, &bar_full[stage], kb * GEMM_BK, m0);
            tma_load_2d(sB + stage * Cfg::B_BYTES, &tmB, &bar_full[stage], kb * GEMM_BK, n0);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        if (RES_TMA) {
          // this tile's residual rows: the buffer is free once every epilogue warp has read the previous tile's rows; the
          // load then has the rest of this tile's main loop to land (own shared memory and barrier, also in a CTA pair)
          mbar_wait_parked(bar_resempty, (rit & 1) ^ 1);
          mbar_expect_tx(bar_resfull, Cfg::RES_BYTES);
#pragma unroll
          for (int pc = 0; pc < BN / 32; ++pc) tma_load_2d(sRes + pc * (GEMM_BM * 64), &tmR, bar_resfull, n0 + pc * 32, m0);
          ++rit;
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (the pair's leader)
    // the whole warp walks the loop and one elected lane issues: with the warp converged the descriptors stay in uniform
    // registers and an MMA costs ~3 instructions (a lane-0 branch makes every operand a per-MMA R2UR waterfall, ~60 cycles)
    if (leader) {
      constexpr uint32_t idesc = umma_idesc_bf16(GEMM_BM * CG, BN);
      const uint32_t a_lo0 = desc_lo(smem_u32(sA)), b_lo0 = desc_lo(smem_u32(sB));
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int t = unit; t < n_tiles; t += n_units, ++it) {
        const int as = it & 1;
        const uint32_t aphase = (it >> 1) & 1;
        mbar_wait_parked(&bar_tempty[as], aphase ^ 1);
        tcgen05_fence_after();
        const uint32_t tmem_d = tmem_base + as * BN;
        for (int kb = 0; kb < n_kb; ++kb) {
          mbar_wait_parked(&bar_full[stage], phase);
          tcgen05_fence_after();
          const uint32_t a_lo = a_lo0 + stage * (Cfg::A_BYTES >> 4), b_lo = b_lo0 + stage * (Cfg::B_BYTES >> 4);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < GEMM_BK / 16; ++k)
              umma_lo_acc<CG>(tmem_d, a_lo + k * 2, b_lo + k * 2, idesc, k != 0 ? 1u : static_cast<uint32_t>(kb));
            if (CG == 2) umma_commit_pair(&bar_empty[stage]);  // frees the stage in both CTAs
            else umma_commit(&bar_empty[stage]);
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        if (elect_one()) {
          if (CG == 2) umma_commit_pair(&bar_tfull[as]);  // each CTA's epilogue drains its own half
          else umma_commit(&bar_tfull[as]);
        }
        __syncwarp();
      }
    }
  } else if (warp == 3) {
    // ------------------------------------------------------------------ row statistics of the folded LayerNorm
    // runs one tile ahead of the epilogue: the L2 round trip for the partials never sits on the epilogue's path
    if (LNIN) {
      const int n_chunks = p.K / p.in_chunk;
      const float cw = static_cast<float>(p.in_chunk);
      auto publish = [&](int it) {
        __syncwarp();
        if (lane == 0) mbar_arrive(&bar_rnfull[it & 1]);
      };
      if (n_chunks == 4) {
        // narrow rows (C = 128 / 256): tiles are short, so the partials of tile i+1 are already in flight while tile i
        // is reduced (two register sets, loop unrolled by two)
        float2 ta[4][4], tb[4][4];
        int t = unit, it = 0;
        if (t < n_tiles) stats_fetch<4, 4>(ta, p.in_stats, p.M, (t / n_tiles_n) * (GEMM_BM * CG) + row_off, lane, 0);
        while (t < n_tiles) {
          int tn = t + n_units;
          if (tn < n_tiles) stats_fetch<4, 4>(tb, p.in_stats, p.M, (tn / n_tiles_n) * (GEMM_BM * CG) + row_off, lane, 0);
          mbar_wait_parked(&bar_rnempty[it & 1], ((it >> 1) & 1) ^ 1);
          stats_reduce<4, 4>(ta, s_rn + (it & 1) * GEMM_BM, lane, 0, cw, p.K, p.in_eps);
          publish(it);
          t = tn; ++it;
          if (t >= n_tiles) break;
          tn = t + n_units;
          if (tn < n_tiles) stats_fetch<4, 4>(ta, p.in_stats, p.M, (tn / n_tiles_n) * (GEMM_BM * CG) + row_off, lane, 0);
          mbar_wait_parked(&bar_rnempty[it & 1], ((it >> 1) & 1) ^ 1);
          stats_reduce<4, 4>(tb, s_rn + (it & 1) * GEMM_BM, lane, 0, cw, p.K, p.in_eps);
          publish(it);
          t = tn; ++it;
        }
      } else {
        int it = 0;
        for (int t = unit; t < n_tiles; t += n_units, ++it) {
          const int m0 = (t / n_tiles_n) * (GEMM_BM * CG) + row_off;
          float2* dst = s_rn + (it & 1) * GEMM_BM;
          if (n_chunks == 8) {
            float2 ta[4][8];
            stats_fetch<8, 4>(ta, p.in_stats, p.M, m0, lane, 0);
            mbar_wait_parked(&bar_rnempty[it & 1], ((it >> 1) & 1) ^ 1);
            stats_reduce<8, 4>(ta, dst, lane, 0, cw, p.K, p.in_eps);
          } else {  // 16 chunks (host-checked): two rows at a time
            mbar_wait_parked(&bar_rnempty[it & 1], ((it >> 1) & 1) ^ 1);
#pragma unroll 1
            for (int i0 = 0; i0 < 4; i0 += 2) {
              float2 ta[2][16];
              stats_fetch<16, 2>(ta, p.in_stats, p.M, m0, lane, i0);
              stats_reduce<16, 2>(ta, dst, lane, i0, cw, p.K, p.in_eps);
            }
          }
          publish(it);
        }
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ epilogue
    const int q = warp & 3;            // TMEM lane quarter this warp may access
    const int part = (warp - 4) >> 2;  // which quarter of the tile's columns
    int it = 0;
    for (int t = unit; t < n_tiles; t += n_units, ++it) {
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      const int m0 = (t / n_tiles_n) * (GEMM_BM * CG) + row_off;
      const int n0 = (t % n_tiles_n) * BN;
      const int row_in_tile = q * 32 + lane;
      const int row = m0 + row_in_tile;
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * BN + part * WCOLS;
      if constexpr (EPI == EPI_BIAS_LN) {
        static_assert(BN == 128 || EPI != EPI_BIAS_LN, "the LayerNorm epilogue normalises over one 128-wide tile");
        mbar_wait_parked(&bar_tfull[as], aphase);
        tcgen05_fence_after();
        // full-row LayerNorm: BN == N == 128, this thread holds 32 of the row's 128 features
        float v[32];
        {
          uint32_t acc[32];
          tmem_ld_32x32(taddr, acc);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(acc[i]) + __ldg(p.bias + part * 32 + i);
        }
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bar_tempty[as]);
        float s = 0.f, ss = 0.f;
#pragma unroll
        for (int i = 0; i < 32; ++i) { s += v[i]; ss += v[i] * v[i]; }
        const int par = it & 1;
        ln_part[((par * 4 + part) * 2 + 0) * GEMM_BM + row_in_tile] = s;
        ln_part[((par * 4 + part) * 2 + 1) * GEMM_BM + row_in_tile] = ss;
        asm volatile("bar.sync 1, 512;" ::: "memory");
        s = ss = 0.f;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          s += ln_part[((par * 4 + k) * 2 + 0) * GEMM_BM + row_in_tile];
          ss += ln_part[((par * 4 + k) * 2 + 1) * GEMM_BM + row_in_tile];
        }
        const float mean = s * (1.0f / BN);
        const float var = fmaxf(ss * (1.0f / BN) - mean * mean, 0.f);
        const float rstd = rsqrtf(var + p.ln_eps);
        if (row < p.M) {
          bf16* o = reinterpret_cast<bf16*>(p.out) + static_cast<size_t>(row) * p.ldo + part * 32;
          float y[32];
#pragma unroll
          for (int j = 0; j < 32; ++j)
            y[j] = (v[j] - mean) * rstd * __ldg(p.ln_g + part * 32 + j) + __ldg(p.ln_b + part * 32 + j);
          uint4 o4[4];
          pack32(y, o4);
#pragma unroll
          for (int j = 0; j < 4; ++j) *reinterpret_cast<uint4*>(o + 8 * j) = o4[j];
          if (p.out_stats != nullptr)
            reinterpret_cast<float2*>(p.out_stats)[static_cast<size_t>(row) * (BN / 32) + part] = stats32(y);
        }
      } else if constexpr (TMA_STORE) {
        // Each warp owns WCOLS columns of its 32 rows, in 32-column pieces: tcgen05.ld -> math -> private 64B-swizzled
        // 2 KB staging slab -> its own TMA bulk store. No CTA-wide barrier: a warp only ever waits for its own previous
        // store to have drained the slab, and the accumulator is released right after the warp's last tcgen05.ld.
        uint8_t* slab = sC + (warp - 4) * Cfg::SLAB_BYTES;
        float2 rn = make_float2(1.f, 0.f);
        // residual rows are fetched one piece ahead (the first one before the accumulator is even ready)
        uint4 res_next[4] = {make_uint4(0, 0, 0, 0), make_uint4(0, 0, 0, 0), make_uint4(0, 0, 0, 0), make_uint4(0, 0, 0, 0)};
        const bf16* res_row = nullptr;
        if (EPI == EPI_BIAS_RESIDUAL && !RES_TMA && row < p.M) {
          res_row = p.residual + static_cast<size_t>(row) * p.ldr + n0 + part * WCOLS;
#pragma unroll
          for (int i = 0; i < 4; ++i) res_next[i] = reinterpret_cast<const uint4*>(res_row)[i];
        }
        if (EPI == EPI_BIAS_RESIDUAL && !RES_TMA) {
          // The residual stream was last touched several kernels ago: its rows come from HBM (~1.5 us under load), which
          // a one-piece-ahead register prefetch cannot cover. Pull this thread's slice of the NEXT tile into L2 now, a whole
          // tile ahead of its use.
          const int tn = t + n_units;
          if (tn < n_tiles) {
            const int rown = (tn / n_tiles_n) * (GEMM_BM * CG) + row_off + row_in_tile;
            if (rown < p.M) {
              const bf16* nxt = p.residual + static_cast<size_t>(rown) * p.ldr + (tn % n_tiles_n) * BN + part * WCOLS;
              asm volatile("prefetch.global.L2 [%0];" ::"l"(nxt));
            }
          }
        }
        if (LNIN) {
          mbar_wait_parked(&bar_rnfull[as], aphase);
          rn = s_rn[as * GEMM_BM + row_in_tile];
          __syncwarp();
          if (lane == 0) mbar_arrive(&bar_rnempty[as]);
        }
        if (RES_TMA) mbar_wait_parked(bar_resfull, it & 1);  // landed long ago: issued a main loop ahead
        mbar_wait_parked(&bar_tfull[as], aphase);
        tcgen05_fence_after();
        float2 st_carry = make_float2(0.f, 0.f);
#pragma unroll
        for (int pc = 0; pc < PIECES; ++pc) {
          const int col_in_tile = part * WCOLS + pc * 32;
          uint32_t acc[32];
          tmem_ld_32x32(taddr + pc * 32, acc);
          uint4 res_cur[4];
          if (RES_TMA) {
            // piece = 128 rows x 64 B, 64B-swizzled like the output slabs: conflict-free row-per-lane reads
            const uint8_t* rp = sRes + (col_in_tile >> 5) * (GEMM_BM * 64) + row_in_tile * 64;
#pragma unroll
            for (int i = 0; i < 4; ++i) res_cur[i] = *reinterpret_cast<const uint4*>(rp + ((i ^ ((row_in_tile >> 1) & 3)) << 4));
            if (pc + 1 == PIECES) {  // this warp is done with the residual tile
              __syncwarp();
              if (lane == 0) mbar_arrive(bar_resempty);
            }
          } else {
#pragma unroll
            for (int i = 0; i < 4; ++i) res_cur[i] = res_next[i];
            if (EPI == EPI_BIAS_RESIDUAL && res_row != nullptr && pc + 1 < PIECES) {
#pragma unroll
              for (int i = 0; i < 4; ++i) res_next[i] = reinterpret_cast<const uint4*>(res_row + (pc + 1) * 32)[i];
            }
          }
          tmem_ld_wait();
          if (pc + 1 == PIECES) {  // every tcgen05.ld of this warp for this tile has completed: release the accumulator
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) {
              if (CG == 2) mbar_arrive_leader(&bar_tempty[as]);  // the leader's MMA thread waits for both halves
              else mbar_arrive(&bar_tempty[as]);
            }
          }
          float v[32];
          epilogue_values<EPI, LNIN>(acc, n0 + col_in_tile, p, rn, res_cur, v);
          uint4 o[4];
          pack32(v, o);
          if (p.out_stats != nullptr) {
            const float2 st_piece = stats32(v);
            if (PIECES == 1) {  // 128-wide tiles: 32-column chunks
              if (row < p.M)
                reinterpret_cast<float2*>(p.out_stats)[static_cast<size_t>(row) * (p.N >> 5) + ((n0 + col_in_tile) >> 5)] = st_piece;
            } else if (pc == 0) {
              st_carry = st_piece;
            } else if (row < p.M) {  // 256-wide tiles: this warp's two pieces form one 64-column chunk
              reinterpret_cast<float2*>(p.out_stats)[static_cast<size_t>(row) * (p.N >> 6) + ((n0 + part * WCOLS) >> 6)] =
                  stats_merge(st_carry, st_piece, 32.0f);
            }
          }
          // the previous TMA store of this warp must have finished reading the slab before it is refilled
          if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
          __syncwarp();
#pragma unroll
          for (int i = 0; i < 4; ++i)  // 64B swizzle: 16-byte chunk c of 64-byte row r lives at chunk position c ^ ((r >> 1) & 3)
            *reinterpret_cast<uint4*>(slab + lane * 64 + ((i ^ ((lane >> 1) & 3)) << 4)) = o[i];
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                             reinterpret_cast<uint64_t>(&tmC)),
                         "r"(smem_u32(slab)), "r"(n0 + col_in_tile), "r"(m0 + q * 32)
                         : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          }
        }
      } else if constexpr (sizeof(OutT) == 4) {
        mbar_wait_parked(&bar_tfull[as], aphase);
        tcgen05_fence_after();
#pragma unroll 1
        for (int c = 0; c < WCOLS; c += 32) {
          uint32_t acc[32];
          tmem_ld_32x32(taddr + c, acc);
          tmem_ld_wait();
          epilogue_chunk<EPI, OutT>(acc, row, n0 + part * WCOLS + c, p);
        }
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bar_tempty[as]);
      }
    }
    if (TMA_STORE && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }

  tcgen05_fence_before();
  if (CG == 2) cluster_sync_all();  // neither CTA may exit while its peer can still read its smem / signal its barriers
  else __syncthreads();
  if (warp == 2) {
    tcgen05_fence_after();
    if (CG == 2) tmem_dealloc_pair(tmem_base, Cfg::TMEM_COLS);
    else tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

template <int BN, int EPI, typename OutT, bool LNIN, int CG>
static int launch_gemm(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC, const GemmParams& p,
                       cudaStream_t stream) {
  using Cfg = GemmCfg<BN, CG, gemm_res_tma(BN, EPI, CG)>;
  CUtensorMap tmR = tmC;  // unused unless the configuration stages its residual through TMA
  if (gemm_res_tma(BN, EPI, CG)) {
    const int rc = make_tmap_2d_bf16(&tmR, p.residual, p.N, p.M, p.ldr, 32, GEMM_BM, /*swizzle_bytes=*/64);
    if (rc != LRCE_OK) return rc;
  }
  auto kern = gemm_tc_kernel<BN, EPI, OutT, LNIN, CG>;
  static thread_local uint64_t configured = 0;  // one bit per device: function attributes are per device
  if (needs_device_setup(&configured)) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
    if (e != cudaSuccess) {
      set_error("cudaFuncSetAttribute(gemm_tc_kernel, smem=%d): %s", Cfg::SMEM_BYTES, cudaGetErrorString(e));
      return LRCE_ECUDA;
    }
    mark_device_setup(&configured);
  }
  const int n_tiles = ((p.M + GEMM_BM * CG - 1) / (GEMM_BM * CG)) * (p.N / BN);
  int units = sm_count() / CG;
  if (n_tiles < units) units = n_tiles;
  if (CG == 1) {
    cudaError_t e = launch_pdl(kern, dim3(units), dim3(GEMM_THREADS), Cfg::SMEM_BYTES, stream, tmA, tmB, tmC, tmR, p);
    if (e != cudaSuccess) {
      set_error("cudaLaunchKernelEx(gemm_tc_kernel): %s", cudaGetErrorString(e));
      return LRCE_ECUDA;
    }
  } else {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(units * CG);
    cfg.blockDim = dim3(GEMM_THREADS);
    cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CG;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1 + pdl_attr(attr + 1);
    cudaError_t e = cudaLaunchKernelEx(&cfg, kern, tmA, tmB, tmC, tmR, p);
    if (e != cudaSuccess) {
      set_error("cudaLaunchKernelEx(gemm_tc_kernel, cluster of %d): %s", CG, cudaGetErrorString(e));
      return LRCE_ECUDA;
    }
  }
  return check_launch("gemm_tc_kernel");
}

// wave efficiency of `tiles` work units on `units` workers
static double wave_eff(int tiles, int units) {
  const int waves = (tiles + units - 1) / units;
  return static_cast<double>(tiles) / (static_cast<double>(waves) * units);
}

template <int BN, int CG>
static int dispatch_epi(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC, const GemmParams& p, int epi,
                        int out_fp32, cudaStream_t stream) {
  const bool lnin = p.in_stats != nullptr;
  if constexpr (CG == 1) {
    if (out_fp32) {
      LRCE_REQUIRE(epi == EPI_BIAS && !lnin, "fp32 output is only available with the plain bias epilogue (epi=%d)", epi);
      return launch_gemm<BN, EPI_BIAS, float, false, 1>(tmA, tmB, tmC, p, stream);
    }
  }
  switch (epi) {
    case EPI_BIAS:
      return lnin ? launch_gemm<BN, EPI_BIAS, bf16, true, CG>(tmA, tmB, tmC, p, stream)
                  : launch_gemm<BN, EPI_BIAS, bf16, false, CG>(tmA, tmB, tmC, p, stream);
    case EPI_BIAS_GELU:
      return lnin ? launch_gemm<BN, EPI_BIAS_GELU, bf16, true, CG>(tmA, tmB, tmC, p, stream)
                  : launch_gemm<BN, EPI_BIAS_GELU, bf16, false, CG>(tmA, tmB, tmC, p, stream);
    case EPI_BIAS_RESIDUAL:
      LRCE_REQUIRE(!lnin, "the residual epilogue does not take a folded LayerNorm input");
      return launch_gemm<BN, EPI_BIAS_RESIDUAL, bf16, false, CG>(tmA, tmB, tmC, p, stream);
    default: break;
  }
  set_error("unknown GEMM epilogue %d", epi);
  return LRCE_EINVAL;
}


// ------------------------------------------------------------------------------------------------------------------------
// Row-tile-fused MLP of a Swin block (stages 2 and 3):  out = x + W2 gelu(W1 LN(x) + b1) + b2  with the hidden rows in L2
// ------------------------------------------------------------------------------------------------------------------------
// fc1 and fc2 as two launches move the hidden activations (M x 4C bf16: 231 MB per stage-3 block) through HBM twice, and
// under sustained load this kernel family runs at the board's power cap, where DRAM traffic is paid in clock. fc2 of a row
// tile needs fc1 of the SAME rows only, so one persistent kernel walks, per CTA pair, a row tile of 256 rows through
//   NT1 = 4C / 256 tiles of fc1 (LayerNorm fold + bias + GELU epilogue)  ->  bf16 hidden rows into the pair's private scratch
//   NT2 =  C / 256 tiles of fc2 (K = 4C, A = the scratch rows the CTA wrote itself; bias + residual + row statistics)
// with the same producer / issuer / epilogue pipeline as gemm_tc_kernel (CTA pairs, 256 x 256 tiles, 2 TMEM accumulators).
// The scratch is 256 x 4C bf16 per pair (74 MB for C = 512 on 74 pairs) and is overwritten by every row tile, so it lives in
// the 126 MB L2: the hidden activations never reach DRAM. No bubble at the fc1 -> fc2 switch: fc2's k-block kb reads hidden
// columns [64 kb, 64 kb + 64) = fc1 tile kb / 4, and only the last four k-blocks need the tile whose epilogue is still running
// when fc2's main loop starts (`h_ready[j]`: 16 epilogue warps -> the CTA's own producer, after cp.async.bulk.wait_group 0).
struct MlpL2Params {
  int M, C;
  const float* in_stats;  // float2 [M][C / in_chunk]: (mean, M2) partials of x (the producer of x emitted them)
  int in_chunk;
  float in_eps;
  int l2_hints;  // 0: none; 1: x / residual loads and out stores evict_first; 2: and the scratch + weights evict_last;
                 // 3: and the hidden rows demoted by their last reader; 4: and discarded (no write-back) once that read has landed
  uint8_t* scratch;
};

template <int BN>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
mlp_l2_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW1,
              const __grid_constant__ CUtensorMap tmHs, const __grid_constant__ CUtensorMap tmHl,
              const __grid_constant__ CUtensorMap tmW2, const __grid_constant__ CUtensorMap tmO,
              const __grid_constant__ CUtensorMap tmR, const GemmParams p1, const GemmParams p2, const MlpL2Params mp) {
  constexpr int CG = 2;
  constexpr int KB_PER_TILE = BN / GEMM_BK;  // k-blocks of fc2 per hidden tile of fc1
  using Cfg = GemmCfg<BN, CG, true>;
  constexpr int STAGES = Cfg::STAGES;
  constexpr int WCOLS = BN / 4, PIECES = WCOLS / 32;
  constexpr int MAX_NT1 = 8;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * Cfg::A_BYTES;
  uint8_t* sC = sB + STAGES * Cfg::B_BYTES;
  uint8_t* sRes = sC + GEMM_EPI_WARPS * Cfg::SLAB_BYTES;
  float2* s_rn = reinterpret_cast<float2*>(sRes + Cfg::RES_BYTES);
  uint64_t* bar_full = reinterpret_cast<uint64_t*>(sRes + Cfg::RES_BYTES + Cfg::RN_BYTES);
  uint64_t* bar_empty = bar_full + STAGES;
  uint64_t* bar_tfull = bar_empty + STAGES;
  uint64_t* bar_tempty = bar_tfull + 2;
  uint64_t* bar_rnfull = bar_tempty + 2;
  uint64_t* bar_rnempty = bar_rnfull + 2;
  uint64_t* bar_resfull = bar_rnempty + 2;
  uint64_t* bar_resempty = bar_resfull + 1;
  uint64_t* h_ready = bar_resempty + 1;  // [MAX_NT1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(h_ready + MAX_NT1);
  static_assert((2 * STAGES + 10 + MAX_NT1) * 8 + 4 <= 256, "barrier block");

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int cta_rank = static_cast<int>(cluster_ctarank());
  const int unit = static_cast<int>(blockIdx.x >> 1);
  const int n_units = static_cast<int>(gridDim.x >> 1);
  const int row_off = cta_rank * GEMM_BM;
  const bool leader = cta_rank == 0;
  const int n_rt = (mp.M + GEMM_BM * CG - 1) / (GEMM_BM * CG);
  const int NT1 = 4 * mp.C / BN, NT2 = mp.C / BN;
  const int nkb1 = mp.C / GEMM_BK, nkb2 = 4 * mp.C / GEMM_BK;
  const int h0 = unit * (GEMM_BM * CG) + row_off;  // this CTA's rows of the scratch

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmW1);
    tma_prefetch_desc(&tmHs);
    tma_prefetch_desc(&tmHl);
    tma_prefetch_desc(&tmW2);
    tma_prefetch_desc(&tmO);
    tma_prefetch_desc(&tmR);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&bar_full[s], 1);
      mbar_init(&bar_empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&bar_tfull[a], 1);
      mbar_init(&bar_tempty[a], GEMM_EPI_WARPS * CG);
      mbar_init(&bar_rnfull[a], 1);
      mbar_init(&bar_rnempty[a], GEMM_EPI_WARPS);
    }
    mbar_init(bar_resfull, 1);
    mbar_init(bar_resempty, GEMM_EPI_WARPS);
    for (int j = 0; j < MAX_NT1; ++j) mbar_init(&h_ready[j], GEMM_EPI_WARPS);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc_pair(tmem_slot, Cfg::TMEM_COLS);
  tcgen05_fence_before();
  cluster_sync_all();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  griddep_wait();
  griddep_launch();

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0, rit = 0, rt_it = 0;
      uint32_t phase = 0;
      // the scratch (and the weights every pair re-reads) must survive in L2 next to the x / out streams of the whole chip
      // (x itself is re-read by the eight fc1 tiles and once more as the residual: normal priority until that last read)
      const uint64_t pol_stream = l2_policy(mp.l2_hints >= 1 ? 1 : 0), pol_keep = l2_policy(mp.l2_hints >= 2 ? 2 : 0), pol_normal = l2_policy(0);
      auto load_kb = [&](const CUtensorMap* ta, int a_row, uint64_t pol_a, const CUtensorMap* tb, int b_row, int kb) {
        mbar_wait_parked(&bar_empty[stage], phase ^ 1);
        if (leader) mbar_expect_tx(&bar_full[stage], 2 * (Cfg::A_BYTES + Cfg::B_BYTES));
        tma_load_2d_pair_hint(sA + stage * Cfg::A_BYTES, ta, &bar_full[stage], kb * GEMM_BK, a_row, pol_a);
        tma_load_2d_pair_hint(sB + stage * Cfg::B_BYTES, tb, &bar_full[stage], kb * GEMM_BK, b_row, pol_keep);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      };
      for (int rt = unit; rt < n_rt; rt += n_units, ++rt_it) {
        const int m0 = rt * (GEMM_BM * CG) + row_off;
        for (int s = 0; s < NT1; ++s)
          for (int kb = 0; kb < nkb1; ++kb) load_kb(&tmX, m0, pol_normal, &tmW1, s * BN + cta_rank * (BN / 2), kb);
        for (int s = 0; s < NT2; ++s) {
          for (int kb = 0; kb < nkb2; ++kb) {
            if (s == 0 && kb % KB_PER_TILE == 0) {
              // hidden columns [64 kb, 64 kb + BN) = fc1 tile kb / KB_PER_TILE of this row tile: stored and complete?
              mbar_wait_parked(&h_ready[kb / KB_PER_TILE], rt_it & 1);
              asm volatile("fence.proxy.async;" ::: "memory");
            }
            // the last fc2 tile is the last reader of the hidden rows: demote them, so that what the L2 evicts next are these dead
            // lines (rewritten by the next row tile anyway) rather than hidden rows that are still waiting to be read
            load_kb(&tmHl, h0, (mp.l2_hints >= 3 && s + 1 == NT2) ? pol_stream : pol_keep, &tmW2, s * BN + cta_rank * (BN / 2), kb);
          }
          // the residual rows of this fc2 tile (= x, not yet overwritten: the tile's own epilogue does that), a main loop ahead
          mbar_wait_parked(bar_resempty, (rit & 1) ^ 1);
          mbar_expect_tx(bar_resfull, Cfg::RES_BYTES);
#pragma unroll
          for (int pc = 0; pc < BN / 32; ++pc)
            tma_load_2d_hint(sRes + pc * (GEMM_BM * 64), &tmR, bar_resfull, s * BN + pc * 32, m0, pol_stream);
          ++rit;
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (the pair's leader; elected lane)
    if (leader) {
      constexpr uint32_t idesc = umma_idesc_bf16(GEMM_BM * CG, BN);
      const uint32_t a_lo0 = desc_lo(smem_u32(sA)), b_lo0 = desc_lo(smem_u32(sB));
      int stage = 0, it = 0;
      uint32_t phase = 0;
      uint8_t* h_pair = mp.scratch + static_cast<size_t>(unit) * (GEMM_BM * CG) * (8 * mp.C);  // the pair's 256 scratch rows (4C bf16 each)
      for (int rt = unit; rt < n_rt; rt += n_units) {
        for (int s = 0; s < NT1 + NT2; ++s, ++it) {
          const int n_kb = s < NT1 ? nkb1 : nkb2;
          const int as = it & 1;
          mbar_wait_parked(&bar_tempty[as], ((it >> 1) & 1) ^ 1);
          tcgen05_fence_after();
          const uint32_t tmem_d = tmem_base + as * BN;
          for (int kb = 0; kb < n_kb; ++kb) {
            mbar_wait_parked(&bar_full[stage], phase);
            tcgen05_fence_after();
            const uint32_t a_lo = a_lo0 + stage * (Cfg::A_BYTES >> 4), b_lo = b_lo0 + stage * (Cfg::B_BYTES >> 4);
            if (elect_one()) {
#pragma unroll
              for (int k = 0; k < GEMM_BK / 16; ++k)
                umma_lo_acc<CG>(tmem_d, a_lo + k * 2, b_lo + k * 2, idesc, k != 0 ? 1u : static_cast<uint32_t>(kb));
              umma_commit_pair(&bar_empty[stage]);
            }
            __syncwarp();
            if (mp.l2_hints >= 4 && s + 1 == NT1 + NT2) {
              // the last fc2 tile's copy of hidden columns [64 kb, 64 kb + 64) has landed in both CTAs: those 256 x 128 B of the
              // scratch are dead until the next row tile rewrites them -> let the L2 drop them without a write-back
#pragma unroll
              for (int r = 0; r < (GEMM_BM * CG) / 32; ++r)
                asm volatile("discard.global.L2 [%0], 128;" ::"l"(h_pair + static_cast<size_t>(r * 32 + lane) * (8 * mp.C) + kb * 128) : "memory");
            }
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
          if (elect_one()) umma_commit_pair(&bar_tfull[as]);
          __syncwarp();
        }
      }
    }
  } else if (warp == 3) {
    // ------------------------------------------------------------------ row statistics of LN(x), once per row tile
    const int n_chunks = mp.C / mp.in_chunk;
    const float cw = static_cast<float>(mp.in_chunk);
    int rt_it = 0;
    for (int rt = unit; rt < n_rt; rt += n_units, ++rt_it) {
      const int m0 = rt * (GEMM_BM * CG) + row_off;
      float2* dst = s_rn + (rt_it & 1) * GEMM_BM;
      if (n_chunks == 4) {
        float2 ta[4][4];
        stats_fetch<4, 4>(ta, mp.in_stats, mp.M, m0, lane, 0);
        mbar_wait_parked(&bar_rnempty[rt_it & 1], ((rt_it >> 1) & 1) ^ 1);
        stats_reduce<4, 4>(ta, dst, lane, 0, cw, mp.C, mp.in_eps);
      } else {  // 8 chunks (host-checked)
        float2 ta[4][8];
        stats_fetch<8, 4>(ta, mp.in_stats, mp.M, m0, lane, 0);
        mbar_wait_parked(&bar_rnempty[rt_it & 1], ((rt_it >> 1) & 1) ^ 1);
        stats_reduce<8, 4>(ta, dst, lane, 0, cw, mp.C, mp.in_eps);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar_rnfull[rt_it & 1]);
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ epilogue
    const int q = warp & 3;
    const int part = (warp - 4) >> 2;
    const int row_in_tile = q * 32 + lane;
    uint8_t* slab = sC + (warp - 4) * Cfg::SLAB_BYTES;
    const uint4 no_res[4] = {make_uint4(0, 0, 0, 0), make_uint4(0, 0, 0, 0), make_uint4(0, 0, 0, 0), make_uint4(0, 0, 0, 0)};
    int it = 0, rit = 0, rt_it = 0;
    const uint64_t pol_stream = l2_policy(mp.l2_hints >= 1 ? 1 : 0), pol_keep = l2_policy(mp.l2_hints >= 2 ? 2 : 0);
    auto store_piece = [&](const uint4 (&o)[4], const CUtensorMap* tm, int col, int row0, uint64_t pol) {
      // the previous TMA store of this warp must have finished reading the slab before it is refilled
      if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      __syncwarp();
#pragma unroll
      for (int i = 0; i < 4; ++i) *reinterpret_cast<uint4*>(slab + lane * 64 + ((i ^ ((lane >> 1) & 3)) << 4)) = o[i];
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group.L2::cache_hint [%0, {%2, %3}], [%1], %4;" ::"l"(
                         reinterpret_cast<uint64_t>(tm)),
                     "r"(smem_u32(slab)), "r"(col), "r"(row0), "l"(pol)
                     : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
    };
    for (int rt = unit; rt < n_rt; rt += n_units, ++rt_it) {
      const int m0 = rt * (GEMM_BM * CG) + row_off;
      const int row = m0 + row_in_tile;
      mbar_wait_parked(&bar_rnfull[rt_it & 1], (rt_it >> 1) & 1);
      const float2 rn = s_rn[(rt_it & 1) * GEMM_BM + row_in_tile];
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar_rnempty[rt_it & 1]);
      // ---- fc1 tiles: hidden = gelu(rstd (acc - mean colsum) + b1') -> scratch rows h0 + ...
      for (int s = 0; s < NT1; ++s, ++it) {
        const int as = it & 1;
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * BN + part * WCOLS;
        mbar_wait_parked(&bar_tfull[as], (it >> 1) & 1);
        tcgen05_fence_after();
#pragma unroll
        for (int pc = 0; pc < PIECES; ++pc) {
          const int col_in_tile = part * WCOLS + pc * 32;
          uint32_t acc[32];
          tmem_ld_32x32(taddr + pc * 32, acc);
          tmem_ld_wait();
          if (pc + 1 == PIECES) {
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_leader(&bar_tempty[as]);
          }
          float v[32];
          epilogue_values<EPI_BIAS_GELU, true>(acc, s * BN + col_in_tile, p1, rn, no_res, v);
          uint4 o[4];
          pack32(v, o);
          store_piece(o, &tmHs, s * BN + col_in_tile, h0 + q * 32, pol_keep);
        }
        // this warp's part of a hidden tile is in L2 once its stores have COMPLETED (not merely been read out of the slab);
        // bulk groups complete in order, so tile s - 1 is reported after tile s has been issued: the completion latency of the
        // stores hides behind the next tile instead of stalling this warp before it
        if (s > 0 && lane == 0) {
          asm volatile("cp.async.bulk.wait_group %0;" ::"n"(PIECES) : "memory");
          mbar_arrive(&h_ready[s - 1]);
        }
        __syncwarp();
      }
      if (lane == 0) {
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        mbar_arrive(&h_ready[NT1 - 1]);
      }
      __syncwarp();
      // ---- fc2 tiles: out = acc + b2 + x, row statistics of the result for the next block's folded LayerNorm
      for (int s = 0; s < NT2; ++s, ++it, ++rit) {
        const int as = it & 1;
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * BN + part * WCOLS;
        mbar_wait_parked(bar_resfull, rit & 1);
        mbar_wait_parked(&bar_tfull[as], (it >> 1) & 1);
        tcgen05_fence_after();
        float2 st_carry = make_float2(0.f, 0.f);
#pragma unroll
        for (int pc = 0; pc < PIECES; ++pc) {
          const int col_in_tile = part * WCOLS + pc * 32;
          uint32_t acc[32];
          tmem_ld_32x32(taddr + pc * 32, acc);
          uint4 res_cur[4];
          const uint8_t* rp = sRes + (col_in_tile >> 5) * (GEMM_BM * 64) + row_in_tile * 64;
#pragma unroll
          for (int i = 0; i < 4; ++i) res_cur[i] = *reinterpret_cast<const uint4*>(rp + ((i ^ ((row_in_tile >> 1) & 3)) << 4));
          if (pc + 1 == PIECES) {
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_resempty);
          }
          tmem_ld_wait();
          if (pc + 1 == PIECES) {
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_leader(&bar_tempty[as]);
          }
          float v[32];
          epilogue_values<EPI_BIAS_RESIDUAL, false>(acc, s * BN + col_in_tile, p2, rn, res_cur, v);
          uint4 o[4];
          pack32(v, o);
          if (p2.out_stats != nullptr) {
            const float2 st_piece = stats32(v);
            if (PIECES == 1) {  // 128-wide tiles: 32-column chunks
              if (row < mp.M)
                reinterpret_cast<float2*>(p2.out_stats)[static_cast<size_t>(row) * (mp.C >> 5) + ((s * BN + col_in_tile) >> 5)] = st_piece;
            } else if (pc == 0) {
              st_carry = st_piece;
            } else if (row < mp.M) {  // 256-wide tiles: this warp's two pieces form one 64-column chunk
              reinterpret_cast<float2*>(p2.out_stats)[static_cast<size_t>(row) * (mp.C >> 6) + ((s * BN + part * WCOLS) >> 6)] =
                  stats_merge(st_carry, st_piece, 32.0f);
            }
          }
          store_piece(o, &tmO, s * BN + col_in_tile, m0 + q * 32, pol_stream);
        }
      }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }

  tcgen05_fence_before();
  cluster_sync_all();
  if (warp == 2) {
    tcgen05_fence_after();
    tmem_dealloc_pair(tmem_base, Cfg::TMEM_COLS);
  }
}

}  // namespace lrce

using namespace lrce;

extern "C" int lrce_gemm_bf16(const void* A, int lda, const void* W, int ldw, int M, int N, int K, const float* bias,
                              const void* residual, int ldr, void* out, int ldo, int epilogue, int out_fp32,
                              const float* ln_gamma, const float* ln_beta, float ln_eps, const float* in_stats,
                              int in_chunk, const float* in_colsum, float in_eps, float* out_stats, void* stream) {
  int rc = require_sm100();
  if (rc != LRCE_OK) return rc;
  LRCE_REQUIRE(A && W && out, "lrce_gemm_bf16: null operand");
  LRCE_REQUIRE(M > 0 && N > 0 && K > 0, "lrce_gemm_bf16: bad shape M=%d N=%d K=%d", M, N, K);
  LRCE_REQUIRE(N % 128 == 0, "lrce_gemm_bf16: N=%d must be a multiple of 128", N);
  LRCE_REQUIRE(K % 8 == 0 && lda % 8 == 0 && ldw % 8 == 0, "lrce_gemm_bf16: K/lda/ldw must be multiples of 8");
  LRCE_REQUIRE(ldo % 8 == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0, "lrce_gemm_bf16: out must be 16B aligned");
  LRCE_REQUIRE(bias == nullptr || (reinterpret_cast<uintptr_t>(bias) & 15) == 0, "lrce_gemm_bf16: bias must be 16B aligned");
  if (epilogue == EPI_BIAS_RESIDUAL)
    LRCE_REQUIRE(residual && ldr % 8 == 0 && (reinterpret_cast<uintptr_t>(residual) & 15) == 0,
                 "lrce_gemm_bf16: residual epilogue needs a 16B-aligned residual");
  if (in_stats != nullptr)
    LRCE_REQUIRE(in_colsum && bias && (in_chunk == 32 || in_chunk == 64) && K % in_chunk == 0 &&
                     (K / in_chunk == 4 || K / in_chunk == 8 || K / in_chunk == 16) && (reinterpret_cast<uintptr_t>(in_colsum) & 15) == 0 &&
                     (reinterpret_cast<uintptr_t>(in_stats) & 15) == 0,
                 "lrce_gemm_bf16: a folded LayerNorm input needs statistics in 4, 8 or 16 chunks of 32 or 64 columns, column sums and a bias (K=%d, chunk=%d)", K, in_chunk);
  if (out_stats != nullptr)
    LRCE_REQUIRE(!out_fp32 && (reinterpret_cast<uintptr_t>(out_stats) & 7) == 0,
                 "lrce_gemm_bf16: row statistics are emitted for bf16 outputs only");
  GemmParams p;
  p.M = M; p.N = N; p.K = K;
  p.bias = bias;
  p.residual = reinterpret_cast<const bf16*>(residual);
  p.ldr = ldr;
  p.out = out;
  p.ldo = ldo;
  p.ln_g = ln_gamma; p.ln_b = ln_beta; p.ln_eps = ln_eps;
  p.in_stats = in_stats; p.in_chunk = in_chunk; p.in_colsum = in_colsum; p.in_eps = in_eps;
  p.out_stats = out_stats;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  CUtensorMap tmA, tmB, tmC;
  rc = make_tmap_2d_bf16(&tmA, A, K, M, lda, GEMM_BK, GEMM_BM);
  if (rc != LRCE_OK) return rc;
  if (!out_fp32 && epilogue != EPI_BIAS_LN) {  // bf16 tiles leave as per-warp TMA stores of 32-row x 32-column slabs
    rc = make_tmap_2d_bf16(&tmC, out, N, M, ldo, 32, 32, /*swizzle_bytes=*/64);
    if (rc != LRCE_OK) return rc;
  } else {
    tmC = tmA;  // unused by these epilogues
  }
  if (epilogue == EPI_BIAS_LN) {
    LRCE_REQUIRE(N == 128 && bias && ln_gamma && ln_beta && !out_fp32 && !in_stats,
                 "lrce_gemm_bf16: the LayerNorm epilogue needs N == 128, a bias and gamma/beta (N=%d)", N);
    rc = make_tmap_2d_bf16(&tmB, W, K, N, ldw, GEMM_BK, 128);
    if (rc != LRCE_OK) return rc;
    return launch_gemm<128, EPI_BIAS_LN, bf16, false, 1>(tmA, tmB, tmC, p, s);
  }
  const bool wide = (N % 256 == 0);
  // CTA pairs (256 x 256 tiles on two SMs) for the compute-bound shapes, unless their wave quantisation is clearly worse
  // than that of single-CTA 128 x 256 tiles
  const int sms = sm_count();
  const int tiles1 = ((M + 127) / 128) * (N / 256), tiles2 = ((M + 255) / 256) * (N / 256);
  const bool pair = wide && !out_fp32 && K >= 256 && sms >= 2 && wave_eff(tiles2, sms / 2) >= wave_eff(tiles1, sms) - 0.02;
  rc = make_tmap_2d_bf16(&tmB, W, K, N, ldw, GEMM_BK, wide ? (pair ? 128 : 256) : 128);
  if (rc != LRCE_OK) return rc;
  if (pair) return dispatch_epi<256, 2>(tmA, tmB, tmC, p, epilogue, out_fp32, s);
  return wide ? dispatch_epi<256, 1>(tmA, tmB, tmC, p, epilogue, out_fp32, s) : dispatch_epi<128, 1>(tmA, tmB, tmC, p, epilogue, out_fp32, s);
}

extern "C" size_t lrce_mlp_l2_scratch_bytes(int C) {
  const int units = sm_count() / 2;
  return static_cast<size_t>(units > 0 ? units : 1) * 256 * 4 * static_cast<size_t>(C > 0 ? C : 0) * 2;
}

extern "C" int lrce_mlp_l2_bf16(const void* x, int ldx, const void* w1, const float* b1, const float* colsum1,
                                const float* in_stats, int in_chunk, float in_eps, const void* w2, const float* b2, void* out,
                                int ldo, float* out_stats, void* scratch, size_t scratch_bytes, int M, int C, void* stream) {
  int rc = require_sm100();
  if (rc != LRCE_OK) return rc;
  LRCE_REQUIRE(x && w1 && b1 && colsum1 && in_stats && w2 && b2 && out && scratch, "lrce_mlp_l2_bf16: null operand");
  LRCE_REQUIRE(M > 0 && (C == 128 || C == 256 || C == 512), "lrce_mlp_l2_bf16: M=%d, C=%d (C must be 128, 256 or 512)", M, C);
  LRCE_REQUIRE(in_chunk == (C == 128 ? 32 : 64), "lrce_mlp_l2_bf16: statistics chunks of %d columns for C=%d", in_chunk, C);
  LRCE_REQUIRE(ldx % 8 == 0 && ldo % 8 == 0 && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out) |
                                                 reinterpret_cast<uintptr_t>(scratch)) & 15) == 0,
               "lrce_mlp_l2_bf16: x / out / scratch must be 16B aligned with row pitches that are multiples of 8");
  LRCE_REQUIRE(((reinterpret_cast<uintptr_t>(b1) | reinterpret_cast<uintptr_t>(colsum1) | reinterpret_cast<uintptr_t>(b2) |
                 reinterpret_cast<uintptr_t>(in_stats)) & 15) == 0 && (out_stats == nullptr || (reinterpret_cast<uintptr_t>(out_stats) & 7) == 0),
               "lrce_mlp_l2_bf16: bias / column sums / statistics must be 16B aligned");
  const int sms = sm_count();
  LRCE_REQUIRE(sms >= 2, "lrce_mlp_l2_bf16: needs CTA pairs");
  const int n_rt = (M + 255) / 256;
  int units = sms / 2;
  if (units > n_rt) units = n_rt;
  LRCE_REQUIRE(scratch_bytes >= static_cast<size_t>(units) * 256 * 4 * C * 2,
               "lrce_mlp_l2_bf16: scratch of %zu bytes, %zu needed (lrce_mlp_l2_scratch_bytes)", scratch_bytes,
               static_cast<size_t>(units) * 256 * 4 * C * 2);
  const int H = 4 * C;
  CUtensorMap tmX, tmW1, tmHs, tmHl, tmW2, tmO, tmR;
  if ((rc = make_tmap_2d_bf16(&tmX, x, C, M, ldx, GEMM_BK, GEMM_BM)) != LRCE_OK) return rc;
  const int BN = C == 128 ? 128 : 256;  // tile width of both products (fc2 has N = C columns)
  if ((rc = make_tmap_2d_bf16(&tmW1, w1, C, H, C, GEMM_BK, BN / 2)) != LRCE_OK) return rc;
  if ((rc = make_tmap_2d_bf16(&tmHs, scratch, H, static_cast<uint64_t>(units) * 256, H, 32, 32, 64)) != LRCE_OK) return rc;
  if ((rc = make_tmap_2d_bf16(&tmHl, scratch, H, static_cast<uint64_t>(units) * 256, H, GEMM_BK, GEMM_BM)) != LRCE_OK) return rc;
  if ((rc = make_tmap_2d_bf16(&tmW2, w2, H, C, H, GEMM_BK, BN / 2)) != LRCE_OK) return rc;
  if ((rc = make_tmap_2d_bf16(&tmO, out, C, M, ldo, 32, 32, 64)) != LRCE_OK) return rc;
  if ((rc = make_tmap_2d_bf16(&tmR, x, C, M, ldx, 32, GEMM_BM, 64)) != LRCE_OK) return rc;
  GemmParams p1 = {}, p2 = {};
  p1.M = M; p1.N = H; p1.K = C; p1.bias = b1; p1.in_stats = in_stats; p1.in_chunk = in_chunk; p1.in_colsum = colsum1; p1.in_eps = in_eps;
  p2.M = M; p2.N = C; p2.K = H; p2.bias = b2; p2.out = out; p2.ldo = ldo; p2.out_stats = out_stats;
  MlpL2Params mp;
  mp.M = M; mp.C = C; mp.in_stats = in_stats; mp.in_chunk = in_chunk; mp.in_eps = in_eps;
  static const int hints = [] {
    const char* e = getenv("LRCE_B200_MLP_L2_HINTS");  // A/B runs of tools/
    return e ? atoi(e) : 4;
  }();
  mp.l2_hints = hints;
  mp.scratch = static_cast<uint8_t*>(scratch);
  auto kern = BN == 128 ? mlp_l2_kernel<128> : mlp_l2_kernel<256>;
  const int smem_bytes = BN == 128 ? GemmCfg<128, 2, true>::SMEM_BYTES : GemmCfg<256, 2, true>::SMEM_BYTES;
  static thread_local uint64_t configured = 0;
  if (needs_device_setup(&configured)) {
    cudaError_t e = cudaFuncSetAttribute(mlp_l2_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, GemmCfg<128, 2, true>::SMEM_BYTES);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(mlp_l2_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, GemmCfg<256, 2, true>::SMEM_BYTES);
    if (e != cudaSuccess) {
      set_error("cudaFuncSetAttribute(mlp_l2_kernel): %s", cudaGetErrorString(e));
      return LRCE_ECUDA;
    }
    mark_device_setup(&configured);
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(units * 2);
  cfg.blockDim = dim3(GEMM_THREADS);
  cfg.dynamicSmemBytes = smem_bytes;
  cfg.stream = reinterpret_cast<cudaStream_t>(stream);
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1 + pdl_attr(attr + 1);
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, tmX, tmW1, tmHs, tmHl, tmW2, tmO, tmR, p1, p2, mp);
  if (e != cudaSuccess) {
    set_error("cudaLaunchKernelEx(mlp_l2_kernel): %s", cudaGetErrorString(e));
    return LRCE_ECUDA;
  }
  return check_launch("mlp_l2_kernel");
}
