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
//                               LayerNorm. Two warps share each 32-lane TMEM quarter and split the columns, so the
//                               erf-heavy epilogues keep up with the tensor pipe. bf16 tiles leave through two
//                               128x64 smem slabs (128B-swizzled) and TMA bulk stores, so global writes are full
//                               lines and the M tail is clipped by the descriptor.
// Both operands are K-major, so no transposes exist anywhere; M and K tails are handled by TMA zero fill.
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
};

constexpr int GEMM_BM = 128;
constexpr int GEMM_BK = 64;
constexpr int GEMM_THREADS = 384;
constexpr int GEMM_EPI_WARPS = 8;

template <int BN>
struct GemmCfg {
  static constexpr int STAGES = (BN == 256) ? 4 : 6;
  static constexpr int A_BYTES = GEMM_BM * GEMM_BK * 2;
  static constexpr int B_BYTES = BN * GEMM_BK * 2;
  static constexpr int TMEM_COLS = 2 * BN;  // 256 or 512: power of two
  static constexpr int SLAB_BYTES = GEMM_BM * 64 * 2;  // one 128 x 64 bf16 output slab
  static constexpr int SMEM_BYTES = STAGES * (A_BYTES + B_BYTES) + 2 * SLAB_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
};

// bias / GELU / residual on 32 accumulator columns of one row; result packed to 4 x 16 B of bf16
template <int EPI>
__device__ __forceinline__ void epilogue_math(const uint32_t (&acc)[32], int col0, const GemmParams& p, const uint4 (&res)[4],
                                              uint4 (&o)[4]) {
  float v[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(acc[i]);
  if (p.bias != nullptr) {
    const float4* b4 = reinterpret_cast<const float4*>(p.bias + col0);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float4 b = __ldg(b4 + i);
      v[4 * i + 0] += b.x;
      v[4 * i + 1] += b.y;
      v[4 * i + 2] += b.z;
      v[4 * i + 3] += b.w;
    }
  }
  if (EPI == EPI_BIAS_GELU) {
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = gelu_erf(v[i]);
  }
  if (EPI == EPI_BIAS_RESIDUAL) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float2 f;
      f = unpack_bf16x2(res[i].x); v[8 * i + 0] += f.x; v[8 * i + 1] += f.y;
      f = unpack_bf16x2(res[i].y); v[8 * i + 2] += f.x; v[8 * i + 3] += f.y;
      f = unpack_bf16x2(res[i].z); v[8 * i + 4] += f.x; v[8 * i + 5] += f.y;
      f = unpack_bf16x2(res[i].w); v[8 * i + 6] += f.x; v[8 * i + 7] += f.y;
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    o[i].x = pack_bf16x2(v[8 * i + 0], v[8 * i + 1]);
    o[i].y = pack_bf16x2(v[8 * i + 2], v[8 * i + 3]);
    o[i].z = pack_bf16x2(v[8 * i + 4], v[8 * i + 5]);
    o[i].w = pack_bf16x2(v[8 * i + 6], v[8 * i + 7]);
  }
}

template <int EPI, typename OutT>
__device__ __forceinline__ void epilogue_chunk(const uint32_t (&acc)[32], int row, int col0, const GemmParams& p) {
  float v[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(acc[i]);
  if (p.bias != nullptr) {
    const float4* b4 = reinterpret_cast<const float4*>(p.bias + col0);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float4 b = __ldg(b4 + i);
      v[4 * i + 0] += b.x;
      v[4 * i + 1] += b.y;
      v[4 * i + 2] += b.z;
      v[4 * i + 3] += b.w;
    }
  }
  if (EPI == EPI_BIAS_GELU) {
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = gelu_erf(v[i]);
  }
  if (row >= p.M) return;
  if (EPI == EPI_BIAS_RESIDUAL) {
    const uint4* r4 = reinterpret_cast<const uint4*>(p.residual + static_cast<size_t>(row) * p.ldr + col0);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      uint4 r = r4[i];
      float2 f;
      f = unpack_bf16x2(r.x); v[8 * i + 0] += f.x; v[8 * i + 1] += f.y;
      f = unpack_bf16x2(r.y); v[8 * i + 2] += f.x; v[8 * i + 3] += f.y;
      f = unpack_bf16x2(r.z); v[8 * i + 4] += f.x; v[8 * i + 5] += f.y;
      f = unpack_bf16x2(r.w); v[8 * i + 6] += f.x; v[8 * i + 7] += f.y;
    }
  }
  if (sizeof(OutT) == 2) {
    uint4* o4 = reinterpret_cast<uint4*>(reinterpret_cast<bf16*>(p.out) + static_cast<size_t>(row) * p.ldo + col0);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      uint4 o;
      o.x = pack_bf16x2(v[8 * i + 0], v[8 * i + 1]);
      o.y = pack_bf16x2(v[8 * i + 2], v[8 * i + 3]);
      o.z = pack_bf16x2(v[8 * i + 4], v[8 * i + 5]);
      o.w = pack_bf16x2(v[8 * i + 6], v[8 * i + 7]);
      o4[i] = o;
    }
  } else {
    float4* o4 = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + static_cast<size_t>(row) * p.ldo + col0);
#pragma unroll
    for (int i = 0; i < 8; ++i) o4[i] = make_float4(v[4 * i + 0], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
  }
}

template <int BN, int EPI, typename OutT>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmC, const GemmParams p) {
  constexpr bool TMA_STORE = (sizeof(OutT) == 2) && (EPI != EPI_BIAS_LN);
  using Cfg = GemmCfg<BN>;
  constexpr int STAGES = Cfg::STAGES;
  // 1024-byte alignment is required by the 128B swizzle atoms; the pointer stays derived from the __shared__ array so
  // that the epilogue's slab writes compile to STS (an integer round trip would demote them to generic stores)
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * Cfg::A_BYTES;
  uint8_t* sC = sB + STAGES * Cfg::B_BYTES;  // 2 output slabs, 1024-byte aligned (stage sizes are multiples of 1024)
  uint64_t* bar_full = reinterpret_cast<uint64_t*>(sC + 2 * Cfg::SLAB_BYTES);
  uint64_t* bar_empty = bar_full + STAGES;
  uint64_t* bar_tfull = bar_empty + STAGES;
  uint64_t* bar_tempty = bar_tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_tempty + 2);
  // LayerNorm epilogue scratch [tile parity][column half][sum | sumsq][row] aliases the (then unused) output slabs
  float* ln_part = reinterpret_cast<float*>(sC);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n_tiles_n = p.N / BN;
  const int n_tiles_m = (p.M + GEMM_BM - 1) / GEMM_BM;
  const int n_tiles = n_tiles_m * n_tiles_n;
  const int n_kb = (p.K + GEMM_BK - 1) / GEMM_BK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (TMA_STORE) tma_prefetch_desc(&tmC);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&bar_full[s], 1);
      mbar_init(&bar_empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&bar_tfull[a], 1);
      mbar_init(&bar_tempty[a], GEMM_EPI_WARPS);
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        const int m0 = (t / n_tiles_n) * GEMM_BM;
        const int n0 = (t % n_tiles_n) * BN;
        for (int kb = 0; kb < n_kb; ++kb) {
          mbar_wait(&bar_empty[stage], phase ^ 1);
          mbar_expect_tx(&bar_full[stage], Cfg::A_BYTES + Cfg::B_BYTES);
          tma_load_2d(sA + stage * Cfg::A_BYTES, &tmA, &bar_full[stage], kb * GEMM_BK, m0);
          tma_load_2d(sB + stage * Cfg::B_BYTES, &tmB, &bar_full[stage], kb * GEMM_BK, n0);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (single thread)
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(GEMM_BM, BN);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it) {
        const int as = it & 1;
        const uint32_t aphase = (it >> 1) & 1;
        mbar_wait(&bar_tempty[as], aphase ^ 1);
        tcgen05_fence_after();
        const uint32_t tmem_d = tmem_base + as * BN;
        for (int kb = 0; kb < n_kb; ++kb) {
          mbar_wait(&bar_full[stage], phase);
          tcgen05_fence_after();
          const uint32_t a_addr = smem_u32(sA + stage * Cfg::A_BYTES);
          const uint32_t b_addr = smem_u32(sB + stage * Cfg::B_BYTES);
#pragma unroll
          for (int k = 0; k < GEMM_BK / 16; ++k) {
            umma_bf16_ss(tmem_d, umma_desc_k_sw128(a_addr + k * 32), umma_desc_k_sw128(b_addr + k * 32), idesc,
                         (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit(&bar_empty[stage]);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(&bar_tfull[as]);
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ epilogue
    const int q = warp & 3;          // TMEM lane quarter this warp may access
    const int half = (warp - 4) >> 2;  // which half of the BN columns
    constexpr int COLS = BN / 2;
    int it = 0;
    uint32_t slab_count = 0;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it) {
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      const int m0 = (t / n_tiles_n) * GEMM_BM;
      const int n0 = (t % n_tiles_n) * BN;
      const int row_in_tile = q * 32 + lane;
      const int row = m0 + row_in_tile;
      // residual rows are fetched one 64-column slab ahead (the first one before the accumulator is even ready), so
      // their latency hides behind the MMA wait / the previous slab's math
      uint4 res_next[4] = {make_uint4(0, 0, 0, 0), make_uint4(0, 0, 0, 0), make_uint4(0, 0, 0, 0), make_uint4(0, 0, 0, 0)};
      const bf16* res_row = nullptr;
      if (TMA_STORE && EPI == EPI_BIAS_RESIDUAL && row < p.M) {
        res_row = p.residual + static_cast<size_t>(row) * p.ldr + n0 + half * 32;
#pragma unroll
        for (int i = 0; i < 4; ++i) res_next[i] = reinterpret_cast<const uint4*>(res_row)[i];
      }
      mbar_wait(&bar_tfull[as], aphase);
      tcgen05_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * BN + half * COLS;
      if (EPI == EPI_BIAS_LN) {
        // full-row LayerNorm: BN == N == 128, this thread holds 64 of the row's 128 features
        float v[COLS];
#pragma unroll
        for (int c = 0; c < COLS; c += 32) {
          uint32_t acc[32];
          tmem_ld_32x32(taddr + c, acc);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) v[c + i] = __uint_as_float(acc[i]) + __ldg(p.bias + half * COLS + c + i);
        }
        float s = 0.f, ss = 0.f;
#pragma unroll
        for (int i = 0; i < COLS; ++i) { s += v[i]; ss += v[i] * v[i]; }
        const int par = it & 1;
        ln_part[((par * 2 + half) * 2 + 0) * GEMM_BM + row_in_tile] = s;
        ln_part[((par * 2 + half) * 2 + 1) * GEMM_BM + row_in_tile] = ss;
        asm volatile("bar.sync 1, 256;" ::: "memory");
        s += ln_part[((par * 2 + (half ^ 1)) * 2 + 0) * GEMM_BM + row_in_tile];
        ss += ln_part[((par * 2 + (half ^ 1)) * 2 + 1) * GEMM_BM + row_in_tile];
        const float mean = s * (1.0f / BN);
        const float var = fmaxf(ss * (1.0f / BN) - mean * mean, 0.f);
        const float rstd = rsqrtf(var + p.ln_eps);
        if (row < p.M) {
          bf16* o = reinterpret_cast<bf16*>(p.out) + static_cast<size_t>(row) * p.ldo + half * COLS;
#pragma unroll
          for (int i = 0; i < COLS; i += 8) {
            float y[8];
#pragma unroll
            for (int j = 0; j < 8; ++j)
              y[j] = (v[i + j] - mean) * rstd * __ldg(p.ln_g + half * COLS + i + j) + __ldg(p.ln_b + half * COLS + i + j);
            uint4 o4;
            o4.x = pack_bf16x2(y[0], y[1]);
            o4.y = pack_bf16x2(y[2], y[3]);
            o4.z = pack_bf16x2(y[4], y[5]);
            o4.w = pack_bf16x2(y[6], y[7]);
            *reinterpret_cast<uint4*>(o + i) = o4;
          }
        }
      } else if (TMA_STORE) {
        // 64-column slabs: the 8 epilogue warps fill one 128 x 64 bf16 slab (this warp: 32 rows x 32 columns), then one
        // thread hands it to the TMA store engine; two slabs alternate so the store of slab g overlaps the math of g+1.
        const bool leader = (warp == 4 && lane == 0);
#pragma unroll 1
        for (int sl = 0; sl < BN / 64; ++sl, ++slab_count) {
          uint32_t acc[32];
          tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * BN + sl * 64 + half * 32, acc);
          uint4 res_cur[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) res_cur[i] = res_next[i];
          if (EPI == EPI_BIAS_RESIDUAL && res_row != nullptr && sl + 1 < BN / 64) {
#pragma unroll
            for (int i = 0; i < 4; ++i) res_next[i] = reinterpret_cast<const uint4*>(res_row + (sl + 1) * 64)[i];
          }
          tmem_ld_wait();
          uint4 o[4];
          epilogue_math<EPI>(acc, n0 + sl * 64 + half * 32, p, res_cur, o);
          uint8_t* slab = sC + (slab_count & 1) * Cfg::SLAB_BYTES + row_in_tile * 128;
#pragma unroll
          for (int i = 0; i < 4; ++i)  // 128B swizzle: 16-byte chunk c of row r lives at chunk position c ^ (r & 7)
            *reinterpret_cast<uint4*>(slab + (((half * 4 + i) ^ (row_in_tile & 7)) << 4)) = o[i];
          fence_proxy_async_smem();
          // the store issued one slab ago must have drained its smem reads before anyone re-fills that buffer next turn
          if (leader) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
          asm volatile("bar.sync 1, 256;" ::: "memory");
          if (leader) {
            asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                             reinterpret_cast<uint64_t>(&tmC)),
                         "r"(smem_u32(sC + (slab_count & 1) * Cfg::SLAB_BYTES)), "r"(n0 + sl * 64), "r"(m0)
                         : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          }
        }
      } else {
#pragma unroll 1
        for (int c = 0; c < COLS; c += 32) {
          uint32_t acc[32];
          tmem_ld_32x32(taddr + c, acc);
          tmem_ld_wait();
          epilogue_chunk<EPI, OutT>(acc, row, n0 + half * COLS + c, p);
        }
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar_tempty[as]);
    }
    if (TMA_STORE && warp == 4 && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 2) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

template <int BN, int EPI, typename OutT>
static int launch_gemm(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC, const GemmParams& p,
                       cudaStream_t stream) {
  using Cfg = GemmCfg<BN>;
  auto kern = gemm_tc_kernel<BN, EPI, OutT>;
  static thread_local bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
    if (e != cudaSuccess) {
      set_error("cudaFuncSetAttribute(gemm_tc_kernel, smem=%d): %s", Cfg::SMEM_BYTES, cudaGetErrorString(e));
      return LRCE_ECUDA;
    }
    configured = true;
  }
  const int n_tiles = ((p.M + GEMM_BM - 1) / GEMM_BM) * (p.N / BN);
  int grid = sm_count();
  if (n_tiles < grid) grid = n_tiles;
  kern<<<grid, GEMM_THREADS, Cfg::SMEM_BYTES, stream>>>(tmA, tmB, tmC, p);
  return check_launch("gemm_tc_kernel");
}

template <int BN>
static int dispatch_epi(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC, const GemmParams& p, int epi,
                        int out_fp32, cudaStream_t stream) {
  if (out_fp32) {
    LRCE_REQUIRE(epi == EPI_BIAS, "fp32 output is only available with the bias epilogue (epi=%d)", epi);
    return launch_gemm<BN, EPI_BIAS, float>(tmA, tmB, tmC, p, stream);
  }
  switch (epi) {
    case EPI_BIAS: return launch_gemm<BN, EPI_BIAS, bf16>(tmA, tmB, tmC, p, stream);
    case EPI_BIAS_GELU: return launch_gemm<BN, EPI_BIAS_GELU, bf16>(tmA, tmB, tmC, p, stream);
    case EPI_BIAS_RESIDUAL: return launch_gemm<BN, EPI_BIAS_RESIDUAL, bf16>(tmA, tmB, tmC, p, stream);
    default: break;
  }
  set_error("unknown GEMM epilogue %d", epi);
  return LRCE_EINVAL;
}

}  // namespace lrce

using namespace lrce;

extern "C" int lrce_gemm_bf16(const void* A, int lda, const void* W, int ldw, int M, int N, int K, const float* bias,
                              const void* residual, int ldr, void* out, int ldo, int epilogue, int out_fp32,
                              const float* ln_gamma, const float* ln_beta, float ln_eps, void* stream) {
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
  GemmParams p;
  p.M = M; p.N = N; p.K = K;
  p.bias = bias;
  p.residual = reinterpret_cast<const bf16*>(residual);
  p.ldr = ldr;
  p.out = out;
  p.ldo = ldo;
  p.ln_g = ln_gamma; p.ln_b = ln_beta; p.ln_eps = ln_eps;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  CUtensorMap tmA, tmB, tmC;
  rc = make_tmap_2d_bf16(&tmA, A, K, M, lda, GEMM_BK, GEMM_BM);
  if (rc != LRCE_OK) return rc;
  if (!out_fp32 && epilogue != EPI_BIAS_LN) {  // bf16 tiles are written by TMA stores of 128 x 64 slabs
    rc = make_tmap_2d_bf16(&tmC, out, N, M, ldo, 64, GEMM_BM);
    if (rc != LRCE_OK) return rc;
  } else {
    tmC = tmA;  // unused by these epilogues
  }
  if (epilogue == EPI_BIAS_LN) {
    LRCE_REQUIRE(N == 128 && bias && ln_gamma && ln_beta && !out_fp32,
                 "lrce_gemm_bf16: the LayerNorm epilogue needs N == 128, a bias and gamma/beta (N=%d)", N);
    rc = make_tmap_2d_bf16(&tmB, W, K, N, ldw, GEMM_BK, 128);
    if (rc != LRCE_OK) return rc;
    return launch_gemm<128, EPI_BIAS_LN, bf16>(tmA, tmB, tmC, p, s);
  }
  const bool wide = (N % 256 == 0);
  rc = make_tmap_2d_bf16(&tmB, W, K, N, ldw, GEMM_BK, wide ? 256 : 128);
  if (rc != LRCE_OK) return rc;
  return wide ? dispatch_epi<256>(tmA, tmB, tmC, p, epilogue, out_fp32, s) : dispatch_epi<128>(tmA, tmB, tmC, p, epilogue, out_fp32, s);
}
