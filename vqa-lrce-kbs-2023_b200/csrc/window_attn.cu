// window_attn.cu — (shifted-)window multi-head attention core of Video Swin (video_swin_ori.py:166-186) with the
// cyclic shift / window_partition / window_reverse remap (video_swin_ori.py:262-276) fused into its loads and stores.
//
// Input  : qkv  bf16 [n_seg * D*H*W, 3C] in NATURAL token order (the qkv Linear is per token, so it runs before any
//          partition); column layout [q | k | v][head][32] (video_swin_ori.py:165).
// Output : out  bf16 [n_seg * D*H*W, C] in natural token order, heads merged (video_swin_ori.py:186), i.e. exactly
//          roll(window_reverse(attn @ v), +shift) — the proj GEMM + residual then runs with no remap at all.
//
// One CTA owns one head and walks a strided list of (segment, window) items:
//   * the head's dense relative-position bias (147 x 152 bf16, pre-multiplied by log2 e) stays resident in smem;
//   * per item the 147 q/k/v rows of the window are gathered from their rolled source tokens with 16-byte cp.async
//     (index = window_source_token(), the same function lrce_remap_index() exports for the bit-exact test);
//   * each warp takes 16-row stripes: S = q k^T on mma.sync m16n8k16 (bf16, fp32 accumulate), + bias, + shift mask
//     (-100 where region ids differ, video_swin_ori.py:357-358), exp2-softmax in registers, P v with P re-used
//     straight from the accumulator registers, 1/rowsum, and a scatter of 64-byte rows through the inverse remap.
// N = 147 is padded to 160 query rows / 152 key columns; padded keys are forced to -inf, padded rows never stored.
#include "host_common.h"
#include "lrce_common.cuh"
#include "remap.cuh"

namespace lrce {

constexpr int WA_N = 147;
constexpr int WA_ROWS = 160;     // 10 stripes of 16 query rows
constexpr int WA_KTILES = 19;    // 152 key columns
constexpr int WA_PITCH = 40;     // smem row pitch in bf16 (32 + 8 pad -> 80 B, conflict-free ldmatrix)
constexpr int WA_BIAS_PITCH = 152;
constexpr int WA_WARPS = 5;
constexpr int WA_THREADS = WA_WARPS * 32;
constexpr int WA_SMEM = 3 * WA_ROWS * WA_PITCH * 2 + WA_N * WA_BIAS_PITCH * 2 + WA_ROWS * 4 + WA_ROWS;

__device__ __forceinline__ void ldmatrix_x4(uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3, const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(smem_u32(p)));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3, const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(smem_u32(p)));
}
__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                               uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
      "{%0, %1, %2, %3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void cp_async_16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

__global__ void __launch_bounds__(WA_THREADS, 2)
window_attention_kernel(const bf16* __restrict__ qkv, bf16* __restrict__ out, const bf16* __restrict__ bias_dense,
                        StageGeom g, int n_seg, int C, float scale_log2e) {
  extern __shared__ __align__(16) uint8_t wa_smem[];
  bf16* sQ = reinterpret_cast<bf16*>(wa_smem);
  bf16* sK = sQ + WA_ROWS * WA_PITCH;
  bf16* sV = sK + WA_ROWS * WA_PITCH;
  bf16* sBias = sV + WA_ROWS * WA_PITCH;
  int* sTok = reinterpret_cast<int*>(sBias + WA_N * WA_BIAS_PITCH);
  uint8_t* sRid = reinterpret_cast<uint8_t*>(sTok + WA_ROWS);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int gq = lane >> 2, tq = lane & 3;
  const int head = blockIdx.y;
  const int nwin = windows_per_segment(g);
  const int T = g.D * g.H * g.W;
  const int n_items = n_seg * nwin;
  const bool shifted = (g.sd | g.sh | g.sw) != 0;
  const float MASK_L2 = -100.0f * 1.4426950408889634f;

  // one-time: zero q/k/v staging (pad rows stay zero forever), load this head's bias
  for (int i = tid; i < 3 * WA_ROWS * WA_PITCH / 8; i += WA_THREADS) reinterpret_cast<uint4*>(sQ)[i] = make_uint4(0, 0, 0, 0);
  {
    const uint4* src = reinterpret_cast<const uint4*>(bias_dense + static_cast<size_t>(head) * WA_N * WA_BIAS_PITCH);
    for (int i = tid; i < WA_N * WA_BIAS_PITCH / 8; i += WA_THREADS) reinterpret_cast<uint4*>(sBias)[i] = __ldg(src + i);
  }
  for (int i = tid; i < WA_ROWS; i += WA_THREADS) { sTok[i] = 0; sRid[i] = 0; }

  for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
    const int seg = item / nwin, win = item - seg * nwin;
    __syncthreads();  // previous item fully consumed (and the one-time init is visible)
    if (tid < WA_N) {
      sTok[tid] = window_source_token(g, win, tid);
      sRid[tid] = static_cast<uint8_t>(shifted ? shift_region_id(g, win, tid) : 0);
    }
    __syncthreads();
    {
      const bf16* base = qkv + static_cast<size_t>(seg) * T * 3 * C + head * 32;
      for (int c = tid; c < WA_N * 12; c += WA_THREADS) {
        const int r = c / 12, rem = c - r * 12, part = rem >> 2, ch = rem & 3;
        const bf16* src = base + static_cast<size_t>(sTok[r]) * 3 * C + part * C + ch * 8;
        cp_async_16(sQ + (part * WA_ROWS + r) * WA_PITCH + ch * 8, src);
      }
      cp_async_wait_all();
    }
    __syncthreads();

    // does this window straddle the wrap line? (otherwise every region id is equal and the mask is all zero)
    bool need_mask = false;
    if (shifted) {
      const int nw = g.W / g.ww, nh = g.H / g.wh;
      need_mask = ((win % nw) == nw - 1 && g.sw) || (((win / nw) % nh) == nh - 1 && g.sh);
    }

    for (int stripe = warp; stripe < WA_ROWS / 16; stripe += WA_WARPS) {
      const int r0 = stripe * 16;
      // ---- Q fragments (2 k-steps of 16 dims)
      uint32_t qa[2][4];
      {
        const int row = r0 + (lane & 7) + ((lane >> 3) & 1) * 8;
        const int col = (lane >> 4) * 8;
        ldmatrix_x4(qa[0][0], qa[0][1], qa[0][2], qa[0][3], sQ + row * WA_PITCH + col);
        ldmatrix_x4(qa[1][0], qa[1][1], qa[1][2], qa[1][3], sQ + row * WA_PITCH + 16 + col);
      }
      // ---- S = Q K^T
      float s[WA_KTILES][4];
#pragma unroll
      for (int j = 0; j < WA_KTILES; ++j) {
        s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f;
        uint32_t b0, b1, b2, b3;
        ldmatrix_x4(b0, b1, b2, b3, sK + (8 * j + (lane & 7)) * WA_PITCH + (lane >> 3) * 8);
        mma_bf16_16816(s[j], qa[0][0], qa[0][1], qa[0][2], qa[0][3], b0, b1);
        mma_bf16_16816(s[j], qa[1][0], qa[1][1], qa[1][2], qa[1][3], b2, b3);
      }
      // ---- scale, + bias, + mask (log2 domain), row max
      const int row_a = r0 + gq, row_b = r0 + gq + 8;
      const int brow_a = min(row_a, WA_N - 1), brow_b = min(row_b, WA_N - 1);
      const int rid_a = sRid[brow_a], rid_b = sRid[brow_b];
      float mx_a = -INFINITY, mx_b = -INFINITY;
#pragma unroll
      for (int j = 0; j < WA_KTILES; ++j) {
        const int col = 8 * j + 2 * tq;
        const float2 ba = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(sBias + brow_a * WA_BIAS_PITCH + col));
        const float2 bb = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(sBias + brow_b * WA_BIAS_PITCH + col));
        s[j][0] = fmaf(s[j][0], scale_log2e, ba.x);
        s[j][1] = fmaf(s[j][1], scale_log2e, ba.y);
        s[j][2] = fmaf(s[j][2], scale_log2e, bb.x);
        s[j][3] = fmaf(s[j][3], scale_log2e, bb.y);
        if (need_mask) {
          const int c0 = sRid[min(col, WA_N - 1)], c1 = sRid[min(col + 1, WA_N - 1)];
          if (c0 != rid_a) s[j][0] += MASK_L2;
          if (c1 != rid_a) s[j][1] += MASK_L2;
          if (c0 != rid_b) s[j][2] += MASK_L2;
          if (c1 != rid_b) s[j][3] += MASK_L2;
        }
        if (j == WA_KTILES - 1) {  // key columns 147..151 do not exist
          if (col >= WA_N) { s[j][0] = -INFINITY; s[j][2] = -INFINITY; }
          if (col + 1 >= WA_N) { s[j][1] = -INFINITY; s[j][3] = -INFINITY; }
        }
        mx_a = fmaxf(mx_a, fmaxf(s[j][0], s[j][1]));
        mx_b = fmaxf(mx_b, fmaxf(s[j][2], s[j][3]));
      }
      mx_a = fmaxf(mx_a, __shfl_xor_sync(0xffffffffu, mx_a, 1));
      mx_a = fmaxf(mx_a, __shfl_xor_sync(0xffffffffu, mx_a, 2));
      mx_b = fmaxf(mx_b, __shfl_xor_sync(0xffffffffu, mx_b, 1));
      mx_b = fmaxf(mx_b, __shfl_xor_sync(0xffffffffu, mx_b, 2));
      // ---- P = exp2(S - max), row sums
      float sum_a = 0.f, sum_b = 0.f;
#pragma unroll
      for (int j = 0; j < WA_KTILES; ++j) {
        s[j][0] = exp2f(s[j][0] - mx_a);
        s[j][1] = exp2f(s[j][1] - mx_a);
        s[j][2] = exp2f(s[j][2] - mx_b);
        s[j][3] = exp2f(s[j][3] - mx_b);
        sum_a += s[j][0] + s[j][1];
        sum_b += s[j][2] + s[j][3];
      }
      sum_a += __shfl_xor_sync(0xffffffffu, sum_a, 1);
      sum_a += __shfl_xor_sync(0xffffffffu, sum_a, 2);
      sum_b += __shfl_xor_sync(0xffffffffu, sum_b, 1);
      sum_b += __shfl_xor_sync(0xffffffffu, sum_b, 2);
      // ---- O = P V  (P taken from the accumulator registers as the A operand)
      float o[4][4];
#pragma unroll
      for (int nn = 0; nn < 4; ++nn) o[nn][0] = o[nn][1] = o[nn][2] = o[nn][3] = 0.f;
#pragma unroll
      for (int kk = 0; kk < (WA_KTILES + 1) / 2; ++kk) {
        const uint32_t a0 = pack_bf16x2(s[2 * kk][0], s[2 * kk][1]);
        const uint32_t a1 = pack_bf16x2(s[2 * kk][2], s[2 * kk][3]);
        uint32_t a2 = 0u, a3 = 0u;
        if (2 * kk + 1 < WA_KTILES) {
          a2 = pack_bf16x2(s[2 * kk + 1][0], s[2 * kk + 1][1]);
          a3 = pack_bf16x2(s[2 * kk + 1][2], s[2 * kk + 1][3]);
        }
        const int vrow = kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
        const int vcol = (lane >> 4) * 8;
        uint32_t b0, b1, b2, b3;
        ldmatrix_x4_trans(b0, b1, b2, b3, sV + vrow * WA_PITCH + vcol);
        mma_bf16_16816(o[0], a0, a1, a2, a3, b0, b1);
        mma_bf16_16816(o[1], a0, a1, a2, a3, b2, b3);
        ldmatrix_x4_trans(b0, b1, b2, b3, sV + vrow * WA_PITCH + 16 + vcol);
        mma_bf16_16816(o[2], a0, a1, a2, a3, b0, b1);
        mma_bf16_16816(o[3], a0, a1, a2, a3, b2, b3);
      }
      // ---- normalise, stage the stripe's 16 x 32 output in this warp's (now dead) Q rows, scatter 64-byte rows
      const float inv_a = 1.0f / sum_a, inv_b = 1.0f / sum_b;
      __syncwarp();
#pragma unroll
      for (int nn = 0; nn < 4; ++nn) {
        *reinterpret_cast<uint32_t*>(sQ + row_a * WA_PITCH + nn * 8 + 2 * tq) = pack_bf16x2(o[nn][0] * inv_a, o[nn][1] * inv_a);
        *reinterpret_cast<uint32_t*>(sQ + row_b * WA_PITCH + nn * 8 + 2 * tq) = pack_bf16x2(o[nn][2] * inv_b, o[nn][3] * inv_b);
      }
      __syncwarp();
#pragma unroll
      for (int h2 = 0; h2 < 2; ++h2) {
        const int row = r0 + h2 * 8 + (lane >> 2);
        if (row < WA_N) {
          const uint4 v = *reinterpret_cast<const uint4*>(sQ + row * WA_PITCH + (lane & 3) * 8);
          bf16* dst = out + (static_cast<size_t>(seg) * T + sTok[row]) * C + head * 32 + (lane & 3) * 8;
          *reinterpret_cast<uint4*>(dst) = v;
        }
      }
    }
  }
}

// bias_dense[h][i][j] = table[rel_index(i, j)][h] * log2(e), j padded to 152 with zeros
__global__ void build_dense_bias_kernel(const float* __restrict__ table, bf16* __restrict__ dense, StageGeom g,
                                        int n_heads) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int total = n_heads * WA_N * WA_BIAS_PITCH;
  if (idx >= total) return;
  const int j = idx % WA_BIAS_PITCH, i = (idx / WA_BIAS_PITCH) % WA_N, h = idx / (WA_BIAS_PITCH * WA_N);
  float v = 0.f;
  if (j < WA_N) {
    const int rel = rel_pos_offset(g, i) - rel_pos_offset(g, j) + REL_POS_CENTER;
    v = table[static_cast<size_t>(rel) * n_heads + h] * 1.4426950408889634f;
  }
  dense[idx] = __float2bfloat16(v);
}

}  // namespace lrce

using namespace lrce;

static int geom_3x7x7(StageGeom* g, int D, int H, int W, int sh, int sw) {
  LRCE_REQUIRE(D == 3 && H % 7 == 0 && W % 7 == 0 && H > 0 && W > 0,
               "window attention is specialised for the clamped (3,7,7) window of LRCE's 5-frame segments; got grid "
               "(%d,%d,%d)", D, H, W);
  LRCE_REQUIRE(sh >= 0 && sh < 7 && sw >= 0 && sw < 7, "shift must be in [0,7)");
  g->D = D; g->H = H; g->W = W; g->wd = 3; g->wh = 7; g->ww = 7; g->sd = 0; g->sh = sh; g->sw = sw;
  return LRCE_OK;
}

extern "C" int lrce_window_attention_bf16(const void* qkv, void* out, const void* bias_dense, int n_seg, int D, int H,
                                          int W, int C, int n_heads, int shift_h, int shift_w, void* stream) {
  int rc = require_sm100();
  if (rc != LRCE_OK) return rc;
  StageGeom g;
  rc = geom_3x7x7(&g, D, H, W, shift_h, shift_w);
  if (rc != LRCE_OK) return rc;
  LRCE_REQUIRE(qkv && out && bias_dense && n_seg > 0, "lrce_window_attention_bf16: null operand");
  LRCE_REQUIRE(n_heads > 0 && C == n_heads * 32, "lrce_window_attention_bf16: head_dim must be 32 (C=%d heads=%d)", C, n_heads);
  static thread_local bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(window_attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, WA_SMEM);
    if (e != cudaSuccess) {
      set_error("cudaFuncSetAttribute(window_attention_kernel): %s", cudaGetErrorString(e));
      return LRCE_ECUDA;
    }
    configured = true;
  }
  const int n_items = n_seg * windows_per_segment(g);
  int groups = (2 * sm_count() + n_heads - 1) / n_heads;
  if (groups > n_items) groups = n_items;
  if (groups < 1) groups = 1;
  dim3 grid(groups, n_heads);
  const float scale_log2e = 0.17677669529663687f * 1.4426950408889634f;  // 32^-0.5 * log2(e)
  window_attention_kernel<<<grid, WA_THREADS, WA_SMEM, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const bf16*>(qkv), reinterpret_cast<bf16*>(out), reinterpret_cast<const bf16*>(bias_dense), g,
      n_seg, C, scale_log2e);
  return check_launch("window_attention_kernel");
}

extern "C" int lrce_window_bias_pack(const float* table, void* bias_dense, int n_heads, void* stream) {
  int rc = require_sm100();
  if (rc != LRCE_OK) return rc;
  LRCE_REQUIRE(table && bias_dense && n_heads > 0, "lrce_window_bias_pack: bad arguments");
  StageGeom g;
  g.D = 3; g.H = 7; g.W = 7; g.wd = 3; g.wh = 7; g.ww = 7; g.sd = g.sh = g.sw = 0;
  const int total = n_heads * WA_N * WA_BIAS_PITCH;
  build_dense_bias_kernel<<<(total + 255) / 256, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      table, reinterpret_cast<bf16*>(bias_dense), g, n_heads);
  return check_launch("build_dense_bias_kernel");
}
