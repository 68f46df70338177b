// encoder.cu — the recurrent cross-modal encoder's small kernels (lrce/models/fusionv3.py, embedding.py).
//
// The heavy part of the encoder — the K/V in-projection of every memory token for all 12 layers — is one tcgen05 GEMM
// (gemm_tc.cu). What remains is the single summarisation token walking 12 layers x S segments (fusionv3.py:41-51):
// a chain of (rows <= 32) x 768 mat-vec-like products that is weight-bandwidth / latency bound. Kernels here:
//   video_posembed_ln / text_posembed_ln : embedding.py:47-63 / :17-23 fused (CLS row, 3 adds, LayerNorm eps 1e-12)
//   skinny_linear   : Y = act(LN?(Xa + Xb) W^T + b) for <= 32 rows per CTA-row; weights streamed once with 16-byte
//                     loads straight into mma.sync B fragments (K order permuted consistently on both operands),
//                     K split across the 8 warps of a CTA, the residual-add + LayerNorm of the PREVIOUS sub-layer
//                     fused as prologue (post-norm decoder: x = LN(x + f(x)))
//   cross_attention : one warp per (row, head): 1 query x (150 video + Lt text) keys of the precomputed K/V
//   recurrent_update: tok = LN_f(tok + LN3(h + y))  (fusionv3.py:47-48 after the 12th layer)
#include "host_common.h"
#include "encoder_common.cuh"

namespace lrce {

// out[b,s,t,p] = LN(emb_pos[p] + emb_len[t] + emb_clip[s] + (p == 0 ? emb_cls : proj[b,s,t,p-1]))
__global__ void __launch_bounds__(256) video_posembed_ln_kernel(const bf16* __restrict__ proj, const float* __restrict__ emb_cls,
                                                                const float* __restrict__ emb_pos, const float* __restrict__ emb_len,
                                                                const float* __restrict__ emb_clip, const float* __restrict__ gamma,
                                                                const float* __restrict__ beta, float eps, bf16* __restrict__ out,
                                                                int B, int S, int T, int P) {
  const int lane = threadIdx.x & 31;
  const long long row = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const long long rows = static_cast<long long>(B) * S * T * (P + 1);
  if (row >= rows) return;
  const int p = static_cast<int>(row % (P + 1));
  const long long frame = row / (P + 1);  // (b*S + s)*T + t
  const int t = static_cast<int>(frame % T);
  const int s = static_cast<int>((frame / T) % S);
  float v[24];
#pragma unroll
  for (int i = 0; i < 24; ++i) v[i] = 0.f;
  if (p == 0) row768_add_f32(v, emb_cls, lane);
  else row768_add_bf16(v, proj + (frame * P + (p - 1)) * ENC_D, lane);
  row768_add_f32(v, emb_pos + static_cast<size_t>(p) * ENC_D, lane);
  row768_add_f32(v, emb_len + static_cast<size_t>(t) * ENC_D, lane);
  row768_add_f32(v, emb_clip + static_cast<size_t>(s) * ENC_D, lane);
  row768_ln_store(v, gamma, beta, eps, lane, out + row * ENC_D, nullptr);
}

// out[b, l] = LN(emb_pos[l] + (l == 0 ? emb_cls : text[b, l-1]))
template <typename TextT>
__global__ void __launch_bounds__(256) text_posembed_ln_kernel(const TextT* __restrict__ text, const float* __restrict__ emb_cls,
                                                               const float* __restrict__ emb_pos, const float* __restrict__ gamma,
                                                               const float* __restrict__ beta, float eps, bf16* __restrict__ out,
                                                               int Bt, int L) {
  const int lane = threadIdx.x & 31;
  const long long row = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  if (row >= static_cast<long long>(Bt) * (L + 1)) return;
  const int l = static_cast<int>(row % (L + 1));
  const long long b = row / (L + 1);
  float v[24];
#pragma unroll
  for (int i = 0; i < 24; ++i) v[i] = 0.f;
  if (l == 0) row768_add_f32(v, emb_cls, lane);
  else if (sizeof(TextT) == 2) row768_add_bf16(v, reinterpret_cast<const bf16*>(text) + (b * L + (l - 1)) * ENC_D, lane);
  else row768_add_f32(v, reinterpret_cast<const float*>(text) + (b * L + (l - 1)) * ENC_D, lane);
  row768_add_f32(v, emb_pos + static_cast<size_t>(l) * ENC_D, lane);
  row768_ln_store(v, gamma, beta, eps, lane, out + row * ENC_D, nullptr);
}

// tok_out = LN_f(tok + LN3(h + y))   (one warp per row)
__global__ void __launch_bounds__(128) recurrent_update_kernel(const float* __restrict__ tok, const float* __restrict__ h,
                                                               const float* __restrict__ y, const float* __restrict__ g3,
                                                               const float* __restrict__ b3, const float* __restrict__ gf,
                                                               const float* __restrict__ bf, float eps, float* __restrict__ tok_out,
                                                               int rows) {
  const int lane = threadIdx.x & 31;
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (row >= rows) return;
  float v[24];
#pragma unroll
  for (int i = 0; i < 24; ++i) v[i] = 0.f;
  row768_add_f32(v, h + static_cast<size_t>(row) * ENC_D, lane);
  row768_add_f32(v, y + static_cast<size_t>(row) * ENC_D, lane);
  row768_ln_store(v, g3, b3, eps, lane, nullptr, nullptr);  // v <- LN3(h + y)
  row768_add_f32(v, tok + static_cast<size_t>(row) * ENC_D, lane);
  row768_ln_store(v, gf, bf, eps, lane, nullptr, tok_out + static_cast<size_t>(row) * ENC_D);
}

// ---------------------------------------------------------------------------------------------------------------
// skinny linear
// ---------------------------------------------------------------------------------------------------------------
constexpr int SK_ROWS = 32;
constexpr int SK_WARPS = 8;
constexpr int SK_THREADS = SK_WARPS * 32;
enum { ACT_NONE = 0, ACT_GELU = 1, ACT_RELU = 2 };

// grid = (ceil(N/8), ceil(rows/32)); dynamic smem = 32*(K+32)*2 + 8*32*8*4
// Xa is fp32 [rows, K] (xa_bf16 == 0) or bf16 [rows, K] (xa_bf16 != 0, plain copy prologue: no Xb / LayerNorm).
__global__ void __launch_bounds__(SK_THREADS) skinny_linear_kernel(const void* __restrict__ Xa_, const float* __restrict__ Xb,
                                                                   const float* __restrict__ ln_g, const float* __restrict__ ln_b,
                                                                   float eps, float* __restrict__ Xout,
                                                                   const bf16* __restrict__ Wt, const float* __restrict__ bias,
                                                                   void* __restrict__ Y_, int rows, int K, int N, int ldy, int act,
                                                                   int xa_bf16, int y_bf16) {
  extern __shared__ __align__(16) uint8_t sk_smem[];
  const int pitch = K + 32;  // bf16 elements; (pitch/2) % 32 == 16 words -> conflict-free 16-byte fragment loads
  bf16* sX = reinterpret_cast<bf16*>(sk_smem);
  float* sRed = reinterpret_cast<float*>(sk_smem + static_cast<size_t>(SK_ROWS) * pitch * 2);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int r_base = blockIdx.y * SK_ROWS;
  const int n0 = blockIdx.x * 8;

  // ---- prologue: x = Xa (+ Xb) (-> LayerNorm) ; bf16 copy to smem ; optional fp32 write-back by the first CTA column
  if (xa_bf16) {
    // producer already emitted bf16: one wave of 16-byte async copies, all in flight together
    const bf16* Xa = reinterpret_cast<const bf16*>(Xa_);
    const int chunks_per_row = K / 8;
    for (int c = tid; c < SK_ROWS * chunks_per_row; c += SK_THREADS) {
      const int r = c / chunks_per_row, k = (c - r * chunks_per_row) * 8;
      bf16* dst = sX + static_cast<size_t>(r) * pitch + k;
      if (r_base + r < rows) {
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)),
                     "l"(Xa + static_cast<size_t>(r_base + r) * K + k) : "memory");
      } else {
        *reinterpret_cast<uint4*>(dst) = make_uint4(0, 0, 0, 0);
      }
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
  } else if (ln_g != nullptr) {
    // K == 768: the 4 rows of this warp are processed together so their loads and reductions overlap
    const float* Xa = reinterpret_cast<const float*>(Xa_);
    float v[SK_ROWS / SK_WARPS][24];
#pragma unroll
    for (int i = 0; i < SK_ROWS / SK_WARPS; ++i) {
      const int row = r_base + warp + i * SK_WARPS;
#pragma unroll
      for (int j = 0; j < 24; ++j) v[i][j] = 0.f;
      if (row < rows) {
        row768_add_f32(v[i], Xa + static_cast<size_t>(row) * K, lane);
        if (Xb) row768_add_f32(v[i], Xb + static_cast<size_t>(row) * K, lane);
      }
    }
#pragma unroll
    for (int i = 0; i < SK_ROWS / SK_WARPS; ++i) {
      const int r = warp + i * SK_WARPS, row = r_base + r;
      row768_ln_store(v[i], ln_g, ln_b, eps, lane, sX + static_cast<size_t>(r) * pitch,
                      (Xout && blockIdx.x == 0 && row < rows) ? Xout + static_cast<size_t>(row) * K : nullptr);
    }
  } else {
    const float* Xa = reinterpret_cast<const float*>(Xa_);
    for (int k = lane * 8; k < K; k += 256) {
      float4 x0[SK_ROWS / SK_WARPS], x1[SK_ROWS / SK_WARPS];
#pragma unroll
      for (int i = 0; i < SK_ROWS / SK_WARPS; ++i) {
        const int row = r_base + warp + i * SK_WARPS;
        x0[i] = x1[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (row < rows) {
          const float* a = Xa + static_cast<size_t>(row) * K + k;
          x0[i] = __ldg(reinterpret_cast<const float4*>(a));
          x1[i] = __ldg(reinterpret_cast<const float4*>(a + 4));
          if (Xb) {
            const float* b = Xb + static_cast<size_t>(row) * K + k;
            const float4 y0 = __ldg(reinterpret_cast<const float4*>(b)), y1 = __ldg(reinterpret_cast<const float4*>(b + 4));
            x0[i].x += y0.x; x0[i].y += y0.y; x0[i].z += y0.z; x0[i].w += y0.w;
            x1[i].x += y1.x; x1[i].y += y1.y; x1[i].z += y1.z; x1[i].w += y1.w;
          }
        }
      }
#pragma unroll
      for (int i = 0; i < SK_ROWS / SK_WARPS; ++i) {
        uint4 u;
        u.x = pack_bf16x2(x0[i].x, x0[i].y); u.y = pack_bf16x2(x0[i].z, x0[i].w);
        u.z = pack_bf16x2(x1[i].x, x1[i].y); u.w = pack_bf16x2(x1[i].z, x1[i].w);
        *reinterpret_cast<uint4*>(sX + static_cast<size_t>(warp + i * SK_WARPS) * pitch + k) = u;
      }
    }
  }
  __syncthreads();

  // ---- main: this warp's K slice; thread (g, t) streams 16 B of weight row n0+g per 32-wide k chunk
  const int g = lane >> 2, t = lane & 3;
  const int k_per_warp = K / SK_WARPS;  // multiple of 32 (K in {768, 3072})
  const int k_begin = warp * k_per_warp;
  const bf16* wrow = Wt + static_cast<size_t>(n0 + g) * K + k_begin + 8 * t;
  const bf16* xa0 = sX + static_cast<size_t>(g) * pitch + k_begin + 8 * t;
  float acc[2][4];
#pragma unroll
  for (int m = 0; m < 2; ++m) acc[m][0] = acc[m][1] = acc[m][2] = acc[m][3] = 0.f;
  const int n_chunks = k_per_warp / 32;
#pragma unroll 4
  for (int c = 0; c < n_chunks; ++c) {
    const uint4 w = __ldg(reinterpret_cast<const uint4*>(wrow + c * 32));
#pragma unroll
    for (int m = 0; m < 2; ++m) {
      const uint4 xlo = *reinterpret_cast<const uint4*>(xa0 + static_cast<size_t>(m * 16) * pitch + c * 32);
      const uint4 xhi = *reinterpret_cast<const uint4*>(xa0 + static_cast<size_t>(m * 16 + 8) * pitch + c * 32);
      mma16816(acc[m], xlo.x, xhi.x, xlo.y, xhi.y, w.x, w.y);
      mma16816(acc[m], xlo.z, xhi.z, xlo.w, xhi.w, w.z, w.w);
    }
  }
  // ---- cross-warp K reduction, bias, activation, store
#pragma unroll
  for (int m = 0; m < 2; ++m) {
    float* r = sRed + (warp * SK_ROWS + m * 16 + g) * 8 + 2 * t;
    r[0] = acc[m][0]; r[1] = acc[m][1];
    r[8 * 8] = acc[m][2]; r[8 * 8 + 1] = acc[m][3];
  }
  __syncthreads();
  {
    const int r = tid >> 3, col = tid & 7;
    float v = 0.f;
#pragma unroll
    for (int w = 0; w < SK_WARPS; ++w) v += sRed[(w * SK_ROWS + r) * 8 + col];
    const int row = r_base + r, n = n0 + col;
    if (row < rows && n < N) {
      if (bias) v += __ldg(bias + n);
      if (act == ACT_GELU) v = gelu_erf(v);
      else if (act == ACT_RELU) v = fmaxf(v, 0.f);
      if (y_bf16) reinterpret_cast<bf16*>(Y_)[static_cast<size_t>(row) * ldy + n] = __float2bfloat16(v);
      else reinterpret_cast<float*>(Y_)[static_cast<size_t>(row) * ldy + n] = v;
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// cross attention of the single query token: grid = (rows, 12 heads), 4 warps; each warp owns a quarter of the keys.
// 8 lanes cover one 128-byte K (or V) head-row with 16-byte loads, 4 keys per warp-iteration, all loads of a warp are
// issued before they are consumed (the kernel is pure latency: 47 KB of K/V per CTA).
// ---------------------------------------------------------------------------------------------------------------
constexpr int CA_MAX_KEYS = 256;
constexpr int CA_WARPS = 4;
constexpr int CA_ITERS = CA_MAX_KEYS / (4 * CA_WARPS);  // 16 key-quads per warp at most

__global__ void __launch_bounds__(128) cross_attention_kernel(const float* __restrict__ q, const bf16* __restrict__ kv_video,
                                                              const bf16* __restrict__ kv_text, bf16* __restrict__ ctx,
                                                              int seg, int S, int Tv, int Lt, int n_cand, int layer,
                                                              int ld_kv) {
  __shared__ float sP[CA_MAX_KEYS];
  __shared__ float sRed[2][CA_WARPS];
  __shared__ float sO[CA_WARPS][64];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x, head = blockIdx.y;
  const int n_keys = Tv + Lt;
  const size_t col_k = static_cast<size_t>(layer) * 2 * ENC_D + head * 64;
  const size_t col_v = col_k + ENC_D;
  const bf16* vid = kv_video + (static_cast<size_t>(b / n_cand) * S + seg) * Tv * ld_kv;
  const bf16* txt = kv_text + static_cast<size_t>(b) * Lt * ld_kv;
  const int l8 = lane & 7, kslot = lane >> 3;
  const int per_warp = (n_keys + CA_WARPS - 1) / CA_WARPS;
  const int k_lo = warp * per_warp, k_hi = min(n_keys, k_lo + per_warp);
  float qv[8];
  {
    const float* qp = q + static_cast<size_t>(b) * ENC_D + head * 64 + l8 * 8;
    const float4 a = __ldg(reinterpret_cast<const float4*>(qp)), c = __ldg(reinterpret_cast<const float4*>(qp + 4));
    qv[0] = a.x; qv[1] = a.y; qv[2] = a.z; qv[3] = a.w; qv[4] = c.x; qv[5] = c.y; qv[6] = c.z; qv[7] = c.w;
  }
  auto row_ptr = [&](int j) { return j < Tv ? vid + static_cast<size_t>(j) * ld_kv : txt + static_cast<size_t>(j - Tv) * ld_kv; };
  // ---- scores
  uint4 kreg[CA_ITERS];
#pragma unroll
  for (int it = 0; it < CA_ITERS; ++it) {
    const int j = k_lo + it * 4 + kslot;
    kreg[it] = make_uint4(0, 0, 0, 0);
    if (j < k_hi) kreg[it] = __ldg(reinterpret_cast<const uint4*>(row_ptr(j) + col_k + l8 * 8));
  }
  float mx = -INFINITY;
#pragma unroll
  for (int it = 0; it < CA_ITERS; ++it) {
    const int j = k_lo + it * 4 + kslot;
    float2 f;
    float d = 0.f;
    f = unpack_bf16x2(kreg[it].x); d += qv[0] * f.x + qv[1] * f.y;
    f = unpack_bf16x2(kreg[it].y); d += qv[2] * f.x + qv[3] * f.y;
    f = unpack_bf16x2(kreg[it].z); d += qv[4] * f.x + qv[5] * f.y;
    f = unpack_bf16x2(kreg[it].w); d += qv[6] * f.x + qv[7] * f.y;
    d += __shfl_xor_sync(0xffffffffu, d, 1);
    d += __shfl_xor_sync(0xffffffffu, d, 2);
    d += __shfl_xor_sync(0xffffffffu, d, 4);
    if (j < k_hi) {
      if (l8 == 0) sP[j] = d;
      mx = fmaxf(mx, d);
    }
  }
  // ---- V rows are independent of the softmax statistics: fetch them now, reduce max / sum meanwhile
  uint4 vreg[CA_ITERS];
#pragma unroll
  for (int it = 0; it < CA_ITERS; ++it) {
    const int j = k_lo + it * 4 + kslot;
    vreg[it] = make_uint4(0, 0, 0, 0);
    if (j < k_hi) vreg[it] = __ldg(reinterpret_cast<const uint4*>(row_ptr(j) + col_v + l8 * 8));
  }
  mx = warp_max(mx);
  if (lane == 0) sRed[0][warp] = mx;
  __syncthreads();
  mx = fmaxf(fmaxf(sRed[0][0], sRed[0][1]), fmaxf(sRed[0][2], sRed[0][3]));
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  float sum = 0.f;
#pragma unroll
  for (int it = 0; it < CA_ITERS; ++it) {
    const int j = k_lo + it * 4 + kslot;
    if (j < k_hi) {
      const float p = __expf(sP[j] - mx);
      if (l8 == 0) sum += p;
      float2 f;
      f = unpack_bf16x2(vreg[it].x); acc[0] = fmaf(p, f.x, acc[0]); acc[1] = fmaf(p, f.y, acc[1]);
      f = unpack_bf16x2(vreg[it].y); acc[2] = fmaf(p, f.x, acc[2]); acc[3] = fmaf(p, f.y, acc[3]);
      f = unpack_bf16x2(vreg[it].z); acc[4] = fmaf(p, f.x, acc[4]); acc[5] = fmaf(p, f.y, acc[5]);
      f = unpack_bf16x2(vreg[it].w); acc[6] = fmaf(p, f.x, acc[6]); acc[7] = fmaf(p, f.y, acc[7]);
    }
  }
  sum = warp_sum(sum);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], 8);
    acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], 16);
  }
  if (lane == 0) sRed[1][warp] = sum;
  if (lane < 8) {
#pragma unroll
    for (int i = 0; i < 8; ++i) sO[warp][lane * 8 + i] = acc[i];
  }
  __syncthreads();
  if (threadIdx.x < 64) {
    const float tot = sRed[1][0] + sRed[1][1] + sRed[1][2] + sRed[1][3];
    const float o = sO[0][threadIdx.x] + sO[1][threadIdx.x] + sO[2][threadIdx.x] + sO[3][threadIdx.x];
    ctx[static_cast<size_t>(b) * ENC_D + head * 64 + threadIdx.x] = __float2bfloat16(o / tot);
  }
}

}  // namespace lrce

using namespace lrce;

extern "C" int lrce_video_posembed_ln(const void* proj, const float* emb_cls, const float* emb_pos, const float* emb_len,
                                      const float* emb_clip, const float* gamma, const float* beta, float eps, void* out,
                                      int B, int S, int T, int P, void* stream) {
  int rc = require_sm100();
  if (rc != LRCE_OK) return rc;
  LRCE_REQUIRE(proj && emb_cls && emb_pos && emb_len && emb_clip && gamma && beta && out && B > 0 && S > 0 && T > 0 && P > 0,
               "lrce_video_posembed_ln: bad arguments");
  const long long rows = static_cast<long long>(B) * S * T * (P + 1);
  video_posembed_ln_kernel<<<static_cast<unsigned>((rows + 7) / 8), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const bf16*>(proj), emb_cls, emb_pos, emb_len, emb_clip, gamma, beta, eps,
      reinterpret_cast<bf16*>(out), B, S, T, P);
  return check_launch("video_posembed_ln_kernel");
}

extern "C" int lrce_text_posembed_ln(const void* text, int text_fp32, const float* emb_cls, const float* emb_pos, const float* gamma,
                                     const float* beta, float eps, void* out, int Bt, int L, void* stream) {
  int rc = require_sm100();
  if (rc != LRCE_OK) return rc;
  LRCE_REQUIRE(text && emb_cls && emb_pos && gamma && beta && out && Bt > 0 && L > 0, "lrce_text_posembed_ln: bad arguments");
  const long long rows = static_cast<long long>(Bt) * (L + 1);
  const unsigned blocks = static_cast<unsigned>((rows + 7) / 8);
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (text_fp32)
    text_posembed_ln_kernel<float><<<blocks, 256, 0, s>>>(reinterpret_cast<const float*>(text), emb_cls, emb_pos, gamma, beta,
                                                          eps, reinterpret_cast<bf16*>(out), Bt, L);
  else
    text_posembed_ln_kernel<bf16><<<blocks, 256, 0, s>>>(reinterpret_cast<const bf16*>(text), emb_cls, emb_pos, gamma, beta,
                                                         eps, reinterpret_cast<bf16*>(out), Bt, L);
  return check_launch("text_posembed_ln_kernel");
}

extern "C" int lrce_recurrent_update(const float* tok, const float* h, const float* y, const float* g3, const float* b3,
                                     const float* gf, const float* bf, float eps, float* tok_out, int rows, void* stream) {
  int rc = require_sm100();
  if (rc != LRCE_OK) return rc;
  LRCE_REQUIRE(tok && h && y && g3 && b3 && gf && bf && tok_out && rows > 0, "lrce_recurrent_update: bad arguments");
  recurrent_update_kernel<<<(rows + 3) / 4, 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(tok, h, y, g3, b3, gf, bf, eps,
                                                                                                tok_out, rows);
  return check_launch("recurrent_update_kernel");
}

extern "C" int lrce_skinny_linear(const void* Xa, int xa_bf16, const float* Xb, const float* ln_gamma, const float* ln_beta,
                                  float eps, float* Xout, const void* W, const float* bias, void* Y, int y_bf16, int rows,
                                  int K, int N, int ldy, int act, void* stream) {
  int rc = require_sm100();
  if (rc != LRCE_OK) return rc;
  LRCE_REQUIRE(Xa && W && Y && rows > 0 && N > 0, "lrce_skinny_linear: bad arguments");
  LRCE_REQUIRE(K == 768 || K == 3072, "lrce_skinny_linear: K must be 768 or 3072 (got %d)", K);
  LRCE_REQUIRE(ln_gamma == nullptr || (K == 768 && ln_beta != nullptr), "lrce_skinny_linear: LayerNorm prologue needs K == 768");
  LRCE_REQUIRE(act >= 0 && act <= 2, "lrce_skinny_linear: unknown activation %d", act);
  LRCE_REQUIRE(!xa_bf16 || (Xb == nullptr && ln_gamma == nullptr), "lrce_skinny_linear: a bf16 input takes no Xb / LayerNorm prologue");
  LRCE_REQUIRE((reinterpret_cast<uintptr_t>(Xa) & 15) == 0, "lrce_skinny_linear: X must be 16-byte aligned");
  const int smem = SK_ROWS * (K + 32) * 2 + SK_WARPS * SK_ROWS * 8 * 4;
  static thread_local bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(skinny_linear_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         SK_ROWS * (3072 + 32) * 2 + SK_WARPS * SK_ROWS * 8 * 4);
    if (e != cudaSuccess) {
      set_error("cudaFuncSetAttribute(skinny_linear_kernel): %s", cudaGetErrorString(e));
      return LRCE_ECUDA;
    }
    configured = true;
  }
  dim3 grid((N + 7) / 8, (rows + SK_ROWS - 1) / SK_ROWS);
  skinny_linear_kernel<<<grid, SK_THREADS, smem, reinterpret_cast<cudaStream_t>(stream)>>>(
      Xa, Xb, ln_gamma, ln_beta, eps, Xout, reinterpret_cast<const bf16*>(W), bias, Y, rows, K, N, ldy, act, xa_bf16, y_bf16);
  return check_launch("skinny_linear_kernel");
}

extern "C" int lrce_cross_attention(const float* q, const void* kv_video, const void* kv_text, void* ctx, int rows, int seg,
                                    int S, int Tv, int Lt, int n_cand, int layer, int ld_kv, void* stream) {
  int rc = require_sm100();
  if (rc != LRCE_OK) return rc;
  LRCE_REQUIRE(q && kv_video && kv_text && ctx && rows > 0 && n_cand > 0, "lrce_cross_attention: bad arguments");
  LRCE_REQUIRE(Tv + Lt <= CA_MAX_KEYS, "lrce_cross_attention: %d memory tokens exceed the %d-key limit", Tv + Lt, CA_MAX_KEYS);
  LRCE_REQUIRE(ld_kv % 8 == 0, "lrce_cross_attention: K/V row pitch must be a multiple of 8");
  dim3 grid(rows, 12);
  cross_attention_kernel<<<grid, 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      q, reinterpret_cast<const bf16*>(kv_video), reinterpret_cast<const bf16*>(kv_text), reinterpret_cast<bf16*>(ctx), seg, S,
      Tv, Lt, n_cand, layer, ld_kv);
  return check_launch("cross_attention_kernel");
}
