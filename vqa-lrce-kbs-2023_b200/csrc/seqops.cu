// seqops.cu — the small row / sequence kernels around the tcgen05 GEMM for (a) BERT-base on liblrce_b200
// (lrce/feature_extractor/text.py:5-17: embeddings + LayerNorm, 12 post-norm layers with a <= 64-token masked attention) and
// (b) the TRAINING step of the recurrent cross-modal encoder (BASELINE.json configs[4]; fusionv3.py:41-51 forward with the
// activations a backward pass needs, and that backward pass: LayerNorm / GELU / single-query cross-attention / pos-embed
// gradients). Every Linear of both paths is lrce_gemm_bf16; these kernels produce its bf16 A operands — row-major for the
// activation-gradient GEMMs (dX = dY W) and TRANSPOSED, stacked over the recurrent steps, for the weight-gradient GEMMs
// (dW = dY^T X as an ordinary K-major product over K = tokens) — so no layout copy sits between two GEMMs.
//
// Row kernels: one warp owns one 768-wide row; lane holds 24 values as 3 chunks of 8 at columns (c * 32 + lane) * 8.
// Dropout (train mode: nn.Dropout / MultiheadAttention dropout of fusionv3.py:8-17, :49, :190-191) is a counter-based
// hash of (seed, site, element index): the backward pass regenerates the mask instead of storing it.
#include "encoder_common.cuh"
#include "host_common.h"

namespace lrce {

// ------------------------------------------------------------------------------------------------------------------
// dropout
// ------------------------------------------------------------------------------------------------------------------
struct Drop {
  float p;      // 0 = off
  float scale;  // 1 / (1 - p)
  unsigned long long seed;
  const unsigned long long* seed_dev;  // optional device-resident seed XOR-ed in at run time: a CUDA graph that captured
                                       // this launch draws a fresh mask on every replay
  uint32_t site;
  uint32_t thresh;  // p * 2^32
};
__host__ inline Drop make_drop(float p, unsigned long long seed, const unsigned long long* seed_dev, int site) {
  Drop d;
  d.p = (p > 0.f && p < 1.f) ? p : 0.f;
  d.scale = d.p > 0.f ? 1.0f / (1.0f - d.p) : 1.0f;
  d.seed = seed;
  d.seed_dev = seed_dev;
  d.site = static_cast<uint32_t>(site);
  d.thresh = static_cast<uint32_t>(static_cast<double>(d.p) * 4294967296.0);
  return d;
}
__device__ __forceinline__ uint32_t rng_u32(unsigned long long seed, uint32_t site, uint32_t idx) {
  unsigned long long z = seed ^ ((static_cast<unsigned long long>(site) << 32) | idx);  // splitmix64 finaliser
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z ^= z >> 31;
  return static_cast<uint32_t>(z >> 32);
}
// multiplier of element idx: 0 (dropped) or 1 / (1 - p)
__device__ __forceinline__ float drop_mul(const Drop& d, uint32_t idx) {
  if (d.p <= 0.f) return 1.f;
  const unsigned long long seed = d.seed_dev ? (d.seed ^ __ldg(d.seed_dev)) : d.seed;
  return rng_u32(seed, d.site, idx) < d.thresh ? 0.f : d.scale;
}

__device__ __forceinline__ void row768_load_f32(float (&v)[24], const float* __restrict__ src, int lane) {
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const int col = (c * 32 + lane) * 8;
    const float4 a = *reinterpret_cast<const float4*>(src + col);
    const float4 b = *reinterpret_cast<const float4*>(src + col + 4);
    v[c * 8 + 0] = a.x; v[c * 8 + 1] = a.y; v[c * 8 + 2] = a.z; v[c * 8 + 3] = a.w;
    v[c * 8 + 4] = b.x; v[c * 8 + 5] = b.y; v[c * 8 + 6] = b.z; v[c * 8 + 7] = b.w;
  }
}
__device__ __forceinline__ void row768_store_f32(const float (&v)[24], float* dst, int lane) {
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const int col = (c * 32 + lane) * 8;
    *reinterpret_cast<float4*>(dst + col) = make_float4(v[c * 8 + 0], v[c * 8 + 1], v[c * 8 + 2], v[c * 8 + 3]);
    *reinterpret_cast<float4*>(dst + col + 4) = make_float4(v[c * 8 + 4], v[c * 8 + 5], v[c * 8 + 6], v[c * 8 + 7]);
  }
}
__device__ __forceinline__ void row768_store_bf16(const float (&v)[24], bf16* dst, int lane) {
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    uint4 u;
    u.x = pack_bf16x2(v[c * 8 + 0], v[c * 8 + 1]); u.y = pack_bf16x2(v[c * 8 + 2], v[c * 8 + 3]);
    u.z = pack_bf16x2(v[c * 8 + 4], v[c * 8 + 5]); u.w = pack_bf16x2(v[c * 8 + 6], v[c * 8 + 7]);
    *reinterpret_cast<uint4*>(dst + (c * 32 + lane) * 8) = u;
  }
}
// transposed copy: dstT[col * ldT + row]
__device__ __forceinline__ void row768_store_T(const float (&v)[24], bf16* dstT, int ldT, long long row, int lane) {
#pragma unroll
  for (int c = 0; c < 3; ++c)
#pragma unroll
    for (int j = 0; j < 8; ++j) dstT[static_cast<size_t>((c * 32 + lane) * 8 + j) * ldT + row] = __float2bfloat16(v[c * 8 + j]);
}
__device__ __forceinline__ void row768_stats(const float (&v)[24], float eps, float& mean, float& rstd) {
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 24; ++i) s += v[i];
  mean = warp_sum(s) * (1.0f / ENC_D);
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < 24; ++i) { const float d = v[i] - mean; ss = fmaf(d, d, ss); }
  rstd = rsqrtf(warp_sum(ss) * (1.0f / ENC_D) + eps);
}

// ------------------------------------------------------------------------------------------------------------------
// y = drop_out( LN( res + drop_a(a) ) )  — post-norm residual step (nn.TransformerDecoderLayer norm1/2/3, the outer
// fusion_layer_norm of fusionv3.py:47-49, BERT's attention.output / output LayerNorms)
// ------------------------------------------------------------------------------------------------------------------
struct AddLnParams {
  const float* a;    // [n, 768] or nullptr
  const float* res;  // [n, 768], or ONE row broadcast to every row (res_bcast)
  int res_bcast;
  const float *gamma, *beta;
  float eps;
  float* u_out;  // nullable: the pre-LayerNorm sum (kept for the backward pass)
  float* y_f32;  // nullable
  bf16* y_bf16;  // nullable
  bf16* yT;      // nullable: transposed bf16 copy, yT[col * ldT + rowT0 + row]
  int ldT, rowT0;
  long long n;
  Drop drop_a, drop_out;
};

__global__ void __launch_bounds__(256) add_ln_kernel(const AddLnParams p) {
  const int lane = threadIdx.x & 31;
  const long long row = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  if (row >= p.n) return;
  float v[24];
  row768_load_f32(v, p.res + (p.res_bcast ? 0 : row * ENC_D), lane);
  if (p.a != nullptr) {
    float a[24];
    row768_load_f32(a, p.a + row * ENC_D, lane);
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
      for (int j = 0; j < 8; ++j)
        v[c * 8 + j] += a[c * 8 + j] * drop_mul(p.drop_a, static_cast<uint32_t>(row * ENC_D + (c * 32 + lane) * 8 + j));
  }
  if (p.u_out) row768_store_f32(v, p.u_out + row * ENC_D, lane);
  float mean, rstd;
  row768_stats(v, p.eps, mean, rstd);
#pragma unroll
  for (int c = 0; c < 3; ++c)
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int col = (c * 32 + lane) * 8 + j;
      v[c * 8 + j] = ((v[c * 8 + j] - mean) * rstd * __ldg(p.gamma + col) + __ldg(p.beta + col)) *
                     drop_mul(p.drop_out, static_cast<uint32_t>(row * ENC_D + col));
    }
  if (p.y_f32) row768_store_f32(v, p.y_f32 + row * ENC_D, lane);
  if (p.y_bf16) row768_store_bf16(v, p.y_bf16 + row * ENC_D, lane);
  if (p.yT) row768_store_T(v, p.yT, p.ldT, p.rowT0 + row, lane);
}

// ------------------------------------------------------------------------------------------------------------------
// LayerNorm backward of the step above:  dy = (dy_a + dy_b) * mask_out ;  g = dy * gamma ;
//   du = rstd * (g - mean(g) - xhat * mean(g * xhat)) ; dgamma += dy * xhat ; dbeta += dy  (atomics over rows)
// du feeds the residual (fp32) and, masked by the branch's own dropout, the branch's backward GEMMs (bf16 + transposed).
// ------------------------------------------------------------------------------------------------------------------
struct LnBwdParams {
  const float *dy_a, *dy_b;  // dy_b nullable
  const float* u;            // pre-LayerNorm input of the forward step
  const float* gamma;
  float eps;
  float* du;        // nullable fp32 [n, 768]
  bf16* dub;        // nullable bf16 [n, 768]: du * mask of drop_a (what flows into the branch)
  bf16* dubT;       // nullable transposed copy of dub
  int ldT, rowT0;
  float *dgamma, *dbeta;  // [768], accumulated
  long long n;
  Drop drop_a, drop_out;
};

__global__ void __launch_bounds__(256) ln_bwd_kernel(const LnBwdParams p) {
  __shared__ float red[8][ENC_D];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long row = static_cast<long long>(blockIdx.x) * 8 + warp;
  const bool active = row < p.n;
  float u[24], dy[24], xg[24];  // xg: dy * xhat (gamma gradient of this row); dy keeps the beta gradient until the flush
#pragma unroll
  for (int i = 0; i < 24; ++i) u[i] = dy[i] = xg[i] = 0.f;
  if (active) {
    row768_load_f32(u, p.u + row * ENC_D, lane);
    row768_load_f32(dy, p.dy_a + row * ENC_D, lane);
    if (p.dy_b) {
      float b[24];
      row768_load_f32(b, p.dy_b + row * ENC_D, lane);
#pragma unroll
      for (int i = 0; i < 24; ++i) dy[i] += b[i];
    }
    float mean, rstd;
    row768_stats(u, p.eps, mean, rstd);
    float g[24], sg = 0.f, sgx = 0.f;
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int col = (c * 32 + lane) * 8 + j, i = c * 8 + j;
        dy[i] *= drop_mul(p.drop_out, static_cast<uint32_t>(row * ENC_D + col));
        u[i] = (u[i] - mean) * rstd;  // xhat
        xg[i] = dy[i] * u[i];
        g[i] = dy[i] * __ldg(p.gamma + col);
        sg += g[i];
        sgx = fmaf(g[i], u[i], sgx);
      }
    sg = warp_sum(sg) * (1.0f / ENC_D);
    sgx = warp_sum(sgx) * (1.0f / ENC_D);
#pragma unroll
    for (int i = 0; i < 24; ++i) g[i] = rstd * (g[i] - sg - u[i] * sgx);
    if (p.du) row768_store_f32(g, p.du + row * ENC_D, lane);
    if (p.dub || p.dubT) {
#pragma unroll
      for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int j = 0; j < 8; ++j) g[c * 8 + j] *= drop_mul(p.drop_a, static_cast<uint32_t>(row * ENC_D + (c * 32 + lane) * 8 + j));
      if (p.dub) row768_store_bf16(g, p.dub + row * ENC_D, lane);
      if (p.dubT) row768_store_T(g, p.dubT, p.ldT, p.rowT0 + row, lane);
    }
  }
  // gamma / beta gradients: sum the CTA's 8 rows in shared memory, then one atomic per column and CTA
  auto flush = [&](const float (&acc)[24], float* dst) {
    __syncthreads();
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
      for (int j = 0; j < 8; ++j) red[warp][(c * 32 + lane) * 8 + j] = acc[c * 8 + j];
    __syncthreads();
    for (int col = threadIdx.x; col < ENC_D; col += 256) {
      float v = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) v += red[w][col];
      atomicAdd(dst + col, v);
    }
  };
  flush(xg, p.dgamma);
  flush(dy, p.dbeta);
}

// ------------------------------------------------------------------------------------------------------------------
// elementwise row pass fp32 [n, C] -> bf16 [n, C] (+ transposed copy): plain cast, GELU forward, GELU backward, with a
// dropout mask per element (group 1) or per group of `group` columns (group 64 = per attention head: the dropout that
// nn.MultiheadAttention applies to the 1 x 1 attention matrix of the length-1 self-attention zeroes whole heads)
// ------------------------------------------------------------------------------------------------------------------
enum { ROWS_CAST = 0, ROWS_GELU_FWD = 1, ROWS_GELU_BWD = 2 };
struct RowsParams {
  const float* x;    // [n, C]
  const float* aux;  // GELU_BWD: the forward pre-activation f
  bf16* y;           // nullable
  bf16* yT;          // nullable
  int ldT, rowT0;
  long long n;
  int C, mode, group;
  Drop drop;
};
__device__ __forceinline__ float gelu_exact(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }
__device__ __forceinline__ float gelu_grad(float x) {
  return 0.5f * (1.0f + erff(x * 0.70710678118654752f)) + x * 0.3989422804014327f * __expf(-0.5f * x * x);
}
__global__ void __launch_bounds__(256) rows_kernel(const RowsParams p) {
  const long long i8 = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int per_row = p.C / 8;
  if (i8 >= p.n * per_row) return;
  const long long row = i8 / per_row;
  const int col = static_cast<int>(i8 - row * per_row) * 8;
  const float* xp = p.x + row * p.C + col;
  float v[8];
  {
    const float4 a = *reinterpret_cast<const float4*>(xp), b = *reinterpret_cast<const float4*>(xp + 4);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  }
  if (p.mode == ROWS_GELU_FWD) {
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = gelu_exact(v[j]);
  } else if (p.mode == ROWS_GELU_BWD) {
    const float* fp = p.aux + row * p.C + col;
    const float4 a = *reinterpret_cast<const float4*>(fp), b = *reinterpret_cast<const float4*>(fp + 4);
    const float f[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] *= gelu_grad(f[j]);
  }
  if (p.drop.p > 0.f) {
#pragma unroll
    for (int j = 0; j < 8; ++j)
      v[j] *= drop_mul(p.drop, static_cast<uint32_t>(p.group == 1 ? row * p.C + col + j : row * (p.C / p.group) + (col + j) / p.group));
  }
  if (p.y) {
    uint4 u;
    u.x = pack_bf16x2(v[0], v[1]); u.y = pack_bf16x2(v[2], v[3]); u.z = pack_bf16x2(v[4], v[5]); u.w = pack_bf16x2(v[6], v[7]);
    *reinterpret_cast<uint4*>(p.y + row * p.C + col) = u;
  }
  if (p.yT) {
#pragma unroll
    for (int j = 0; j < 8; ++j) p.yT[static_cast<size_t>(col + j) * p.ldT + p.rowT0 + row] = __float2bfloat16(v[j]);
  }
}

// dropout of bf16 rows in place (video_dropout / question_dropout on the embedded memory, fusionv3.py:190-191)
__global__ void __launch_bounds__(256) dropout_bf16_kernel(bf16* x, long long n_elems8, Drop drop) {
  const long long i8 = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i8 >= n_elems8) return;
  uint4 u = reinterpret_cast<uint4*>(x)[i8];
  uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    float2 f = unpack_bf16x2(w[k]);
    f.x *= drop_mul(drop, static_cast<uint32_t>(i8 * 8 + 2 * k));
    f.y *= drop_mul(drop, static_cast<uint32_t>(i8 * 8 + 2 * k + 1));
    w[k] = pack_bf16x2(f.x, f.y);
  }
  reinterpret_cast<uint4*>(x)[i8] = make_uint4(w[0], w[1], w[2], w[3]);
}

// ------------------------------------------------------------------------------------------------------------------
// single-query cross attention of the summarisation token (nn.MultiheadAttention inside the decoder layer, 12 heads x 64,
// memory = [video segment s ; text], no masks: fusionv3.py:45-46), forward with the probabilities kept, and backward.
// One warp per (row, head). Keys are read straight from the K/V GEMM output (kv_video / kv_text, row pitch ld_kv, K of
// this layer at column kcol, V at kcol + 768): a lane owns keys lane, lane + 32, ... for the score / dV / dK passes and
// two of the 64 head dims for the P V / dq passes.
// ------------------------------------------------------------------------------------------------------------------
constexpr int XA_MAXK = 256;  // Tv + Lt <= 256 (8 keys per lane)
struct XAttnParams {
  const float* q;  // [R, 768] fp32 (un-scaled; 1/8 applied here)
  const bf16 *kv_video, *kv_text;
  int ld_kv, kcol;
  int R, S, seg, Tv, Lt, n_cand;
  float* P;     // [R * 12, XA_MAXK] fp32 probabilities (before dropout)
  bf16* ctx;    // fwd out [R, 768]
  bf16* ctxT;   // fwd: transposed copy
  int ldT, rowT0;
  Drop drop;    // dropout on the probabilities
  // backward
  const float* dctx;  // [R, 768] fp32
  bf16* dq;           // [R, 768] bf16 (already scaled by 1/8)
  bf16* dqT;
  bf16 *dkv_video, *dkv_text;  // same layout as kv_*; text rows of THIS recurrent step (accumulated over steps by the caller)
};

__device__ __forceinline__ const bf16* xa_key_row(const XAttnParams& p, int b, int j) {
  return j < p.Tv ? p.kv_video + ((static_cast<size_t>(b / p.n_cand) * p.S + p.seg) * p.Tv + j) * p.ld_kv
                  : p.kv_text + (static_cast<size_t>(b) * p.Lt + (j - p.Tv)) * p.ld_kv;
}
__device__ __forceinline__ bf16* xa_dkey_row(const XAttnParams& p, int b, int j) {
  return j < p.Tv ? p.dkv_video + ((static_cast<size_t>(b / p.n_cand) * p.S + p.seg) * p.Tv + j) * p.ld_kv
                  : p.dkv_text + (static_cast<size_t>(b) * p.Lt + (j - p.Tv)) * p.ld_kv;
}

// One CTA (128 threads) per (row, head). The head's K and V slices (n_keys x 64 bf16 each) are staged in shared memory with
// coalesced 16-byte loads (row pitch 144 B: conflict-free when a thread reads one key's whole row); thread = key for the
// score / dV / dK passes, thread = (head dim, key half) for the P V / dq passes.
constexpr int XA_PITCH = 72;  // bf16 elements per staged key row
constexpr int XA_THREADS = 128;
constexpr int XA_SMEM = 2 * XA_MAXK * XA_PITCH * 2 + (XA_MAXK + 64 + 64 + 2 * 64 + 8) * 4;

struct XaSmem {
  bf16* K;
  bf16* V;
  float* s;     // [XA_MAXK] scores / probabilities / dS
  float* q;     // [64]
  float* dc;    // [64]
  float* part;  // [2][64] partial sums of the two key halves
  float* red;   // [8] cross-warp reduction scratch
};
__device__ __forceinline__ XaSmem xa_carve(uint8_t* smem) {
  XaSmem m;
  m.K = reinterpret_cast<bf16*>(smem);
  m.V = m.K + XA_MAXK * XA_PITCH;
  m.s = reinterpret_cast<float*>(m.V + XA_MAXK * XA_PITCH);
  m.q = m.s + XA_MAXK;
  m.dc = m.q + 64;
  m.part = m.dc + 64;
  m.red = m.part + 128;
  return m;
}
__device__ __forceinline__ void xa_stage_kv(const XAttnParams& p, const XaSmem& m, int b, int head, int n_keys) {
  for (int c = threadIdx.x; c < n_keys * 16; c += XA_THREADS) {
    const int j = c >> 4, part = (c >> 3) & 1, ch = c & 7;  // part 0: K, 1: V
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(xa_key_row(p, b, j) + p.kcol + part * ENC_D + head * 64) + ch);
    *reinterpret_cast<uint4*>((part ? m.V : m.K) + j * XA_PITCH + ch * 8) = v;
  }
}
__device__ __forceinline__ float xa_dot64(const bf16* row, const float* vec) {
  float acc = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const uint4 u = *reinterpret_cast<const uint4*>(row + 8 * i);
    float2 f;
    f = unpack_bf16x2(u.x); acc = fmaf(vec[8 * i + 0], f.x, acc); acc = fmaf(vec[8 * i + 1], f.y, acc);
    f = unpack_bf16x2(u.y); acc = fmaf(vec[8 * i + 2], f.x, acc); acc = fmaf(vec[8 * i + 3], f.y, acc);
    f = unpack_bf16x2(u.z); acc = fmaf(vec[8 * i + 4], f.x, acc); acc = fmaf(vec[8 * i + 5], f.y, acc);
    f = unpack_bf16x2(u.w); acc = fmaf(vec[8 * i + 6], f.x, acc); acc = fmaf(vec[8 * i + 7], f.y, acc);
  }
  return acc;
}
// block-wide sum / max over the 4 warps (result in every thread)
template <bool IS_MAX>
__device__ __forceinline__ float xa_block_reduce(float v, float* red) {
  v = IS_MAX ? warp_max(v) : warp_sum(v);
  __syncthreads();  // `red` may still be read from a previous reduction
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = red[0];
#pragma unroll
  for (int w = 1; w < XA_THREADS / 32; ++w) r = IS_MAX ? fmaxf(r, red[w]) : r + red[w];
  return r;
}
// out[d] = scale * sum_j w[j] M[j][d] for the 64 head dims: thread (d = tid & 63, half = tid >> 6) takes every other key
__device__ __forceinline__ float xa_weighted_rows(const bf16* M, const float* w, int n_keys, float* part) {
  const int d = threadIdx.x & 63, half = threadIdx.x >> 6;
  float acc = 0.f;
  for (int j = half; j < n_keys; j += 2) acc = fmaf(w[j], __bfloat162float(M[j * XA_PITCH + d]), acc);
  part[half * 64 + d] = acc;
  __syncthreads();
  return part[d] + part[64 + d];
}

__global__ void __launch_bounds__(XA_THREADS) xattn_fwd_kernel(const XAttnParams p) {
  extern __shared__ __align__(16) uint8_t xa_smem[];
  const XaSmem m = xa_carve(xa_smem);
  const int unit = blockIdx.x, b = unit / 12, head = unit % 12, tid = threadIdx.x;
  const int n_keys = p.Tv + p.Lt;
  xa_stage_kv(p, m, b, head, n_keys);
  if (tid < 64) m.q[tid] = __ldg(p.q + static_cast<size_t>(b) * ENC_D + head * 64 + tid) * 0.125f;
  __syncthreads();
  float s[2], mx = -INFINITY;
#pragma unroll
  for (int t = 0; t < 2; ++t) {
    const int j = t * XA_THREADS + tid;
    s[t] = j < n_keys ? xa_dot64(m.K + j * XA_PITCH, m.q) : -INFINITY;
    mx = fmaxf(mx, s[t]);
  }
  mx = xa_block_reduce<true>(mx, m.red);
  float sum = 0.f;
#pragma unroll
  for (int t = 0; t < 2; ++t) {
    s[t] = (t * XA_THREADS + tid < n_keys) ? __expf(s[t] - mx) : 0.f;
    sum += s[t];
  }
  sum = xa_block_reduce<false>(sum, m.red);
  const float inv = 1.0f / sum;
  float* Prow = p.P + static_cast<size_t>(unit) * XA_MAXK;
#pragma unroll
  for (int t = 0; t < 2; ++t) {
    const int j = t * XA_THREADS + tid;
    const float pj = s[t] * inv;
    Prow[j] = pj;
    m.s[j] = pj * drop_mul(p.drop, static_cast<uint32_t>(unit * XA_MAXK + j));
  }
  __syncthreads();
  const float o = xa_weighted_rows(m.V, m.s, n_keys, m.part);
  if (tid < 64) {
    const int col = head * 64 + tid;
    p.ctx[static_cast<size_t>(b) * ENC_D + col] = __float2bfloat16(o);
    if (p.ctxT) p.ctxT[static_cast<size_t>(col) * p.ldT + p.rowT0 + b] = __float2bfloat16(o);
  }
}

template <bool ATOMIC_VIDEO>
__global__ void __launch_bounds__(XA_THREADS) xattn_bwd_kernel(const XAttnParams p) {
  extern __shared__ __align__(16) uint8_t xa_smem[];
  const XaSmem m = xa_carve(xa_smem);
  const int unit = blockIdx.x, b = unit / 12, head = unit % 12, tid = threadIdx.x;
  const int n_keys = p.Tv + p.Lt;
  xa_stage_kv(p, m, b, head, n_keys);
  if (tid < 64) {
    m.q[tid] = __ldg(p.q + static_cast<size_t>(b) * ENC_D + head * 64 + tid) * 0.125f;
    m.dc[tid] = __ldg(p.dctx + static_cast<size_t>(b) * ENC_D + head * 64 + tid);
  }
  __syncthreads();
  const float* Prow = p.P + static_cast<size_t>(unit) * XA_MAXK;
  // stores of one key's 64 gradient values (128 B): plain for rows this (row, head) owns, atomic for the video rows that
  // the candidates of a multiple-choice clip share
  auto store_row = [&](bf16* dst, float w, const float* vec, bool shared_row) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (ATOMIC_VIDEO && shared_row) {
#pragma unroll
        for (int k = 0; k < 4; ++k)
          atomicAdd(reinterpret_cast<__nv_bfloat162*>(dst + 8 * i + 2 * k), __floats2bfloat162_rn(w * vec[8 * i + 2 * k], w * vec[8 * i + 2 * k + 1]));
      } else {
        uint4 o;
        o.x = pack_bf16x2(w * vec[8 * i + 0], w * vec[8 * i + 1]); o.y = pack_bf16x2(w * vec[8 * i + 2], w * vec[8 * i + 3]);
        o.z = pack_bf16x2(w * vec[8 * i + 4], w * vec[8 * i + 5]); o.w = pack_bf16x2(w * vec[8 * i + 6], w * vec[8 * i + 7]);
        *reinterpret_cast<uint4*>(dst + 8 * i) = o;
      }
    }
  };
  // pass 1 (thread = key): dP'_j = dctx . V_j ; dV_j = p'_j dctx ; D = sum_j P_j dP_j
  float pj[2], dp[2], dsum = 0.f;
#pragma unroll
  for (int t = 0; t < 2; ++t) {
    const int j = t * XA_THREADS + tid;
    pj[t] = dp[t] = 0.f;
    if (j < n_keys) {
      pj[t] = Prow[j];
      const float mk = drop_mul(p.drop, static_cast<uint32_t>(unit * XA_MAXK + j));
      store_row(xa_dkey_row(p, b, j) + p.kcol + ENC_D + head * 64, pj[t] * mk, m.dc, j < p.Tv);
      dp[t] = xa_dot64(m.V + j * XA_PITCH, m.dc) * mk;
      dsum = fmaf(pj[t], dp[t], dsum);
    }
  }
  dsum = xa_block_reduce<false>(dsum, m.red);
  // pass 2 (thread = key): dS_j = P_j (dP_j - D) ; dK_j = dS_j q / 8
#pragma unroll
  for (int t = 0; t < 2; ++t) {
    const int j = t * XA_THREADS + tid;
    const float ds = pj[t] * (dp[t] - dsum);
    m.s[j] = ds;
    if (j < n_keys) store_row(xa_dkey_row(p, b, j) + p.kcol + head * 64, ds, m.q, j < p.Tv);
  }
  __syncthreads();
  // pass 3: dq = sum_j dS_j K_j / 8
  const float g = xa_weighted_rows(m.K, m.s, n_keys, m.part) * 0.125f;
  if (tid < 64) {
    const int col = head * 64 + tid;
    p.dq[static_cast<size_t>(b) * ENC_D + col] = __float2bfloat16(g);
    if (p.dqT) p.dqT[static_cast<size_t>(col) * p.ldT + p.rowT0 + b] = __float2bfloat16(g);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// out[i] (+)= sum_{c < cols} src[i * ld + c]   (bias gradients from the transposed dY^T operands: one warp per row)
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) rowsum_bf16_kernel(const bf16* __restrict__ src, int ld, int cols, float* out, int rows,
                                                          int accumulate) {
  const int lane = threadIdx.x & 31;
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (row >= rows) return;
  float s = 0.f;
  for (int c = lane; c < cols; c += 32) s += __bfloat162float(src[static_cast<size_t>(row) * ld + c]);
  s = warp_sum(s);
  if (lane == 0) out[row] = accumulate ? out[row] + s : s;
}

// out[c] (+)= sum_r src[r * ld + c] for fp32 / bf16 row-major inputs with many rows (projection bias, K/V bias): one thread
// per column and row slab, atomics across slabs
template <typename T>
__global__ void __launch_bounds__(256) colsum_kernel(const T* __restrict__ src, long long rows, int cols, long long ld, float* out,
                                                     int rows_per_block) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= cols) return;
  const long long r0 = static_cast<long long>(blockIdx.y) * rows_per_block;
  const long long r1 = r0 + rows_per_block < rows ? r0 + rows_per_block : rows;
  float s = 0.f;
  for (long long r = r0; r < r1; ++r) s += static_cast<float>(src[r * ld + c]);
  atomicAdd(out + c, s);
}

// dst[c * ld_dst + r] = src[r * ld_src + c]  (bf16, 64 x 64 tiles through shared memory)
__global__ void __launch_bounds__(256) transpose_bf16_kernel(const bf16* __restrict__ src, long long rows, int cols, long long ld_src,
                                                             bf16* __restrict__ dst, long long ld_dst) {
  __shared__ bf16 tile[64][66];
  const long long r0 = static_cast<long long>(blockIdx.y) * 64;
  const int c0 = blockIdx.x * 64;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  for (int i = ty; i < 64; i += 8) {
    const long long r = r0 + i;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const int c = c0 + tx * 2 + k;
      tile[i][tx * 2 + k] = (r < rows && c < cols) ? src[r * ld_src + c] : __float2bfloat16(0.f);
    }
  }
  __syncthreads();
  for (int i = ty; i < 64; i += 8) {
    const int c = c0 + i;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const long long r = r0 + tx * 2 + k;
      if (c < cols && r < rows) dst[static_cast<long long>(c) * ld_dst + r] = tile[tx * 2 + k][i];
    }
  }
}

// dst = a + b (+ c) over bf16 (sum of the per-step text-key gradients)
__global__ void __launch_bounds__(256) add_bf16_kernel(bf16* dst, const bf16* a, const bf16* b, const bf16* c, long long n8) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n8) return;
  const uint4 ua = reinterpret_cast<const uint4*>(a)[i], ub = reinterpret_cast<const uint4*>(b)[i];
  uint4 uc = make_uint4(0, 0, 0, 0);
  if (c) uc = reinterpret_cast<const uint4*>(c)[i];
  const uint32_t wa[4] = {ua.x, ua.y, ua.z, ua.w}, wb[4] = {ub.x, ub.y, ub.z, ub.w}, wc[4] = {uc.x, uc.y, uc.z, uc.w};
  uint32_t o[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float2 fa = unpack_bf16x2(wa[k]), fb = unpack_bf16x2(wb[k]), fc = unpack_bf16x2(wc[k]);
    o[k] = pack_bf16x2(fa.x + fb.x + fc.x, fa.y + fb.y + fc.y);
  }
  reinterpret_cast<uint4*>(dst)[i] = make_uint4(o[0], o[1], o[2], o[3]);
}

// ------------------------------------------------------------------------------------------------------------------
// backward of VideoPosEmbed / TextPosEmbed (embedding.py:47-63, :17-23): dy fp32 [rows, 768] (gradient of the embedded
// memory rows = output of the dX GEMM of the K/V projection) -> LayerNorm backward through the recomputed pre-LN sum;
// dproj bf16 for the projection_layer's gradient GEMMs, table gradients by atomics
// ------------------------------------------------------------------------------------------------------------------
struct PosBwdParams {
  const float* dy;
  const bf16* proj;  // video: projection output [B*S*T*P, 768]; text: nullptr
  const float* text_f32;  // text features [Bt, L, 768] fp32 or nullptr
  const bf16* text_bf16;  // or bf16
  const float *emb_cls, *emb_pos, *emb_len, *emb_clip, *gamma;
  float eps;
  bf16* dproj;  // video only: [B*S*T*P, 768]
  float *d_cls, *d_pos, *d_len, *d_clip, *dgamma, *dbeta;
  int B, S, T, P;  // text: B = Bt, P = L, S = T = 1
  int is_text;
  Drop drop;  // dropout that was applied to the embedded rows
};
// One CTA (8 warps) per frame = the P + 1 rows that share (t, s): LayerNorm-parameter and emb_len / emb_clip gradients are
// summed in registers over the frame's rows and across the warps in shared memory before ONE atomic per column and CTA
// (14 400 video rows would otherwise queue 14 400 atomics on each of the 768 gamma addresses); emb_pos is per row.
__global__ void __launch_bounds__(256) posembed_bwd_kernel(const PosBwdParams p) {
  __shared__ float red[8][ENC_D];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long frame = blockIdx.x;  // (b * S + s) * T + t
  const int t = static_cast<int>(frame % p.T);
  const int s = static_cast<int>((frame / p.T) % p.S);
  float acc_g[24], acc_b[24], acc_u[24];
#pragma unroll
  for (int i = 0; i < 24; ++i) acc_g[i] = acc_b[i] = acc_u[i] = 0.f;
  for (int pp = warp; pp <= p.P; pp += 8) {
    const long long row = frame * (p.P + 1) + pp;
    float v[24];
#pragma unroll
    for (int i = 0; i < 24; ++i) v[i] = 0.f;
    if (pp == 0) row768_add_f32(v, p.emb_cls, lane);
    else if (!p.is_text) row768_add_bf16(v, p.proj + (frame * p.P + (pp - 1)) * ENC_D, lane);
    else if (p.text_f32) row768_add_f32(v, p.text_f32 + (frame * p.P + (pp - 1)) * ENC_D, lane);
    else row768_add_bf16(v, p.text_bf16 + (frame * p.P + (pp - 1)) * ENC_D, lane);
    row768_add_f32(v, p.emb_pos + static_cast<size_t>(pp) * ENC_D, lane);
    if (!p.is_text) {
      row768_add_f32(v, p.emb_len + static_cast<size_t>(t) * ENC_D, lane);
      row768_add_f32(v, p.emb_clip + static_cast<size_t>(s) * ENC_D, lane);
    }
    float mean, rstd;
    row768_stats(v, p.eps, mean, rstd);
    float dy[24];
    row768_load_f32(dy, p.dy + row * ENC_D, lane);
    float sg = 0.f, sgx = 0.f;
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int col = (c * 32 + lane) * 8 + j, i = c * 8 + j;
        dy[i] *= drop_mul(p.drop, static_cast<uint32_t>(row * ENC_D + col));
        v[i] = (v[i] - mean) * rstd;
        acc_g[i] = fmaf(dy[i], v[i], acc_g[i]);
        acc_b[i] += dy[i];
        dy[i] *= __ldg(p.gamma + col);
        sg += dy[i];
        sgx = fmaf(dy[i], v[i], sgx);
      }
    sg = warp_sum(sg) * (1.0f / ENC_D);
    sgx = warp_sum(sgx) * (1.0f / ENC_D);
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int col = (c * 32 + lane) * 8 + j, i = c * 8 + j;
        const float du = rstd * (dy[i] - sg - v[i] * sgx);
        dy[i] = du;
        acc_u[i] += du;
        atomicAdd(p.d_pos + static_cast<size_t>(pp) * ENC_D + col, du);
        if (pp == 0) atomicAdd(p.d_cls + col, du);
      }
    if (!p.is_text && pp > 0) row768_store_bf16(dy, p.dproj + (frame * p.P + (pp - 1)) * ENC_D, lane);
  }
  // cross-warp sums, one quantity at a time through the same shared buffer
  auto flush = [&](const float (&acc)[24], float* dst0, float* dst1) {
    __syncthreads();
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
      for (int j = 0; j < 8; ++j) red[warp][(c * 32 + lane) * 8 + j] = acc[c * 8 + j];
    __syncthreads();
    for (int col = threadIdx.x; col < ENC_D; col += 256) {
      float v = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) v += red[w][col];
      atomicAdd(dst0 + col, v);
      if (dst1) atomicAdd(dst1 + col, v);
    }
  };
  flush(acc_g, p.dgamma, nullptr);
  flush(acc_b, p.dbeta, nullptr);
  if (!p.is_text) flush(acc_u, p.d_len + static_cast<size_t>(t) * ENC_D, p.d_clip + static_cast<size_t>(s) * ENC_D);
}

// ------------------------------------------------------------------------------------------------------------------
// Skinny GEMM for the token path of the training step: out[M, N] fp32 = A[M, K] bf16 x W[N, K]^T (+ bias), M = a few dozen
// rows (the summarisation tokens of a batch), K in {768, 3072}. A 128-row tcgen05 tile would be three-quarters padding and
// the persistent GEMM's prologue (TMEM allocation, barrier ring, descriptor fetch) costs more than this whole product, which
// is bound by streaming W once: a CTA owns 8 output columns of a 32-row tile, its 8 warps split K and stream their slices of
// the 8 weight rows straight into mma.sync m16n8k16 B fragments (16-byte loads), the 32 activation rows sit in shared memory.
// ------------------------------------------------------------------------------------------------------------------
constexpr int SK_ROWS = 32, SK_WARPS = 8, SK_THREADS = 256;
template <int NCH>  // 32-wide k chunks per warp: K = 8 * 32 * NCH
__global__ void __launch_bounds__(SK_THREADS) skinny_gemm_kernel(const bf16* __restrict__ A, int lda, const bf16* __restrict__ W, int ldw,
                                                                 const float* __restrict__ bias, float* __restrict__ out, int ldo, int M,
                                                                 int N) {
  constexpr int K = SK_WARPS * 32 * NCH;
  constexpr int PITCH = K + 32;  // bf16 elements; (PITCH / 2) % 32 == 16 words: conflict-free 16-byte fragment loads
  extern __shared__ __align__(16) uint8_t sk_smem[];
  bf16* sX = reinterpret_cast<bf16*>(sk_smem);
  float* sRed = reinterpret_cast<float*>(sk_smem + SK_ROWS * PITCH * 2);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t4 = lane & 3;
  const int n0 = blockIdx.x * 8, r_base = blockIdx.y * SK_ROWS;
  const int k_begin = warp * (NCH * 32);
  // this warp's K slice of the tile's 8 weight rows: issued first, they do not depend on the activations
  uint4 wreg[NCH];
  {
    const int nrow = min(n0 + g, N - 1);
    const bf16* wrow = W + static_cast<size_t>(nrow) * ldw + k_begin + 8 * t4;
#pragma unroll
    for (int c = 0; c < NCH; ++c) wreg[c] = __ldg(reinterpret_cast<const uint4*>(wrow + c * 32));
  }
  for (int c = tid; c < SK_ROWS * (K / 8); c += SK_THREADS) {
    const int r = c / (K / 8), k = (c - r * (K / 8)) * 8;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (r_base + r < M) v = *reinterpret_cast<const uint4*>(A + static_cast<size_t>(r_base + r) * lda + k);
    *reinterpret_cast<uint4*>(sX + static_cast<size_t>(r) * PITCH + k) = v;
  }
  __syncthreads();
  const bf16* xa0 = sX + static_cast<size_t>(g) * PITCH + k_begin + 8 * t4;
  float acc[2][4];
#pragma unroll
  for (int mi = 0; mi < 2; ++mi) acc[mi][0] = acc[mi][1] = acc[mi][2] = acc[mi][3] = 0.f;
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
#pragma unroll
    for (int mi = 0; mi < 2; ++mi) {
      const uint4 xlo = *reinterpret_cast<const uint4*>(xa0 + static_cast<size_t>(mi * 16) * PITCH + c * 32);
      const uint4 xhi = *reinterpret_cast<const uint4*>(xa0 + static_cast<size_t>(mi * 16 + 8) * PITCH + c * 32);
      mma16816(acc[mi], xlo.x, xhi.x, xlo.y, xhi.y, wreg[c].x, wreg[c].y);
      mma16816(acc[mi], xlo.z, xhi.z, xlo.w, xhi.w, wreg[c].z, wreg[c].w);
    }
  }
  // cross-warp K reduction, bias, store
#pragma unroll
  for (int mi = 0; mi < 2; ++mi) {
    float* r = sRed + (warp * SK_ROWS + mi * 16 + g) * 8 + 2 * t4;
    r[0] = acc[mi][0]; r[1] = acc[mi][1];
    r[8 * 8] = acc[mi][2]; r[8 * 8 + 1] = acc[mi][3];
  }
  __syncthreads();
  {
    const int r = tid >> 3, col = tid & 7;
    float v = 0.f;
#pragma unroll
    for (int w = 0; w < SK_WARPS; ++w) v += sRed[(w * SK_ROWS + r) * 8 + col];
    const int row = r_base + r, n = n0 + col;
    if (row < M && n < N) out[static_cast<size_t>(row) * ldo + n] = v + (bias ? __ldg(bias + n) : 0.f);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// BERT (HF BertModel as text.py:9-17 uses it): embeddings + LayerNorm, and the masked multi-head self-attention
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) bert_embed_ln_kernel(const long long* __restrict__ ids, const long long* __restrict__ type_ids,
                                                            const float* __restrict__ word, const float* __restrict__ pos,
                                                            const float* __restrict__ type, const float* __restrict__ gamma,
                                                            const float* __restrict__ beta, float eps, float* out_f32, bf16* out_bf16,
                                                            long long n, int L, int vocab, int n_types) {
  const int lane = threadIdx.x & 31;
  const long long row = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  if (row >= n) return;
  long long id = ids[row], tt = type_ids ? type_ids[row] : 0;
  id = id < 0 ? 0 : (id >= vocab ? vocab - 1 : id);  // host code validates; never index out of bounds
  tt = tt < 0 ? 0 : (tt >= n_types ? n_types - 1 : tt);
  float v[24];
#pragma unroll
  for (int i = 0; i < 24; ++i) v[i] = 0.f;
  row768_add_f32(v, word + id * ENC_D, lane);
  row768_add_f32(v, pos + (row % L) * ENC_D, lane);
  row768_add_f32(v, type + tt * ENC_D, lane);
  row768_ln_store(v, gamma, beta, eps, lane, out_bf16 + row * ENC_D, out_f32 + row * ENC_D);
}

// one CTA (4 warps) per (sequence, head); K / V head slices in shared memory (row pitch 72 bf16: conflict-free 16-byte
// reads with lane = key); keys with attention_mask == 0 get -inf (HF adds finfo.min: the same softmax); L <= 64
constexpr int BA_MAXL = 64;
__global__ void __launch_bounds__(128) bert_attention_kernel(const bf16* __restrict__ qkv, const long long* __restrict__ mask,
                                                             bf16* __restrict__ out, int L, int n_heads) {
  __shared__ __align__(16) bf16 sK[BA_MAXL][72];
  __shared__ __align__(16) bf16 sV[BA_MAXL][72];
  __shared__ float sQ[4][64];
  const int seq = blockIdx.x / n_heads, head = blockIdx.x % n_heads;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int D = n_heads * 64;
  const bf16* base = qkv + static_cast<size_t>(seq) * L * 3 * D;
  for (int i = threadIdx.x; i < L * 8; i += 128) {
    const int j = i >> 3, ch = i & 7;
    *reinterpret_cast<uint4*>(&sK[j][ch * 8]) = *reinterpret_cast<const uint4*>(base + static_cast<size_t>(j) * 3 * D + D + head * 64 + ch * 8);
    *reinterpret_cast<uint4*>(&sV[j][ch * 8]) = *reinterpret_cast<const uint4*>(base + static_cast<size_t>(j) * 3 * D + 2 * D + head * 64 + ch * 8);
  }
  __syncthreads();
  bool keep[2];
#pragma unroll
  for (int t = 0; t < 2; ++t) {
    const int j = t * 32 + lane;
    keep[t] = j < L && (mask == nullptr || mask[static_cast<size_t>(seq) * L + j] != 0);
  }
  for (int i = warp; i < L; i += 4) {
    {
      const float2 f = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(base + static_cast<size_t>(i) * 3 * D + head * 64 + 2 * lane));
      sQ[warp][2 * lane] = f.x * 0.125f;
      sQ[warp][2 * lane + 1] = f.y * 0.125f;
    }
    __syncwarp();
    float s[2], mx = -INFINITY;
#pragma unroll
    for (int t = 0; t < 2; ++t) {
      const int j = t * 32 + lane;
      s[t] = -INFINITY;
      if (j < L) {
        float acc = 0.f;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const uint4 u = *reinterpret_cast<const uint4*>(&sK[j][c * 8]);
          float2 f;
          f = unpack_bf16x2(u.x); acc = fmaf(sQ[warp][c * 8 + 0], f.x, acc); acc = fmaf(sQ[warp][c * 8 + 1], f.y, acc);
          f = unpack_bf16x2(u.y); acc = fmaf(sQ[warp][c * 8 + 2], f.x, acc); acc = fmaf(sQ[warp][c * 8 + 3], f.y, acc);
          f = unpack_bf16x2(u.z); acc = fmaf(sQ[warp][c * 8 + 4], f.x, acc); acc = fmaf(sQ[warp][c * 8 + 5], f.y, acc);
          f = unpack_bf16x2(u.w); acc = fmaf(sQ[warp][c * 8 + 6], f.x, acc); acc = fmaf(sQ[warp][c * 8 + 7], f.y, acc);
        }
        // a fully masked row cannot occur ([CLS] is always attended); masked keys contribute exactly zero
        s[t] = keep[t] ? acc : -INFINITY;
        mx = fmaxf(mx, s[t]);
      }
    }
    mx = warp_max(mx);
    float sum = 0.f;
#pragma unroll
    for (int t = 0; t < 2; ++t) {
      s[t] = (s[t] == -INFINITY) ? 0.f : __expf(s[t] - mx);
      sum += s[t];
    }
    sum = warp_sum(sum);
    const float inv = sum > 0.f ? 1.0f / sum : 0.f;
    float o0 = 0.f, o1 = 0.f;
#pragma unroll
    for (int t = 0; t < 2; ++t) {
      if (t * 32 >= L) break;
      for (int l = 0; l < 32 && t * 32 + l < L; ++l) {
        const float pj = __shfl_sync(0xffffffffu, s[t], l);
        const float2 f = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(&sV[t * 32 + l][2 * lane]));
        o0 = fmaf(pj, f.x, o0);
        o1 = fmaf(pj, f.y, o1);
      }
    }
    *reinterpret_cast<uint32_t*>(out + (static_cast<size_t>(seq) * L + i) * D + head * 64 + 2 * lane) = pack_bf16x2(o0 * inv, o1 * inv);
    __syncwarp();
  }
}

}  // namespace lrce

using namespace lrce;

static inline unsigned blocks_for_rows(long long rows) { return static_cast<unsigned>((rows + 7) / 8); }  // 8 warps per CTA

extern "C" int lrce_add_ln_768(const float* a, const float* res, int res_bcast, const float* gamma, const float* beta, float eps,
                               float* u_out, float* y_f32, void* y_bf16, void* yT, int ldT, int rowT0, long long n, float p_a,
                               int site_a, float p_out, int site_out, unsigned long long seed, const unsigned long long* seed_dev, void* stream) {
  int rc = require_sm100();
  if (rc != LRCE_OK) return rc;
  LRCE_REQUIRE(res && gamma && beta && n > 0 && (y_f32 || y_bf16 || yT), "lrce_add_ln_768: bad arguments");
  AddLnParams p;
  p.a = a; p.res = res; p.res_bcast = res_bcast; p.gamma = gamma; p.beta = beta; p.eps = eps; p.u_out = u_out; p.y_f32 = y_f32;
  p.y_bf16 = reinterpret_cast<bf16*>(y_bf16); p.yT = reinterpret_cast<bf16*>(yT); p.ldT = ldT; p.rowT0 = rowT0; p.n = n;
  p.drop_a = make_drop(p_a, seed, seed_dev, site_a);
  p.drop_out = make_drop(p_out, seed, seed_dev, site_out);
  add_ln_kernel<<<blocks_for_rows(n), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(p);
  return check_launch("add_ln_kernel");
}

extern "C" int lrce_ln_bwd_768(const float* dy_a, const float* dy_b, const float* u, const float* gamma, float eps, float* du,
                               void* dub, void* dubT, int ldT, int rowT0, float* dgamma, float* dbeta, long long n, float p_a,
                               int site_a, float p_out, int site_out, unsigned long long seed, const unsigned long long* seed_dev, void* stream) {
  int rc = require_sm100();
  if (rc != LRCE_OK) return rc;
  LRCE_REQUIRE(dy_a && u && gamma && dgamma && dbeta && n > 0, "lrce_ln_bwd_768: bad arguments");
  LnBwdParams p;
  p.dy_a = dy_a; p.dy_b = dy_b; p.u = u; p.gamma = gamma; p.eps = eps; p.du = du; p.dub = reinterpret_cast<bf16*>(dub);
  p.dubT = reinterpret_cast<bf16*>(dubT); p.ldT = ldT; p.rowT0 = rowT0; p.dgamma = dgamma; p.dbeta = dbeta; p.n = n;
  p.drop_a = make_drop(p_a, seed, seed_dev, site_a);
  p.drop_out = make_drop(p_out, seed, seed_dev, site_out);
  ln_bwd_kernel<<<blocks_for_rows(n), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(p);
  return check_launch("ln_bwd_kernel");
}

extern "C" int lrce_rows_f32_to_bf16(const float* x, const float* aux, void* y, void* yT, int ldT, int rowT0, long long n, int C,
                                     int mode, int group, float p_drop, int site, unsigned long long seed, const unsigned long long* seed_dev, void* stream) {
  int rc = require_sm100();
  if (rc != LRCE_OK) return rc;
  LRCE_REQUIRE(x && (y || yT) && n > 0 && C > 0 && C % 8 == 0 && mode >= 0 && mode <= 2 && group >= 1 && C % group == 0,
               "lrce_rows_f32_to_bf16: bad arguments (C=%d mode=%d group=%d)", C, mode, group);
  LRCE_REQUIRE(mode != ROWS_GELU_BWD || aux, "lrce_rows_f32_to_bf16: GELU backward needs the forward pre-activation");
  RowsParams p;
  p.x = x; p.aux = aux; p.y = reinterpret_cast<bf16*>(y); p.yT = reinterpret_cast<bf16*>(yT); p.ldT = ldT; p.rowT0 = rowT0;
  p.n = n; p.C = C; p.mode = mode; p.group = group; p.drop = make_drop(p_drop, seed, seed_dev, site);
  const long long threads = n * (C / 8);
  rows_kernel<<<static_cast<unsigned>((threads + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(p);
  return check_launch("rows_kernel");
}

extern "C" int lrce_dropout_bf16(void* x, long long n_elems, float p_drop, int site, unsigned long long seed, const unsigned long long* seed_dev, void* stream) {
  int rc = require_sm100();
  if (rc != LRCE_OK) return rc;
  LRCE_REQUIRE(x && n_elems > 0 && n_elems % 8 == 0, "lrce_dropout_bf16: bad arguments");
  if (!(p_drop > 0.f)) return LRCE_OK;
  dropout_bf16_kernel<<<static_cast<unsigned>((n_elems / 8 + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<bf16*>(x), n_elems / 8, make_drop(p_drop, seed, seed_dev, site));
  return check_launch("dropout_bf16_kernel");
}

static int xa_configure() {
  static thread_local uint64_t configured = 0;  // one bit per device
  if (needs_device_setup(&configured)) {
    cudaError_t e = cudaFuncSetAttribute(xattn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, XA_SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(xattn_bwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, XA_SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(xattn_bwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, XA_SMEM);
    if (e != cudaSuccess) {
      set_error("cudaFuncSetAttribute(xattn kernels, smem=%d): %s", XA_SMEM, cudaGetErrorString(e));
      return LRCE_ECUDA;
    }
    mark_device_setup(&configured);
  }
  return LRCE_OK;
}

static int fill_xattn(XAttnParams* p, const float* q, const void* kv_video, const void* kv_text, int ld_kv, int kcol, int R, int S,
                      int seg, int Tv, int Lt, int n_cand, float* P, int ldT, int rowT0, float p_drop, int site,
                      unsigned long long seed, const unsigned long long* seed_dev) {
  LRCE_REQUIRE(q && kv_video && kv_text && P && R > 0 && S > 0 && seg >= 0 && seg < S && Tv > 0 && Lt > 0 && n_cand > 0 &&
                   R % n_cand == 0 && Tv + Lt <= XA_MAXK && ld_kv % 8 == 0 && kcol % 8 == 0,
               "cross attention: bad arguments (R=%d S=%d seg=%d Tv=%d Lt=%d cand=%d)", R, S, seg, Tv, Lt, n_cand);
  p->q = q; p->kv_video = reinterpret_cast<const bf16*>(kv_video); p->kv_text = reinterpret_cast<const bf16*>(kv_text);
  p->ld_kv = ld_kv; p->kcol = kcol; p->R = R; p->S = S; p->seg = seg; p->Tv = Tv; p->Lt = Lt; p->n_cand = n_cand; p->P = P;
  p->ldT = ldT; p->rowT0 = rowT0; p->drop = make_drop(p_drop, seed, seed_dev, site);
  p->ctx = nullptr; p->ctxT = nullptr; p->dctx = nullptr; p->dq = nullptr; p->dqT = nullptr; p->dkv_video = nullptr; p->dkv_text = nullptr;
  return LRCE_OK;
}

extern "C" int lrce_xattn_fwd(const float* q, const void* kv_video, const void* kv_text, int ld_kv, int kcol, int R, int S, int seg,
                              int Tv, int Lt, int n_cand, float* P, void* ctx, void* ctxT, int ldT, int rowT0, float p_drop,
                              int site, unsigned long long seed, const unsigned long long* seed_dev, void* stream) {
  int rc = require_sm100();
  if (rc != LRCE_OK) return rc;
  XAttnParams p;
  rc = fill_xattn(&p, q, kv_video, kv_text, ld_kv, kcol, R, S, seg, Tv, Lt, n_cand, P, ldT, rowT0, p_drop, site, seed, seed_dev);
  if (rc != LRCE_OK) return rc;
  LRCE_REQUIRE(ctx, "lrce_xattn_fwd: null output");
  p.ctx = reinterpret_cast<bf16*>(ctx); p.ctxT = reinterpret_cast<bf16*>(ctxT);
  rc = xa_configure();
  if (rc != LRCE_OK) return rc;
  xattn_fwd_kernel<<<R * 12, XA_THREADS, XA_SMEM, reinterpret_cast<cudaStream_t>(stream)>>>(p);
  return check_launch("xattn_fwd_kernel");
}

extern "C" int lrce_xattn_bwd(const float* q, const void* kv_video, const void* kv_text, int ld_kv, int kcol, int R, int S, int seg,
                              int Tv, int Lt, int n_cand, const float* P, const float* dctx, void* dq, void* dqT, int ldT, int rowT0,
                              void* dkv_video, void* dkv_text, float p_drop, int site, unsigned long long seed, const unsigned long long* seed_dev, void* stream) {
  int rc = require_sm100();
  if (rc != LRCE_OK) return rc;
  XAttnParams p;
  rc = fill_xattn(&p, q, kv_video, kv_text, ld_kv, kcol, R, S, seg, Tv, Lt, n_cand, const_cast<float*>(P), ldT, rowT0, p_drop, site, seed, seed_dev);
  if (rc != LRCE_OK) return rc;
  LRCE_REQUIRE(dctx && dq && dkv_video && dkv_text, "lrce_xattn_bwd: null argument");
  p.dctx = dctx; p.dq = reinterpret_cast<bf16*>(dq); p.dqT = reinterpret_cast<bf16*>(dqT);
  p.dkv_video = reinterpret_cast<bf16*>(dkv_video); p.dkv_text = reinterpret_cast<bf16*>(dkv_text);
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  rc = xa_configure();
  if (rc != LRCE_OK) return rc;
  if (n_cand > 1) xattn_bwd_kernel<true><<<R * 12, XA_THREADS, XA_SMEM, s>>>(p);
  else xattn_bwd_kernel<false><<<R * 12, XA_THREADS, XA_SMEM, s>>>(p);
  return check_launch("xattn_bwd_kernel");
}

extern "C" int lrce_rowsum_bf16(const void* src, int ld, int cols, float* out, int rows, int accumulate, void* stream) {
  int rc = require_sm100();
  if (rc != LRCE_OK) return rc;
  LRCE_REQUIRE(src && out && rows > 0 && cols > 0 && ld >= cols, "lrce_rowsum_bf16: bad arguments");
  rowsum_bf16_kernel<<<(rows + 7) / 8, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(reinterpret_cast<const bf16*>(src), ld, cols,
                                                                                       out, rows, accumulate);
  return check_launch("rowsum_bf16_kernel");
}

extern "C" int lrce_colsum(const void* src, int is_bf16, long long rows, int cols, long long ld, float* out, void* stream) {
  int rc = require_sm100();
  if (rc != LRCE_OK) return rc;
  LRCE_REQUIRE(src && out && rows > 0 && cols > 0 && ld >= cols, "lrce_colsum: bad arguments");
  const int rpb = 128;
  dim3 grid((cols + 255) / 256, static_cast<unsigned>((rows + rpb - 1) / rpb));
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (is_bf16) colsum_kernel<bf16><<<grid, 256, 0, s>>>(reinterpret_cast<const bf16*>(src), rows, cols, ld, out, rpb);
  else colsum_kernel<float><<<grid, 256, 0, s>>>(reinterpret_cast<const float*>(src), rows, cols, ld, out, rpb);
  return check_launch("colsum_kernel");
}

extern "C" int lrce_transpose_bf16(const void* src, long long rows, int cols, long long ld_src, void* dst, long long ld_dst,
                                   void* stream) {
  int rc = require_sm100();
  if (rc != LRCE_OK) return rc;
  LRCE_REQUIRE(src && dst && src != dst && rows > 0 && cols > 0 && ld_src >= cols && ld_dst >= rows, "lrce_transpose_bf16: bad arguments");
  dim3 grid((cols + 63) / 64, static_cast<unsigned>((rows + 63) / 64));
  transpose_bf16_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(reinterpret_cast<const bf16*>(src), rows, cols, ld_src,
                                                                                reinterpret_cast<bf16*>(dst), ld_dst);
  return check_launch("transpose_bf16_kernel");
}

extern "C" int lrce_add_bf16(void* dst, const void* a, const void* b, const void* c, long long n_elems, void* stream) {
  int rc = require_sm100();
  if (rc != LRCE_OK) return rc;
  LRCE_REQUIRE(dst && a && b && n_elems > 0 && n_elems % 8 == 0, "lrce_add_bf16: bad arguments");
  add_bf16_kernel<<<static_cast<unsigned>((n_elems / 8 + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<bf16*>(dst), reinterpret_cast<const bf16*>(a), reinterpret_cast<const bf16*>(b),
      reinterpret_cast<const bf16*>(c), n_elems / 8);
  return check_launch("add_bf16_kernel");
}

extern "C" int lrce_posembed_bwd(const float* dy, const void* proj, const void* text, int text_fp32, const float* emb_cls,
                                 const float* emb_pos, const float* emb_len, const float* emb_clip, const float* gamma, float eps,
                                 void* dproj, float* d_cls, float* d_pos, float* d_len, float* d_clip, float* dgamma, float* dbeta,
                                 int B, int S, int T, int P, int is_text, float p_drop, int site, unsigned long long seed,
                                 const unsigned long long* seed_dev, void* stream) {
  int rc = require_sm100();
  if (rc != LRCE_OK) return rc;
  LRCE_REQUIRE(dy && emb_cls && emb_pos && gamma && d_cls && d_pos && dgamma && dbeta && B > 0 && S > 0 && T > 0 && P > 0,
               "lrce_posembed_bwd: bad arguments");
  if (is_text) LRCE_REQUIRE(text && S == 1 && T == 1, "lrce_posembed_bwd: text rows need the text features and S = T = 1");
  else LRCE_REQUIRE(proj && dproj && emb_len && emb_clip && d_len && d_clip, "lrce_posembed_bwd: video rows need proj / dproj / tables");
  PosBwdParams p;
  p.dy = dy; p.proj = reinterpret_cast<const bf16*>(proj);
  p.text_f32 = (is_text && text_fp32) ? reinterpret_cast<const float*>(text) : nullptr;
  p.text_bf16 = (is_text && !text_fp32) ? reinterpret_cast<const bf16*>(text) : nullptr;
  p.emb_cls = emb_cls; p.emb_pos = emb_pos; p.emb_len = emb_len; p.emb_clip = emb_clip; p.gamma = gamma; p.eps = eps;
  p.dproj = reinterpret_cast<bf16*>(dproj); p.d_cls = d_cls; p.d_pos = d_pos; p.d_len = d_len; p.d_clip = d_clip;
  p.dgamma = dgamma; p.dbeta = dbeta; p.B = B; p.S = S; p.T = T; p.P = P; p.is_text = is_text;
  p.drop = make_drop(p_drop, seed, seed_dev, site);
  const long long frames = static_cast<long long>(B) * S * T;
  posembed_bwd_kernel<<<static_cast<unsigned>(frames), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(p);
  return check_launch("posembed_bwd_kernel");
}

extern "C" int lrce_bert_embed_ln(const long long* ids, const long long* type_ids, const float* word, const float* pos,
                                  const float* type, const float* gamma, const float* beta, float eps, float* out_f32, void* out_bf16,
                                  long long n, int L, int vocab, int n_types, int max_pos, void* stream) {
  int rc = require_sm100();
  if (rc != LRCE_OK) return rc;
  LRCE_REQUIRE(ids && word && pos && type && gamma && beta && out_f32 && out_bf16 && n > 0 && L > 0 && n % L == 0 && L <= max_pos &&
                   vocab > 0 && n_types > 0,
               "lrce_bert_embed_ln: bad arguments (n=%lld L=%d max_pos=%d)", n, L, max_pos);
  bert_embed_ln_kernel<<<blocks_for_rows(n), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      ids, type_ids, word, pos, type, gamma, beta, eps, out_f32, reinterpret_cast<bf16*>(out_bf16), n, L, vocab, n_types);
  return check_launch("bert_embed_ln_kernel");
}

extern "C" int lrce_bert_attention(const void* qkv, const long long* mask, void* out, int n_seq, int L, int n_heads, void* stream) {
  int rc = require_sm100();
  if (rc != LRCE_OK) return rc;
  LRCE_REQUIRE(qkv && out && n_seq > 0 && L > 0 && L <= BA_MAXL && n_heads > 0, "lrce_bert_attention: bad arguments (L=%d, limit %d)", L,
               BA_MAXL);
  bert_attention_kernel<<<n_seq * n_heads, 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const bf16*>(qkv), mask, reinterpret_cast<bf16*>(out), L, n_heads);
  return check_launch("bert_attention_kernel");
}

extern "C" int lrce_gemm_skinny_bf16(const void* A, int lda, const void* W, int ldw, int M, int N, int K, const float* bias, float* out,
                                     int ldo, void* stream) {
  int rc = require_sm100();
  if (rc != LRCE_OK) return rc;
  LRCE_REQUIRE(A && W && out && M > 0 && N > 0, "lrce_gemm_skinny_bf16: bad arguments");
  LRCE_REQUIRE(K == 768 || K == 3072, "lrce_gemm_skinny_bf16: K must be 768 or 3072 (got %d)", K);
  LRCE_REQUIRE(lda % 8 == 0 && ldw % 8 == 0 && lda >= K && ldw >= K && ldo >= N && (reinterpret_cast<uintptr_t>(A) & 15) == 0 &&
                   (reinterpret_cast<uintptr_t>(W) & 15) == 0,
               "lrce_gemm_skinny_bf16: operands must be 16-byte aligned with pitches that are multiples of 8 elements");
  const int smem = SK_ROWS * (K + 32) * 2 + SK_WARPS * SK_ROWS * 8 * 4;
  static thread_local uint64_t configured = 0;  // one bit per device
  if (needs_device_setup(&configured)) {
    cudaError_t e = cudaFuncSetAttribute(skinny_gemm_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         SK_ROWS * (768 + 32) * 2 + SK_WARPS * SK_ROWS * 8 * 4);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(skinny_gemm_kernel<12>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               SK_ROWS * (3072 + 32) * 2 + SK_WARPS * SK_ROWS * 8 * 4);
    if (e != cudaSuccess) {
      set_error("cudaFuncSetAttribute(skinny_gemm_kernel): %s", cudaGetErrorString(e));
      return LRCE_ECUDA;
    }
    mark_device_setup(&configured);
  }
  dim3 grid((N + 7) / 8, (M + SK_ROWS - 1) / SK_ROWS);
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const bf16 *a = reinterpret_cast<const bf16*>(A), *w = reinterpret_cast<const bf16*>(W);
  if (K == 768) skinny_gemm_kernel<3><<<grid, SK_THREADS, smem, s>>>(a, lda, w, ldw, bias, out, ldo, M, N);
  else skinny_gemm_kernel<12><<<grid, SK_THREADS, smem, s>>>(a, lda, w, ldw, bias, out, ldo, M, N);
  return check_launch("skinny_gemm_kernel");
}
