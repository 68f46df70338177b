// encoder_walk.cu — the summarisation token's walk through the recurrent cross-modal encoder as ONE persistent
// cooperative kernel (fusionv3.py:41-51 FusionTransformer.forward + :195 final_fc, nn.TransformerDecoderLayer post-norm).
//
// The walk is S segments x 12 layers of six dependent sub-steps on a (rows <= 160) x 768 state: a chain of mat-vec-like
// products that is bound by weight streaming (14.2 MB of bf16 per layer-step) and by dependency latency, not by FLOPs.
// As ~220 separate launches it cost ~15 us per step even inside a CUDA graph; here every SM stays resident, each phase
// is distributed over all CTAs, and phases are separated by a grid-wide barrier (one atomic + one acquire spin, ~1 us):
//
//   P1  h1pre = x + W_sa x + b_sa        x = tok0 | LN3(prev layer) | LN_f(tok + LN3(layer 12)) computed in the prologue;
//                                        length-1 self-attention == out_proj(v_proj(x)), folded into one matrix at pack time
//   P2  q = W_q LN1(h1pre) + b_q          (1/8 scale folded into W_q); h1 = LN1(h1pre) kept for the residual
//   P3  ctx = softmax(q K^T) V            per (row, head) over [video segment s ; text] from the precomputed K/V GEMM
//   P4  h2pre = h1 + W_o ctx + b_o
//   P5  hdn = gelu(W_1 LN2(h2pre) + b_1)  ; h2 = LN2(h2pre)
//   P6  xpre = h2 + W_2 hdn + b_2         (LN3 is applied by the next P1's prologue)
//   end logits = act(W_fc LN_f(tok + LN3(xpre)) + b_fc)
//
// Linear phases: a CTA owns (32-row tile, 8 output columns) pairs; the 8 warps split K and stream the weight rows with
// 16-byte loads straight into mma.sync m16n8k16 B fragments; the 32 activation rows (after the fused residual / LayerNorm
// prologue) sit in shared memory as bf16, plus an fp32 copy for exact residual adds.
#include <cooperative_groups.h>

#include "encoder_common.cuh"
#include "host_common.h"

namespace lrce {

struct EncLayerW {  // device pointers of one decoder layer; the host passes an array of these in device memory
  const bf16 *sa_w, *q_w, *o_w, *w1, *w2;
  const float *sa_b, *q_b, *o_b, *b1, *b2, *n1g, *n1b, *n2g, *n2b, *n3g, *n3b;
};
static_assert(sizeof(EncLayerW) == 16 * sizeof(void*), "layer table layout is part of the C ABI (16 pointers per layer)");

struct WalkParams {
  const EncLayerW* layers;
  int n_layers;
  const bf16 *kv_video, *kv_text;
  int ld_kv;
  const float *tok0, *f_g, *f_b;
  const bf16* fc_w;
  const float* fc_b;
  int n_out, act;
  float* out;         // [R, n_out]
  float* tokens_tap;  // nullptr or [S, R, 768]: the token after every segment (tests)
  float *tok[2], *xp, *a, *h1, *q, *h2;  // fp32 [R, 768] workspace rows
  bf16 *ctx, *hdn;                       // bf16 [R, 768], [R, 3072]
  unsigned* barrier;                     // zeroed by the host wrapper before every launch
  int R, S, Tv, Lt, n_cand;
  float eps;
  unsigned long long* timing;  // profiling hook (lrce_debug_walk_timing): globaltimer of CTA 0 at every phase boundary
};

constexpr int WK_THREADS = 256;
constexpr int WK_WARPS = 8;
constexpr int WK_ROWS = 32;
constexpr int WK_KMAX = 3072;
constexpr int WK_SX_BYTES = WK_ROWS * (WK_KMAX + 32) * 2;            // bf16 activation rows, pitch K + 32
constexpr int WK_SXF_OFF = WK_ROWS * (ENC_D + 32) * 2;               // fp32 copy (K = 768 phases only) behind the bf16 rows
constexpr int WK_RED_BYTES = WK_WARPS * WK_ROWS * 8 * 4;             // cross-warp K reduction / attention scratch
constexpr int WK_SMEM = WK_SX_BYTES + WK_RED_BYTES;
static_assert(WK_SXF_OFF + WK_ROWS * ENC_D * 4 <= WK_SX_BYTES, "fp32 row copy must fit behind the bf16 rows");
enum { ACT_NONE = 0, ACT_GELU = 1, ACT_RELU = 2 };
enum { PRO_BCAST = 0, PRO_LN = 1, PRO_LN2 = 2, PRO_BF16 = 3 };

struct LinPhase {
  int pro;
  const float* xin;   // PRO_LN / PRO_LN2: rows to normalise; PRO_BCAST: one row broadcast to every row
  const float* xres;  // PRO_LN2: rows added between the two LayerNorms
  const float *g1, *b1, *g2, *b2;
  const bf16* xbf;  // PRO_BF16: rows already in bf16
  int K, N;
  const bf16* W;
  const float* bias;
  int res_mode;  // 0 none, 1 the fp32 prologue rows (smem), 2 global fp32 rows `res`
  const float* res;
  int act;
  float* out_f32;
  bf16* out_bf16;
  int ldo;
  float* side;       // optional copy of the fp32 prologue rows, written column-wise by the tiles that own columns < 768
  float* side_full;  // optional copy of the whole fp32 prologue row tile, written by the tile with n0 == 0 (token tap)
};

__device__ __forceinline__ void stamp(const WalkParams& p, int& slot) {
  if (p.timing != nullptr && blockIdx.x == 1 && threadIdx.x == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    p.timing[slot] = t;
  }
  ++slot;
}

__device__ __forceinline__ void grid_barrier(unsigned* bar, unsigned& target, unsigned n_blocks) {
  __syncthreads();
  if (threadIdx.x == 0) {
    target += n_blocks;
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(bar) : "memory");
    unsigned v;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(bar) : "memory");
    } while (v < target);
  }
  __syncthreads();
}

// Row helpers of the fp32 prologues: one warp owns a 768-wide row, lane holds 24 values as 6 chunks of 4 consecutive
// floats at columns (c * 32 + lane) * 4, so that every 16-byte load / store of a warp covers 512 contiguous bytes.
// Rows produced by other CTAs in the previous phase are read through L2 (ld.global.cg): L1 is not coherent across SMs.
__device__ __forceinline__ void row_add_cg(float (&v)[24], const float* src, int lane) {
#pragma unroll
  for (int c = 0; c < 6; ++c) {
    const float4 a = __ldcg(reinterpret_cast<const float4*>(src + (c * 32 + lane) * 4));
    v[c * 4 + 0] += a.x; v[c * 4 + 1] += a.y; v[c * 4 + 2] += a.z; v[c * 4 + 3] += a.w;
  }
}
__device__ __forceinline__ void row_load_param(float (&g)[24], const float* src, int lane) {
#pragma unroll
  for (int c = 0; c < 6; ++c) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(src + (c * 32 + lane) * 4));
    g[c * 4 + 0] = a.x; g[c * 4 + 1] = a.y; g[c * 4 + 2] = a.z; g[c * 4 + 3] = a.w;
  }
}
__device__ __forceinline__ void row_layernorm(float (&v)[24], const float (&g)[24], const float (&b)[24], float eps) {
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 24; ++i) s += v[i];
  const float mean = warp_sum(s) * (1.0f / ENC_D);
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < 24; ++i) { const float d = v[i] - mean; ss = fmaf(d, d, ss); }
  const float rstd = rsqrtf(warp_sum(ss) * (1.0f / ENC_D) + eps);
#pragma unroll
  for (int i = 0; i < 24; ++i) v[i] = fmaf((v[i] - mean) * rstd, g[i], b[i]);
}
__device__ __forceinline__ void row_store(const float (&v)[24], bf16* xb, float* xf, int lane) {
#pragma unroll
  for (int c = 0; c < 6; ++c) {
    const int col = (c * 32 + lane) * 4;
    uint2 u;
    u.x = pack_bf16x2(v[c * 4 + 0], v[c * 4 + 1]);
    u.y = pack_bf16x2(v[c * 4 + 2], v[c * 4 + 3]);
    *reinterpret_cast<uint2*>(xb + col) = u;
    *reinterpret_cast<float4*>(xf + col) = make_float4(v[c * 4 + 0], v[c * 4 + 1], v[c * 4 + 2], v[c * 4 + 3]);
  }
}

__device__ void linear_prologue(const LinPhase& ph, const WalkParams& p, int r_base, bf16* sX, float* sXf) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int pitch = ph.K + 32;
  if (ph.pro == PRO_BF16) {
    const int chunks_per_row = ph.K / 8;
    for (int c = tid; c < WK_ROWS * chunks_per_row; c += WK_THREADS) {
      const int r = c / chunks_per_row, k = (c - r * chunks_per_row) * 8;
      bf16* dst = sX + static_cast<size_t>(r) * pitch + k;
      if (r_base + r < p.R) {
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)),
                     "l"(ph.xbf + static_cast<size_t>(r_base + r) * ph.K + k) : "memory");
      } else {
        *reinterpret_cast<uint4*>(dst) = make_uint4(0, 0, 0, 0);
      }
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
    return;
  }
  // fp32 prologues (K == 768): the 4 rows of this warp are processed together so their loads and reductions overlap;
  // rows are taken in a CTA-dependent rotation so that the ~100 CTAs reading the same 32 rows do not march through the
  // same L2 lines in lock step
  constexpr int RPW = WK_ROWS / WK_WARPS;
  const int rot = (blockIdx.x * 5) & (WK_ROWS - 1);
  float v[RPW][24];
#pragma unroll
  for (int i = 0; i < RPW; ++i) {
    const int r = (warp + i * WK_WARPS + rot) & (WK_ROWS - 1), row = r_base + r;
#pragma unroll
    for (int j = 0; j < 24; ++j) v[i][j] = 0.f;
    if (row < p.R) row_add_cg(v[i], ph.pro == PRO_BCAST ? ph.xin : ph.xin + static_cast<size_t>(row) * ENC_D, lane);
  }
  if (ph.pro != PRO_BCAST) {
    float g[24], b[24];
    row_load_param(g, ph.g1, lane);
    row_load_param(b, ph.b1, lane);
#pragma unroll
    for (int i = 0; i < RPW; ++i) row_layernorm(v[i], g, b, p.eps);
    if (ph.pro == PRO_LN2) {  // LN_outer(xres + LN_inner(xin))
#pragma unroll
      for (int i = 0; i < RPW; ++i) {
        const int row = r_base + ((warp + i * WK_WARPS + rot) & (WK_ROWS - 1));
        if (row < p.R) row_add_cg(v[i], ph.xres + static_cast<size_t>(row) * ENC_D, lane);
      }
      row_load_param(g, ph.g2, lane);
      row_load_param(b, ph.b2, lane);
#pragma unroll
      for (int i = 0; i < RPW; ++i) row_layernorm(v[i], g, b, p.eps);
    }
  }
#pragma unroll
  for (int i = 0; i < RPW; ++i) {
    const int r = (warp + i * WK_WARPS + rot) & (WK_ROWS - 1);
    if (r_base + r >= p.R) {
#pragma unroll
      for (int j = 0; j < 24; ++j) v[i][j] = 0.f;  // rows beyond the batch: zero operand rows, never stored
    }
    row_store(v[i], sX + static_cast<size_t>(r) * pitch, sXf + static_cast<size_t>(r) * ENC_D, lane);
  }
}

__device__ __forceinline__ void prefetch_l2(const void* ptr) { asm volatile("prefetch.global.L2 [%0];" ::"l"(ptr)); }

// Pull the weight rows this CTA will stream in a LATER linear phase (same tile assignment as run_linear) into L2 now, so
// that the phase itself only pays L2 latency: weights never depend on the token state.
__device__ void prefetch_phase_weights(const bf16* W, int K, int N, const WalkParams& p) {
  const int n_tiles_n = (N + 7) / 8;
  const int row_tiles = (p.R + WK_ROWS - 1) / WK_ROWS;
  const long long total = static_cast<long long>(n_tiles_n) * row_tiles;
  const int lo = static_cast<int>(total * blockIdx.x / gridDim.x);
  const int hi = min(static_cast<int>(total * (blockIdx.x + 1) / gridDim.x), lo + n_tiles_n);  // distinct n-tiles only
  const int lines_per_row = K / 64;  // 128-byte lines
  for (int idx = threadIdx.x; idx < (hi - lo) * 8 * lines_per_row; idx += WK_THREADS) {
    const int t = lo + idx / (8 * lines_per_row);
    const int row = (t % n_tiles_n) * 8 + (idx / lines_per_row) % 8;
    prefetch_l2(W + static_cast<size_t>(row) * K + (idx % lines_per_row) * 64);
  }
}

// Same for the K / V head-rows of this CTA's (row, head) units of the coming attention phase (first round).
template <int WPU>
__device__ void prefetch_attention_kv(const WalkParams& p, int seg, int layer) {
  constexpr int UNITS = WK_WARPS / WPU;
  const int n_keys = p.Tv + p.Lt;
  const int total_units = p.R * 12;
  for (int ul = 0; ul < UNITS; ++ul) {
    const int u = blockIdx.x * UNITS + ul;
    if (u >= total_units) break;
    const int b = u / 12, head = u % 12;
    const bf16* vid = p.kv_video + (static_cast<size_t>(b / p.n_cand) * p.S + seg) * p.Tv * p.ld_kv;
    const bf16* txt = p.kv_text + static_cast<size_t>(b) * p.Lt * p.ld_kv;
    for (int idx = threadIdx.x; idx < n_keys * 2; idx += WK_THREADS) {
      const int j = idx >> 1, part = idx & 1;
      const size_t col = static_cast<size_t>(layer) * 2 * ENC_D + head * 64 + part * ENC_D;
      prefetch_l2((j < p.Tv ? vid + static_cast<size_t>(j) * p.ld_kv : txt + static_cast<size_t>(j - p.Tv) * p.ld_kv) + col);
    }
  }
}

// NCH = 32-wide k chunks per warp: 3 (K = 768) or 12 (K = 3072)
template <int NCH>
__device__ void run_linear_t(const LinPhase& ph, const WalkParams& p, uint8_t* smem, int& ts) {
  bf16* sX = reinterpret_cast<bf16*>(smem);
  float* sXf = reinterpret_cast<float*>(smem + WK_SXF_OFF);
  float* sRed = reinterpret_cast<float*>(smem + WK_SX_BYTES);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int pitch = ph.K + 32;  // bf16 elements; (pitch/2) % 32 == 16 words -> conflict-free 16-byte fragment loads
  const int n_tiles_n = (ph.N + 7) / 8;
  const int row_tiles = (p.R + WK_ROWS - 1) / WK_ROWS;
  const long long total = static_cast<long long>(n_tiles_n) * row_tiles;
  const int lo = static_cast<int>(total * blockIdx.x / gridDim.x);
  const int hi = static_cast<int>(total * (blockIdx.x + 1) / gridDim.x);
  int cur_rt = -1;
  const int g = lane >> 2, t4 = lane & 3;
  const int k_begin = warp * (NCH * 32);
  // this warp's K slice of the tile's 8 weight rows: thread (g, t4) holds 16 B of row n0+g per 32-wide k chunk. The
  // weights of a tile are fetched before the prologue / before the previous tile's reduction: they never depend on it.
  uint4 wreg[NCH];
  auto load_w = [&](int t) {
    const int nt = t % n_tiles_n;
    const bf16* wrow = ph.W + static_cast<size_t>(nt * 8 + g) * ph.K + k_begin + 8 * t4;
#pragma unroll
    for (int c = 0; c < NCH; ++c) wreg[c] = __ldg(reinterpret_cast<const uint4*>(wrow + c * 32));
  };
  if (lo < hi) load_w(lo);
  for (int t = lo; t < hi; ++t) {
    const int rt = t / n_tiles_n, nt = t - rt * n_tiles_n;
    const int r_base = rt * WK_ROWS, n0 = nt * 8;
    if (rt != cur_rt) {
      __syncthreads();  // every warp is done with the previous row tile's activations
      linear_prologue(ph, p, r_base, sX, sXf);
      __syncthreads();
      if (cur_rt < 0) stamp(p, ts);
      cur_rt = rt;
    }
    const bf16* xa0 = sX + static_cast<size_t>(g) * pitch + k_begin + 8 * t4;
    float acc[2][4];
#pragma unroll
    for (int m = 0; m < 2; ++m) acc[m][0] = acc[m][1] = acc[m][2] = acc[m][3] = 0.f;
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
#pragma unroll
      for (int m = 0; m < 2; ++m) {
        const uint4 xlo = *reinterpret_cast<const uint4*>(xa0 + static_cast<size_t>(m * 16) * pitch + c * 32);
        const uint4 xhi = *reinterpret_cast<const uint4*>(xa0 + static_cast<size_t>(m * 16 + 8) * pitch + c * 32);
        mma16816(acc[m], xlo.x, xhi.x, xlo.y, xhi.y, wreg[c].x, wreg[c].y);
        mma16816(acc[m], xlo.z, xhi.z, xlo.w, xhi.w, wreg[c].z, wreg[c].w);
      }
    }
    if (t + 1 < hi) load_w(t + 1);
    // ---- cross-warp K reduction, bias, residual, activation, store
#pragma unroll
    for (int m = 0; m < 2; ++m) {
      float* r = sRed + (warp * WK_ROWS + m * 16 + g) * 8 + 2 * t4;
      r[0] = acc[m][0]; r[1] = acc[m][1];
      r[8 * 8] = acc[m][2]; r[8 * 8 + 1] = acc[m][3];
    }
    __syncthreads();
    {
      const int r = tid >> 3, col = tid & 7;
      float v = 0.f;
#pragma unroll
      for (int w = 0; w < WK_WARPS; ++w) v += sRed[(w * WK_ROWS + r) * 8 + col];
      const int row = r_base + r, n = n0 + col;
      if (row < p.R && n < ph.N) {
        if (ph.bias) v += __ldg(ph.bias + n);
        if (ph.res_mode == 1) v += sXf[r * ENC_D + n];
        else if (ph.res_mode == 2) v += __ldcg(ph.res + static_cast<size_t>(row) * ENC_D + n);
        if (ph.act == ACT_GELU) v = gelu_erf(v);
        else if (ph.act == ACT_RELU) v = fmaxf(v, 0.f);
        if (ph.out_bf16) ph.out_bf16[static_cast<size_t>(row) * ph.ldo + n] = __float2bfloat16(v);
        else ph.out_f32[static_cast<size_t>(row) * ph.ldo + n] = v;
        if (ph.side && n < ENC_D) ph.side[static_cast<size_t>(row) * ENC_D + n] = sXf[r * ENC_D + n];
      }
    }
    if (ph.side_full && nt == 0) {
      for (int i = tid; i < WK_ROWS * ENC_D / 4; i += WK_THREADS) {
        const int r = i / (ENC_D / 4);
        if (r_base + r < p.R)
          reinterpret_cast<float4*>(ph.side_full + static_cast<size_t>(r_base) * ENC_D)[i] = reinterpret_cast<const float4*>(sXf)[i];
      }
    }
    __syncthreads();  // sRed is reused by the next tile
  }
  if (cur_rt < 0) stamp(p, ts);  // no tile for this CTA: keep the stamp layout fixed
}

__device__ void run_linear(const LinPhase& ph, const WalkParams& p, uint8_t* smem, int& ts) {
  if (ph.K == ENC_D) run_linear_t<ENC_D / (WK_WARPS * 32)>(ph, p, smem, ts);
  else run_linear_t<4 * ENC_D / (WK_WARPS * 32)>(ph, p, smem, ts);
}

// ctx[row, head*64 ..] = softmax(q . K^T) V for (row, head) units; WPU warps per unit, 8 / WPU units per CTA and round.
// The unit's K and V head-rows (n_keys x 128 B each) are staged in shared memory with two waves of cp.async (K, then V:
// the scores are computed while V is still landing). 8 lanes cover one 128-byte row; a warp handles 4 keys per step and
// keeps its scores in registers (fully unrolled over the 256-key limit).
template <int WPU>
__device__ void run_attention(const WalkParams& p, int seg, int layer, uint8_t* smem, int& ts) {
  bool stamped = false;
  constexpr int UNITS = WK_WARPS / WPU;
  const int n_keys = p.Tv + p.Lt;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int slot = warp / WPU, wi = warp % WPU;
  const int gtid = wi * 32 + lane;  // thread index inside the unit group
  uint8_t* stage = smem + static_cast<size_t>(slot) * n_keys * 256;  // [K rows | V rows]
  float* scratch = reinterpret_cast<float*>(smem + WK_SX_BYTES);  // per slot: [264] scores, [WPU] max, [WPU] sum, [WPU][64] out
  float* sP = scratch + slot * (264 + 16 + WPU * 64);
  float* sMax = sP + 264;
  float* sSum = sMax + 8;
  float* sOut = sSum + 8;
  const int total_units = p.R * 12;
  const int l8 = lane & 7, kslot = lane >> 3;
  const int per_warp = (((n_keys + WPU - 1) / WPU) + 3) & ~3;
  const int k_lo = wi * per_warp, k_hi = min(n_keys, k_lo + per_warp);
  for (int base = blockIdx.x * UNITS; base < total_units; base += gridDim.x * UNITS) {
    const int u = base + slot;
    const bool active = u < total_units;
    const int b = active ? u / 12 : 0, head = active ? u % 12 : 0;
    {
      const size_t col_k = static_cast<size_t>(layer) * 2 * ENC_D + head * 64;
      const bf16* vid = p.kv_video + (static_cast<size_t>(b / p.n_cand) * p.S + seg) * p.Tv * p.ld_kv + col_k;
      const bf16* txt = p.kv_text + static_cast<size_t>(b) * p.Lt * p.ld_kv + col_k;
#pragma unroll
      for (int part = 0; part < 2; ++part) {  // 0: K, 1: V
        if (active) {
          for (int c = gtid; c < n_keys * 8; c += WPU * 32) {
            const int j = c >> 3, ch = c & 7;
            const bf16* src = (j < p.Tv ? vid + static_cast<size_t>(j) * p.ld_kv : txt + static_cast<size_t>(j - p.Tv) * p.ld_kv) +
                              part * ENC_D + ch * 8;
            const uint32_t dst = smem_u32(stage) + (part * n_keys + j) * 128 + ch * 16;
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
          }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
      }
    }
    float qv[8];
    {
      const float* qp = p.q + static_cast<size_t>(b) * ENC_D + head * 64 + l8 * 8;
      const float4 a = __ldcg(reinterpret_cast<const float4*>(qp)), c = __ldcg(reinterpret_cast<const float4*>(qp + 4));
      qv[0] = a.x; qv[1] = a.y; qv[2] = a.z; qv[3] = a.w; qv[4] = c.x; qv[5] = c.y; qv[6] = c.z; qv[7] = c.w;
    }
    asm volatile("cp.async.wait_group 1;" ::: "memory");
    asm volatile("bar.sync %0, %1;" ::"r"(1 + slot), "r"(WPU * 32) : "memory");
    if (!stamped) { stamp(p, ts); stamped = true; }
    // ---- scores of this warp's keys (every lane of an octet ends up with the score of the octet's key); they are parked
    // in shared memory, each octet reads back only what it wrote
    float mx = -INFINITY;
#pragma unroll 4
    for (int j0 = k_lo; j0 < k_hi; j0 += 4) {
      const int j = j0 + kslot;
      const int jr = min(j, n_keys - 1);
      const uint4 kk = *reinterpret_cast<const uint4*>(stage + static_cast<size_t>(jr) * 128 + l8 * 16);
      const float2 f0 = unpack_bf16x2(kk.x), f1 = unpack_bf16x2(kk.y), f2 = unpack_bf16x2(kk.z), f3 = unpack_bf16x2(kk.w);
      float d = (qv[0] * f0.x + qv[1] * f0.y) + (qv[2] * f1.x + qv[3] * f1.y) + (qv[4] * f2.x + qv[5] * f2.y) +
                (qv[6] * f3.x + qv[7] * f3.y);
      d += __shfl_xor_sync(0xffffffffu, d, 1);
      d += __shfl_xor_sync(0xffffffffu, d, 2);
      d += __shfl_xor_sync(0xffffffffu, d, 4);
      if (j >= k_hi) d = -INFINITY;
      if (l8 == 0) sP[j0 + kslot] = d;
      mx = fmaxf(mx, d);
    }
    mx = warp_max(mx);
    if (lane == 0) sMax[wi] = mx;
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    asm volatile("bar.sync %0, %1;" ::"r"(1 + slot), "r"(WPU * 32) : "memory");  // max exchange + V landed + scores visible
#pragma unroll
    for (int w = 0; w < WPU; ++w) mx = fmaxf(mx, sMax[w]);
    // ---- p = exp(s - max), partial sums and partial P V over this warp's keys
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    float sum = 0.f;
    const uint8_t* vstage = stage + static_cast<size_t>(n_keys) * 128;
#pragma unroll 4
    for (int j0 = k_lo; j0 < k_hi; j0 += 4) {
      const int jr = min(j0 + kslot, n_keys - 1);
      const float pj = __expf(sP[j0 + kslot] - mx);  // exactly 0 for the keys beyond this warp's range (score -inf)
      sum += pj;
      const uint4 vv = *reinterpret_cast<const uint4*>(vstage + static_cast<size_t>(jr) * 128 + l8 * 16);
      float2 f;
      f = unpack_bf16x2(vv.x); acc[0] = fmaf(pj, f.x, acc[0]); acc[1] = fmaf(pj, f.y, acc[1]);
      f = unpack_bf16x2(vv.y); acc[2] = fmaf(pj, f.x, acc[2]); acc[3] = fmaf(pj, f.y, acc[3]);
      f = unpack_bf16x2(vv.z); acc[4] = fmaf(pj, f.x, acc[4]); acc[5] = fmaf(pj, f.y, acc[5]);
      f = unpack_bf16x2(vv.w); acc[6] = fmaf(pj, f.x, acc[6]); acc[7] = fmaf(pj, f.y, acc[7]);
    }
    sum = warp_sum(sum) * 0.125f;  // every key was counted by the 8 lanes of its octet
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], 8);
      acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], 16);
    }
    if (lane == 0) sSum[wi] = sum;
    if (lane < 8) {
#pragma unroll
      for (int i = 0; i < 8; ++i) sOut[wi * 64 + lane * 8 + i] = acc[i];
    }
    asm volatile("bar.sync %0, %1;" ::"r"(1 + slot), "r"(WPU * 32) : "memory");
    if (active && gtid < 64) {
      float tot = 0.f, o = 0.f;
#pragma unroll
      for (int w = 0; w < WPU; ++w) { tot += sSum[w]; o += sOut[w * 64 + gtid]; }
      p.ctx[static_cast<size_t>(b) * ENC_D + head * 64 + gtid] = __float2bfloat16(o / tot);
    }
    __syncthreads();  // staging and scratch are reused by the next round
  }
  if (!stamped) stamp(p, ts);
}

__global__ void __launch_bounds__(WK_THREADS, 1) encoder_walk_kernel(const WalkParams p) {
  extern __shared__ __align__(16) uint8_t wk_smem[];
  unsigned target = 0;
  const unsigned n_blocks = gridDim.x;
  int ts = 0;
  stamp(p, ts);
  const bool wide_units = (p.Tv + p.Lt) * 256 * 4 <= WK_SX_BYTES;  // four (row, head) units fit the staging area
  for (int s = 0; s < p.S; ++s) {
    for (int n = 0; n < p.n_layers; ++n) {
      const EncLayerW L = p.layers[n];
      LinPhase ph;
      // ---- P1: h1pre = x + W_sa x + b_sa
      ph = LinPhase();
      if (n == 0 && s == 0) {
        ph.pro = PRO_BCAST; ph.xin = p.tok0;
      } else if (n == 0) {
        const EncLayerW Lp = p.layers[p.n_layers - 1];
        ph.pro = PRO_LN2; ph.xin = p.xp; ph.g1 = Lp.n3g; ph.b1 = Lp.n3b; ph.xres = p.tok[(s - 1) & 1]; ph.g2 = p.f_g; ph.b2 = p.f_b;
        if (p.tokens_tap) ph.side_full = p.tokens_tap + static_cast<size_t>(s - 1) * p.R * ENC_D;
      } else {
        const EncLayerW Lp = p.layers[n - 1];
        ph.pro = PRO_LN; ph.xin = p.xp; ph.g1 = Lp.n3g; ph.b1 = Lp.n3b;
      }
      if (n == 0) ph.side = p.tok[s & 1];
      ph.K = ENC_D; ph.N = ENC_D; ph.W = L.sa_w; ph.bias = L.sa_b; ph.res_mode = 1; ph.out_f32 = p.a; ph.ldo = ENC_D;
      prefetch_phase_weights(L.q_w, ENC_D, ENC_D, p);
      if (wide_units) prefetch_attention_kv<2>(p, s, n);
      else prefetch_attention_kv<4>(p, s, n);
      run_linear(ph, p, wk_smem, ts);
      stamp(p, ts);
      grid_barrier(p.barrier, target, n_blocks);
      stamp(p, ts);
      // ---- P2: q = W_q LN1(h1pre) + b_q ; h1 = LN1(h1pre)
      ph = LinPhase();
      ph.pro = PRO_LN; ph.xin = p.a; ph.g1 = L.n1g; ph.b1 = L.n1b;
      ph.K = ENC_D; ph.N = ENC_D; ph.W = L.q_w; ph.bias = L.q_b; ph.out_f32 = p.q; ph.ldo = ENC_D; ph.side = p.h1;
      prefetch_phase_weights(L.o_w, ENC_D, ENC_D, p);
      run_linear(ph, p, wk_smem, ts);
      stamp(p, ts);
      grid_barrier(p.barrier, target, n_blocks);
      stamp(p, ts);
      // ---- P3: cross attention over [video segment s ; text]
      prefetch_phase_weights(L.w1, ENC_D, 4 * ENC_D, p);
      if (wide_units) run_attention<2>(p, s, n, wk_smem, ts);
      else run_attention<4>(p, s, n, wk_smem, ts);
      stamp(p, ts);
      grid_barrier(p.barrier, target, n_blocks);
      stamp(p, ts);
      // ---- P4: h2pre = h1 + W_o ctx + b_o
      ph = LinPhase();
      ph.pro = PRO_BF16; ph.xbf = p.ctx;
      ph.K = ENC_D; ph.N = ENC_D; ph.W = L.o_w; ph.bias = L.o_b; ph.res_mode = 2; ph.res = p.h1; ph.out_f32 = p.a; ph.ldo = ENC_D;
      prefetch_phase_weights(L.w2, 4 * ENC_D, ENC_D, p);
      run_linear(ph, p, wk_smem, ts);
      stamp(p, ts);
      grid_barrier(p.barrier, target, n_blocks);
      stamp(p, ts);
      // ---- P5: hdn = gelu(W_1 LN2(h2pre) + b_1) ; h2 = LN2(h2pre)
      ph = LinPhase();
      ph.pro = PRO_LN; ph.xin = p.a; ph.g1 = L.n2g; ph.b1 = L.n2b;
      ph.K = ENC_D; ph.N = 4 * ENC_D; ph.W = L.w1; ph.bias = L.b1; ph.act = ACT_GELU; ph.out_bf16 = p.hdn; ph.ldo = 4 * ENC_D;
      ph.side = p.h2;
      {
        const bool last = (n + 1 == p.n_layers);
        if (last && s + 1 == p.S) prefetch_phase_weights(p.fc_w, ENC_D, p.n_out, p);
        else prefetch_phase_weights(p.layers[last ? 0 : n + 1].sa_w, ENC_D, ENC_D, p);
      }
      run_linear(ph, p, wk_smem, ts);
      stamp(p, ts);
      grid_barrier(p.barrier, target, n_blocks);
      stamp(p, ts);
      // ---- P6: xpre = h2 + W_2 hdn + b_2
      ph = LinPhase();
      ph.pro = PRO_BF16; ph.xbf = p.hdn;
      ph.K = 4 * ENC_D; ph.N = ENC_D; ph.W = L.w2; ph.bias = L.b2; ph.res_mode = 2; ph.res = p.h2; ph.out_f32 = p.xp; ph.ldo = ENC_D;
      run_linear(ph, p, wk_smem, ts);
      stamp(p, ts);
      grid_barrier(p.barrier, target, n_blocks);
      stamp(p, ts);
    }
  }
  // ---- answer head on the final token LN_f(tok + LN3(xpre))
  {
    const EncLayerW Lp = p.layers[p.n_layers - 1];
    LinPhase ph = LinPhase();
    ph.pro = PRO_LN2; ph.xin = p.xp; ph.g1 = Lp.n3g; ph.b1 = Lp.n3b; ph.xres = p.tok[(p.S - 1) & 1]; ph.g2 = p.f_g; ph.b2 = p.f_b;
    if (p.tokens_tap) ph.side_full = p.tokens_tap + static_cast<size_t>(p.S - 1) * p.R * ENC_D;
    ph.K = ENC_D; ph.N = p.n_out; ph.W = p.fc_w; ph.bias = p.fc_b; ph.act = p.act; ph.out_f32 = p.out; ph.ldo = p.n_out;
    run_linear(ph, p, wk_smem, ts);
    stamp(p, ts);
  }
}

static unsigned long long* g_walk_timing = nullptr;

}  // namespace lrce

using namespace lrce;

// Profiling hook, not part of the product path: when `buf` (device, >= 3 + 18 * S * n_layers entries) is non-NULL, the last
// CTA of every following lrce_encoder_walk launch records %globaltimer three times per phase: after its first prologue
// (after the K/V staging for the attention phase), before the grid barrier and after it.
extern "C" int lrce_debug_walk_timing(unsigned long long* buf) {
  g_walk_timing = buf;
  return LRCE_OK;
}

extern "C" size_t lrce_encoder_walk_workspace_bytes(int rows) {
  // 7 fp32 [rows, 768] buffers, bf16 [rows, 768] + [rows, 3072], one 256-byte slot for the grid barrier
  return static_cast<size_t>(rows) * ENC_D * 4 * 7 + static_cast<size_t>(rows) * ENC_D * 2 * 5 + 256;
}

extern "C" int lrce_encoder_walk(const void* layer_table, int n_layers, const void* kv_video, const void* kv_text, int ld_kv,
                                 const float* tok0, const float* f_gamma, const float* f_beta, float eps, const void* fc_w,
                                 const float* fc_b, int n_out, int act, float* out, float* tokens_tap, void* workspace,
                                 int rows, int S, int Tv, int Lt, int n_cand, void* stream) {
  int rc = require_sm100();
  if (rc != LRCE_OK) return rc;
  LRCE_REQUIRE(layer_table && kv_video && kv_text && tok0 && f_gamma && f_beta && fc_w && fc_b && out && workspace,
               "lrce_encoder_walk: null argument");
  LRCE_REQUIRE(n_layers > 0 && rows > 0 && S > 0 && Tv > 0 && Lt > 0 && n_cand > 0 && n_out > 0 && rows % n_cand == 0,
               "lrce_encoder_walk: bad sizes (layers=%d rows=%d S=%d Tv=%d Lt=%d cand=%d out=%d)", n_layers, rows, S, Tv, Lt,
               n_cand, n_out);
  LRCE_REQUIRE(Tv + Lt <= 256, "lrce_encoder_walk: %d memory tokens exceed the 256-key limit", Tv + Lt);
  LRCE_REQUIRE(ld_kv % 8 == 0 && ld_kv >= n_layers * 2 * ENC_D, "lrce_encoder_walk: K/V row pitch %d too small / unaligned", ld_kv);
  LRCE_REQUIRE(act >= 0 && act <= 2, "lrce_encoder_walk: unknown activation %d", act);
  LRCE_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "lrce_encoder_walk: workspace must be 256-byte aligned");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  static thread_local uint64_t configured = 0;  // one bit per device
  if (needs_device_setup(&configured)) {
    cudaError_t e = cudaFuncSetAttribute(encoder_walk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, WK_SMEM);
    int per_sm = 0;
    if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, encoder_walk_kernel, WK_THREADS, WK_SMEM);
    if (e != cudaSuccess || per_sm < 1) {
      set_error("encoder_walk_kernel cannot be made resident (smem=%d): %s", WK_SMEM, cudaGetErrorString(e));
      return LRCE_ECUDA;
    }
    mark_device_setup(&configured);
  }
  const int max_grid = sm_count();  // one CTA per SM: every CTA of the cooperative grid is co-resident
  WalkParams p;
  p.layers = reinterpret_cast<const EncLayerW*>(layer_table);
  p.n_layers = n_layers;
  p.kv_video = reinterpret_cast<const bf16*>(kv_video);
  p.kv_text = reinterpret_cast<const bf16*>(kv_text);
  p.ld_kv = ld_kv;
  p.tok0 = tok0; p.f_g = f_gamma; p.f_b = f_beta;
  p.fc_w = reinterpret_cast<const bf16*>(fc_w);
  p.fc_b = fc_b;
  p.n_out = n_out; p.act = act;
  p.out = out; p.tokens_tap = tokens_tap;
  uint8_t* ws = reinterpret_cast<uint8_t*>(workspace);
  p.barrier = reinterpret_cast<unsigned*>(ws);
  float* f = reinterpret_cast<float*>(ws + 256);
  const size_t rowsz = static_cast<size_t>(rows) * ENC_D;
  p.tok[0] = f; p.tok[1] = f + rowsz; p.xp = f + 2 * rowsz; p.a = f + 3 * rowsz; p.h1 = f + 4 * rowsz; p.q = f + 5 * rowsz;
  p.h2 = f + 6 * rowsz;
  bf16* h = reinterpret_cast<bf16*>(f + 7 * rowsz);
  p.ctx = h; p.hdn = h + rowsz;
  p.R = rows; p.S = S; p.Tv = Tv; p.Lt = Lt; p.n_cand = n_cand; p.eps = eps;
  p.timing = g_walk_timing;
  cudaError_t e = cudaMemsetAsync(p.barrier, 0, 256, s);
  if (e != cudaSuccess) {
    set_error("lrce_encoder_walk: cudaMemsetAsync: %s", cudaGetErrorString(e));
    return LRCE_ECUDA;
  }
  void* args[] = {&p};
  e = cudaLaunchCooperativeKernel(reinterpret_cast<void*>(encoder_walk_kernel), dim3(max_grid), dim3(WK_THREADS), args, WK_SMEM, s);
  if (e != cudaSuccess) {
    set_error("cudaLaunchCooperativeKernel(encoder_walk_kernel): %s", cudaGetErrorString(e));
    return LRCE_ECUDA;
  }
  return check_launch("encoder_walk_kernel");
}
