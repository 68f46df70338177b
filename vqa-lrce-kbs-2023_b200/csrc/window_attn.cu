// window_attn.cu — (shifted-)window multi-head attention core of Video Swin (video_swin_ori.py:166-186) on tcgen05
// tensor cores, with the cyclic shift / window_partition / window_reverse remap (video_swin_ori.py:262-276) fused into
// its loads and stores.
//
// Input  : qkv  bf16 [n_seg * D*H*W, 3C] in NATURAL token order (the qkv Linear is per token, so it runs before any
//          partition); column layout [q | k | v][head][32] (video_swin_ori.py:165).
// Output : out  bf16 [n_seg * D*H*W, C] in natural token order, heads merged (video_swin_ori.py:186), i.e. exactly
//          roll(window_reverse(attn @ v), +shift) — the proj GEMM + residual then runs with no remap at all.
//
// One persistent CTA per SM walks a contiguous range of (head, segment, window) work units; 14 warps:
//   warps 9,10,13 loaders: one warp each for q, k and v: gather the window's 147 rows (64 B each) from their rolled
//                          source tokens with 16-byte cp.async (4 lanes per row -> full 32-byte sectors) into UMMA
//                          "core matrix" order (8 rows x 16 B contiguous), double-buffered; q/k buffers are recycled
//                          as soon as S has been computed, v buffers after P v; the row index is
//                          window_source_token(), the function lrce_remap_index() exports for the bit-exact test
//   warp  11     MMA     : one thread issues S = q k^T (2 row tiles x [M=128, N=160, K=32]) and O = P v (2 x [M=128,
//                          N=32, K=160], v consumed MN-major exactly as it sits in memory) with tcgen05.mma into TMEM
//   warps 0-8,12 softmax : thread-per-row on the TMEM accumulator (two warps split the 160 columns of each 32-row
//                          quarter): t = s*scale*log2e + bias (dense 147x152 bf16 table of this head, resident in
//                          smem) + shift mask (-100 where region ids differ, video_swin_ori.py:357-358); row max;
//                          p = exp2(t - max) written as bf16 A-operand tiles to smem; 1/rowsum applied to O in the
//                          epilogue, which scatters 32-byte row pieces through the inverse remap.
// S(i+1) and P v(i) run on the tensor core while the softmax warps are still busy with item i's epilogue / item i+1's
// first pass, so the kernel is bound by the softmax ALU work, not by the MMAs. N = 147 is padded to 2 x 128 query rows
// and 160 key columns; padded keys get p = 0, padded rows are never stored.
#include "host_common.h"
#include "lrce_common.cuh"
#include "remap.cuh"

namespace lrce {

constexpr int WA_N = 147;           // tokens per (3,7,7) window
constexpr int WA_KEYS = 160;        // key columns of the S tile (multiple of 16)
constexpr int WA_BIAS_PITCH = 152;  // dense bias row pitch (bf16)
constexpr int WA_THREADS = 14 * 32;

// shared memory map (bytes)
constexpr int WA_Q_BYTES = 256 * 64;  // 2 row tiles x 128 rows x 32 dims
constexpr int WA_K_BYTES = 160 * 64;
constexpr int WA_V_BYTES = 160 * 64;
constexpr int WA_QKV_BYTES = WA_Q_BYTES + WA_K_BYTES + WA_V_BYTES;  // 36864 per buffer
constexpr int WA_P_TILE_BYTES = 128 * WA_KEYS * 2;                  // 40960 per row tile
constexpr int WA_BIAS_BYTES = ((WA_N * WA_BIAS_PITCH * 2 + 127) / 128) * 128;
constexpr int WA_OFF_QKV = 0;
constexpr int WA_OFF_P = 2 * WA_QKV_BYTES;
constexpr int WA_OFF_BIAS = WA_OFF_P + 2 * WA_P_TILE_BYTES;
// token / region tables are a 4-deep ring (slot = item & 3): the q loader refills a slot as soon as S(item - 2) has been
// issued, while the softmax warps of item - 2 are still reading theirs
constexpr int WA_RING = 4;
constexpr int WA_OFF_TOK = WA_OFF_BIAS + WA_BIAS_BYTES;         // int [3 loaders][WA_RING][160]
constexpr int WA_OFF_RID = WA_OFF_TOK + 3 * WA_RING * 160 * 4;  // uint8 [WA_RING][160]
constexpr int WA_OFF_XCHG = WA_OFF_RID + WA_RING * 160;         // float [2 kinds][2 halves][160 rows]
constexpr int WA_OFF_BAR = WA_OFF_XCHG + 2 * 2 * 160 * 4;       // mbarriers + tmem slot
constexpr int WA_SMEM = WA_OFF_BAR + 128 + 128 /*align slack*/;

// TMEM columns
constexpr int WA_TM_S0 = 0, WA_TM_S1 = 160, WA_TM_O0 = 320, WA_TM_O1 = 352, WA_TM_COLS = 512;

// UMMA shared-memory descriptor without swizzle: operands live as 8-row x 16-byte "core matrices" (128 contiguous
// bytes); lbo / sbo are the byte distances between core matrices (K-major: lbo along K, sbo along M/N;
// MN-major: sbo along M/N, lbo along K).
__device__ __forceinline__ uint64_t umma_desc_nosw(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(lbo_bytes >> 4) << 16;
  d |= static_cast<uint64_t>(sbo_bytes >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;  // descriptor version (Blackwell)
  return d;                             // layout type 0 = no swizzle
}

__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void cp_async_16(uint32_t smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_dst), "l"(gsrc) : "memory");
}
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// byte offset of the 16-byte chunk (row, chunk) inside an operand stored as core matrices with `cores_per_group` cores
// along the contiguous (K for q/k/P, dims for v) direction
__device__ __forceinline__ uint32_t core_off(int row, int chunk, int cores_per_group) {
  return static_cast<uint32_t>(((row >> 3) * cores_per_group + chunk) * 128 + (row & 7) * 16);
}

__global__ void __launch_bounds__(WA_THREADS, 1)
window_attention_kernel(const bf16* __restrict__ qkv, bf16* __restrict__ out, const bf16* __restrict__ bias_dense,
                        StageGeom g, int n_seg, int C, int n_heads, float scale_log2e) {
  // NOTE: pointers must stay derived from the __shared__ array itself (no integer round trip), otherwise the compiler
  // falls back to generic LD/ST for every shared-memory access
  extern __shared__ __align__(128) uint8_t smem[];
  const bf16* sBias = reinterpret_cast<const bf16*>(smem + WA_OFF_BIAS);
  int* sTok = reinterpret_cast<int*>(smem + WA_OFF_TOK);
  uint8_t* sRid = smem + WA_OFF_RID;
  float* sX = reinterpret_cast<float*>(smem + WA_OFF_XCHG);
  uint64_t* bar_qk_full = reinterpret_cast<uint64_t*>(smem + WA_OFF_BAR);  // [2]
  uint64_t* bar_qk_empty = bar_qk_full + 2;                                // [2]
  uint64_t* bar_v_full = bar_qk_empty + 2;                                 // [2]
  uint64_t* bar_v_empty = bar_v_full + 2;                                  // [2]
  uint64_t* bar_s_full = bar_v_empty + 2;
  uint64_t* bar_p_full = bar_s_full + 1;
  uint64_t* bar_o_full = bar_p_full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_o_full + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nwin = windows_per_segment(g);
  const int T = g.D * g.H * g.W;
  const int n_items = n_seg * nwin;
  const bool shifted = (g.sd | g.sh | g.sw) != 0;
  // work units are (head, segment, window) triples in head-major order; every CTA takes one contiguous, equally sized
  // range, so a CTA changes head (and reloads the 44 KB bias table) at most ceil(heads / CTAs) + 1 times
  const long long n_units = static_cast<long long>(n_heads) * n_items;
  const int u_lo = static_cast<int>(n_units * blockIdx.x / gridDim.x);
  const int u_hi = static_cast<int>(n_units * (blockIdx.x + 1) / gridDim.x);
  const int n_my = u_hi - u_lo;

  // ---- one-time setup: zero the q/k/v staging (pad rows stay zero forever), load this head's bias, barriers, TMEM
  for (int i = tid; i < 2 * WA_QKV_BYTES / 16; i += WA_THREADS) reinterpret_cast<uint4*>(smem + WA_OFF_QKV)[i] = make_uint4(0, 0, 0, 0);
  for (int i = tid; i < 2 * WA_P_TILE_BYTES / 16; i += WA_THREADS) reinterpret_cast<uint4*>(smem + WA_OFF_P)[i] = make_uint4(0, 0, 0, 0);
  if (warp == 11 && lane == 0) {
    for (int b = 0; b < 2; ++b) {
      mbar_init(&bar_qk_full[b], 2);  // the q loader and the k loader
      mbar_init(&bar_qk_empty[b], 1);
      mbar_init(&bar_v_full[b], 1);
      mbar_init(&bar_v_empty[b], 1);
    }
    mbar_init(bar_s_full, 1);
    mbar_init(bar_p_full, 10);
    mbar_init(bar_o_full, 1);
    fence_barrier_init();
  }
  if (warp == 10) tmem_alloc(tmem_slot, WA_TM_COLS);
  fence_proxy_async_smem();  // the zero fill above must be visible to the tensor core's operand reads
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 9 || warp == 10 || warp == 13) {
    // ===================================================================== loaders: warp 9 -> q, 10 -> k, 13 -> v
    const int part = (warp == 9) ? 0 : (warp == 10 ? 1 : 2);
    int* myTok = sTok + part * WA_RING * 160;
    uint64_t* full = (part == 2) ? bar_v_full : bar_qk_full;
    uint64_t* empty = (part == 2) ? bar_v_empty : bar_qk_empty;
    const uint32_t part_off = (part == 0) ? 0u : (part == 1 ? static_cast<uint32_t>(WA_Q_BYTES) : static_cast<uint32_t>(WA_Q_BYTES + WA_K_BYTES));
    const int sub = lane >> 2, ch = lane & 3;  // 4 lanes fetch the 64 contiguous bytes of one row
    for (int j = 0; j < n_my; ++j) {
      const int buf = j & 1, slot = j & (WA_RING - 1);
      const int unit = u_lo + j;
      const int head = unit / n_items, item = unit - head * n_items;
      const int seg = item / nwin, win = item - seg * nwin;
      mbar_wait(&empty[buf], ((j >> 1) & 1) ^ 1);
      for (int r = lane; r < WA_N; r += 32) {
        myTok[slot * 160 + r] = window_source_token(g, win, r);
        if (part == 0) sRid[slot * 160 + r] = static_cast<uint8_t>(shifted ? shift_region_id(g, win, r) : 0);
      }
      __syncwarp();
      const bf16* base = qkv + static_cast<size_t>(seg) * T * 3 * C + part * C + head * 32 + ch * 8;
      const uint32_t sbuf = smem_u32(smem + WA_OFF_QKV + buf * WA_QKV_BYTES) + part_off;
#pragma unroll 4
      for (int r = sub; r < WA_N; r += 8)
        cp_async_16(sbuf + core_off(r, ch, 4), base + static_cast<size_t>(myTok[slot * 160 + r]) * 3 * C);
      asm volatile("cp.async.wait_all;" ::: "memory");
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&full[buf]);
    }
  } else if (warp == 11) {
    // ===================================================================== MMA issuer
    if (lane == 0 && n_my > 0) {
      constexpr uint32_t idesc_s = umma_idesc_bf16(128, WA_KEYS);               // A, B K-major
      constexpr uint32_t idesc_o = umma_idesc_bf16(128, 32) | (1u << 16);       // B (= v) MN-major
      auto issue_s = [&](int j) {
        const uint32_t b = smem_u32(smem + WA_OFF_QKV + (j & 1) * WA_QKV_BYTES);
        const uint32_t q_addr = b, k_addr = b + WA_Q_BYTES;
#pragma unroll
        for (int tile = 0; tile < 2; ++tile)
#pragma unroll
          for (int kk = 0; kk < 2; ++kk)
            umma_bf16_ss(tmem_base + (tile ? WA_TM_S1 : WA_TM_S0),
                         umma_desc_nosw(q_addr + tile * (16 * 512) + kk * 256, 128, 512),
                         umma_desc_nosw(k_addr + kk * 256, 128, 512), idesc_s, kk);
        umma_commit(bar_s_full);
      };
      mbar_wait(&bar_qk_full[0], 0);
      tcgen05_fence_after();
      issue_s(0);
      umma_commit(&bar_qk_empty[0]);
      for (int j = 0; j < n_my; ++j) {
        mbar_wait(bar_p_full, j & 1);  // softmax(j) done: S buffer free, P(j) in smem, O(j-1) drained
        tcgen05_fence_after();
        if (j + 1 < n_my) {
          mbar_wait(&bar_qk_full[(j + 1) & 1], ((j + 1) >> 1) & 1);
          tcgen05_fence_after();
          issue_s(j + 1);
          umma_commit(&bar_qk_empty[(j + 1) & 1]);
        }
        mbar_wait(&bar_v_full[j & 1], (j >> 1) & 1);
        tcgen05_fence_after();
        const uint32_t v_addr = smem_u32(smem + WA_OFF_QKV + (j & 1) * WA_QKV_BYTES) + WA_Q_BYTES + WA_K_BYTES;
        const uint32_t p_addr = smem_u32(smem + WA_OFF_P);
#pragma unroll
        for (int tile = 0; tile < 2; ++tile)
#pragma unroll
          for (int kk = 0; kk < WA_KEYS / 16; ++kk)
            umma_bf16_ss(tmem_base + (tile ? WA_TM_O1 : WA_TM_O0),
                         umma_desc_nosw(p_addr + tile * WA_P_TILE_BYTES + kk * 256, 128, (WA_KEYS / 8) * 128),
                         umma_desc_nosw(v_addr + kk * 1024, /*lbo: key groups*/ 512, /*sbo: dim groups*/ 128), idesc_o, kk);
        umma_commit(bar_o_full);
        umma_commit(&bar_v_empty[j & 1]);
      }
    }
  } else {
    // ===================================================================== softmax + epilogue (warps 0-8, 12)
    const int tile = (warp >= 8) ? 1 : 0;
    const int q = warp & 3;                              // TMEM lane quarter
    const int half = (warp >= 8) ? (warp == 12) : (warp >> 2);
    const int row = tile * 128 + q * 32 + lane;          // query row inside the window
    const int brow = min(row, WA_N - 1);
    const uint32_t lane_addr = static_cast<uint32_t>(q * 32) << 16;
    const int pair_bar = 1 + tile * 4 + q;               // named barrier shared with the warp owning the other columns
    const float MASK_L2 = -100.0f * 1.4426950408889634f;
    const int col0 = half * 80;
    float inv_prev = 0.f;
    int tok_prev = 0, seg_prev = 0, head_prev = 0, head_loaded = -1;
    const int st = (warp < 9) ? tid : (tid - 96);  // index among the 320 softmax threads (warps 0-8 and 12)

    auto store_o = [&](float inv, int seg, int tok, int head, bool valid) {
      uint32_t acc[16];
      tmem_ld_32x16(tmem_base + lane_addr + (tile ? WA_TM_O1 : WA_TM_O0) + half * 16, acc);
      tmem_ld_wait();
      if (valid) {
        uint4 o0, o1;
        o0.x = pack_bf16x2(__uint_as_float(acc[0]) * inv, __uint_as_float(acc[1]) * inv);
        o0.y = pack_bf16x2(__uint_as_float(acc[2]) * inv, __uint_as_float(acc[3]) * inv);
        o0.z = pack_bf16x2(__uint_as_float(acc[4]) * inv, __uint_as_float(acc[5]) * inv);
        o0.w = pack_bf16x2(__uint_as_float(acc[6]) * inv, __uint_as_float(acc[7]) * inv);
        o1.x = pack_bf16x2(__uint_as_float(acc[8]) * inv, __uint_as_float(acc[9]) * inv);
        o1.y = pack_bf16x2(__uint_as_float(acc[10]) * inv, __uint_as_float(acc[11]) * inv);
        o1.z = pack_bf16x2(__uint_as_float(acc[12]) * inv, __uint_as_float(acc[13]) * inv);
        o1.w = pack_bf16x2(__uint_as_float(acc[14]) * inv, __uint_as_float(acc[15]) * inv);
        uint4* dst = reinterpret_cast<uint4*>(out + (static_cast<size_t>(seg) * T + tok) * C + head * 32 + half * 16);
        dst[0] = o0;
        dst[1] = o1;
      }
    };

    for (int j = 0; j < n_my; ++j) {
      const int slot = j & (WA_RING - 1);
      const int unit = u_lo + j;
      const int head = unit / n_items, item = unit - head * n_items;
      const int seg = item / nwin, win = item - seg * nwin;
      if (head != head_loaded) {  // (re)load this head's dense bias; uniform across the softmax warps
        asm volatile("bar.sync 7, 320;" ::: "memory");
        const uint4* src = reinterpret_cast<const uint4*>(bias_dense + static_cast<size_t>(head) * WA_N * WA_BIAS_PITCH);
        uint4* dst = reinterpret_cast<uint4*>(smem + WA_OFF_BIAS);
        for (int i = st; i < WA_N * WA_BIAS_PITCH / 8; i += 320) dst[i] = __ldg(src + i);
        asm volatile("bar.sync 7, 320;" ::: "memory");
        head_loaded = head;
      }
      bool need_mask = false;
      if (shifted) {
        const int nw = g.W / g.ww, nh = g.H / g.wh;
        need_mask = ((win % nw) == nw - 1 && g.sw) || (((win / nw) % nh) == nh - 1 && g.sh);
      }
      mbar_wait(bar_s_full, j & 1);
      tcgen05_fence_after();
      const int tok = sTok[slot * 160 + brow];
      const int rid = sRid[slot * 160 + brow];
      // ---- pass 1: t = s * scale * log2e + bias (+ mask), running max; 80 columns per thread kept in registers
      float t[80];
      float mx = -INFINITY;
      const uint32_t s_addr = tmem_base + lane_addr + (tile ? WA_TM_S1 : WA_TM_S0) + col0;
#pragma unroll
      for (int c = 0; c < 80; c += 16) {
        uint32_t acc[16];
        tmem_ld_32x16(s_addr + c, acc);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; i += 8) {
          const int col = col0 + c + i;  // 8 consecutive key columns
          if (col < WA_BIAS_PITCH) {     // compile-time after unrolling except for `half`
            const uint4 b4 = *reinterpret_cast<const uint4*>(sBias + brow * WA_BIAS_PITCH + col);
            const float2 b01 = unpack_bf16x2(b4.x), b23 = unpack_bf16x2(b4.y), b45 = unpack_bf16x2(b4.z), b67 = unpack_bf16x2(b4.w);
            const float bb[8] = {b01.x, b01.y, b23.x, b23.y, b45.x, b45.y, b67.x, b67.y};
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              float v = fmaf(__uint_as_float(acc[i + e]), scale_log2e, bb[e]);
              if (need_mask && sRid[slot * 160 + min(col + e, WA_N - 1)] != rid) v += MASK_L2;
              if (col + e >= WA_N) v = -INFINITY;
              t[c + i + e] = v;
              mx = fmaxf(mx, v);
            }
          } else {
#pragma unroll
            for (int e = 0; e < 8; ++e) t[c + i + e] = -INFINITY;
          }
        }
      }
      tcgen05_fence_before();  // all TMEM reads of S(j) are complete (wait::ld above)
      sX[(0 * 2 + half) * 160 + row] = mx;  // rows of tile 1 are stored at 32 + (row - 128)
      asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory");
      mx = fmaxf(mx, sX[(0 * 2 + (half ^ 1)) * 160 + row]);
      // ---- epilogue of the previous item (its P v finished long ago); also frees the P buffer for this item
      if (j > 0) {
        mbar_wait(bar_o_full, (j - 1) & 1);
        tcgen05_fence_after();
        store_o(inv_prev, seg_prev, tok_prev, head_prev, row < WA_N);
        tcgen05_fence_before();
      }
      // ---- pass 2: p = exp2(t - max) -> bf16 A-operand tile, row sum
      float sum = 0.f;
      uint8_t* p_row = smem + WA_OFF_P + tile * WA_P_TILE_BYTES;
#pragma unroll
      for (int c = 0; c < 80; c += 8) {
        float p[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          p[e] = ex2_approx(t[c + e] - mx);
          sum += p[e];
        }
        uint4 u;
        u.x = pack_bf16x2(p[0], p[1]); u.y = pack_bf16x2(p[2], p[3]);
        u.z = pack_bf16x2(p[4], p[5]); u.w = pack_bf16x2(p[6], p[7]);
        *reinterpret_cast<uint4*>(p_row + core_off(q * 32 + lane, (col0 + c) >> 3, WA_KEYS / 8)) = u;
      }
      sX[(1 * 2 + half) * 160 + row] = sum;
      fence_proxy_async_smem();  // P writes -> visible to the tensor core
      asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory");
      sum += sX[(1 * 2 + (half ^ 1)) * 160 + row];
      inv_prev = 1.0f / sum;
      tok_prev = tok;
      seg_prev = seg;
      head_prev = head;
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_p_full);
    }
    if (n_my > 0) {
      mbar_wait(bar_o_full, (n_my - 1) & 1);
      tcgen05_fence_after();
      store_o(inv_prev, seg_prev, tok_prev, head_prev, row < WA_N);
      tcgen05_fence_before();
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 10) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, WA_TM_COLS);
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
  const long long n_units = static_cast<long long>(n_seg) * windows_per_segment(g) * n_heads;
  int grid = sm_count();  // one persistent CTA per SM (it owns all 512 TMEM columns)
  if (grid > n_units) grid = static_cast<int>(n_units);
  const float scale_log2e = 0.17677669529663687f * 1.4426950408889634f;  // 32^-0.5 * log2(e)
  window_attention_kernel<<<grid, WA_THREADS, WA_SMEM, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const bf16*>(qkv), reinterpret_cast<bf16*>(out), reinterpret_cast<const bf16*>(bias_dense), g,
      n_seg, C, n_heads, scale_log2e);
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
