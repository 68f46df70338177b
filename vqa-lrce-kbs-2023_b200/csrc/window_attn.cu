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
//   warp  10     loader   : gathers the window's 147 q, k and v rows (64 B each) from their rolled source tokens with
//                           16-byte cp.async into UMMA "core matrix" order (8 rows x 16 B contiguous), double-buffered;
//                           the row index is window_source_token_377(), the function lrce_remap_index() exports for the
//                           bit-exact test. v rows carry a 5th 16-byte chunk holding a constant 1 so that the P v product
//                           also yields the softmax row sums.
//   warps 11,14  MMA      : one thread each (issuing 24 tcgen05.mma per item from ONE thread took ~2600 cycles and delayed
//                           the P v results; warp 11 issues S and the row-tile-0 products, warp 14 the row-tile-1 products): S = q k^T (2 row tiles x [M=128, N=160, K=32]) and, per row tile, TWO
//                           products O_A = P[:, 0:80] v[0:80], O_B = P[:, 80:160] v[80:160] ([M=128, N=48, K=80], v consumed
//                           MN-major exactly as it sits in memory) with tcgen05.mma into TMEM
//   warps 0-7    softmax  : rows 0..127 of the window (thread = row, warp w and w+4 split the 160 key columns in halves)
//   warps 8,12 / 9,13     : rows 128..146 of even / odd work units (placed in TMEM lane quarter 0 / 1 so that the extra
//                           load is spread over two warp schedulers)
// Softmax on the TMEM accumulator: t = s*scale*log2e + bias (dense 147x152 bf16 table of this head, resident in smem)
// + shift mask (-100 where region ids differ, video_swin_ori.py:357-358; keys are staged grouped by their mask class —
// remap.cuh key_slot_377 — so the mask is one additive constant per 8-column group of the tile);
// p = exp2(t - m_half) with the maximum of the thread's OWN 80 columns — the two halves never synchronise: their
// products are kept in separate accumulators and merged in the epilogue, out = (a O_A + b O_B) / (a l_A + b l_B),
// a = 2^(m_A - m), b = 2^(m_B - m), l = the ones-column sums. The S accumulator is released as soon as a warp has loaded
// its 80 scores into registers, so S(i+1) runs on the tensor core under the softmax of item i. All twelve softmax warps
// execute ONE copy of the (fully unrolled) code: four template instances thrashed the instruction cache.
#include "host_common.h"
#include "lrce_common.cuh"
#include "remap.cuh"

namespace lrce {

constexpr int WA_N = 147;           // tokens per (3,7,7) window
constexpr int WA_KEYS = 160;        // key columns of the S tile (multiple of 16)
constexpr int WA_HALF = 80;         // key columns per softmax thread
constexpr int WA_BIAS_PITCH = 160;  // dense bias row pitch (bf16) = key columns of the score tile
constexpr int WA_VCH = 6;           // 16-byte chunks per staged v row: 4 of data, 1 with the constant 1, 1 of zeros
constexpr int WA_ON = 8 * WA_VCH;   // N of the P v products (48)
constexpr int WA_THREADS = 15 * 32;
constexpr int WA_SOFTMAX_ARRIVALS = 12;  // 8 main warps + 2 leftover warps of the item + 2 leftover warps of the other parity

// shared memory map (bytes)
constexpr int WA_Q_BYTES = 256 * 64;  // 2 row tiles x 128 rows x 32 dims
constexpr int WA_K_BYTES = WA_KEYS * 64;
constexpr int WA_V_BYTES = WA_KEYS * 16 * WA_VCH;
constexpr int WA_QKV_BYTES = WA_Q_BYTES + WA_K_BYTES + WA_V_BYTES;  // 41984 per buffer
constexpr int WA_P_TILE_BYTES = 128 * WA_KEYS * 2;                  // 40960 per row tile
constexpr int WA_BIAS_BYTES = ((WA_N * WA_BIAS_PITCH * 2 + 127) / 128) * 128;
constexpr int WA_OFF_QKV = 0;
constexpr int WA_OFF_P = 2 * WA_QKV_BYTES;
constexpr int WA_OFF_BIAS = WA_OFF_P + 2 * WA_P_TILE_BYTES;
constexpr int WA_OFF_M = WA_OFF_BIAS + WA_BIAS_BYTES;         // float [4 slots][2 halves][160 rows]: per-half row maxima
constexpr int WA_OFF_BAR = WA_OFF_M + 4 * 2 * 160 * 4;        // mbarriers + tmem slot
constexpr int WA_SMEM = WA_OFF_BAR + 256;
static_assert(WA_SMEM <= 227 * 1024, "window attention shared-memory budget");

// TMEM columns: S tiles, then per row tile the two partial outputs
constexpr int WA_TM_S0 = 0, WA_TM_S1 = 160, WA_TM_O = 320, WA_TM_COLS = 512;
static_assert(WA_TM_O + 4 * WA_ON <= WA_TM_COLS, "TMEM budget");

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
__device__ __forceinline__ uint32_t tmem_ld_32x1(uint32_t taddr) {
  uint32_t v;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(v) : "r"(taddr) : "memory");
  return v;
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

struct WaShared {
  uint8_t* smem;
  uint64_t *qk_full, *qk_empty, *v_full, *v_empty, *s_full, *s_free, *p_full, *o_full;
  long long* prof;  // profiling hook (lrce_debug_attention_timing): [warp][8] cycle counters of CTA 0, or nullptr
};

__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n"
      "selp.u32 %0, 1, 0, P1;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}

// mbarrier wait that also accounts the stall to counter `slot` of this warp when the profiling hook is armed. With the
// hook armed it doubles as a watchdog: a wait longer than ~50 ms records (warp, slot, item) in prof[132..] and raises
// a CTA-wide abort flag that makes every later wait fall through, so a protocol deadlock ends with a report, not a hang.
__device__ __forceinline__ void timed_wait(const WaShared& sh, uint64_t* bar, uint32_t parity, int slot, int item = -1) {
  if (sh.prof == nullptr) {
    mbar_wait(bar, parity);
    return;
  }
  volatile int* abort_flag = reinterpret_cast<volatile int*>(sh.smem + WA_OFF_BAR + 200);
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (*abort_flag) break;
    if (clock64() - t0 > 100000000LL) {
      if ((threadIdx.x & 31) == 0) {
        *abort_flag = 1;
        sh.prof[132 + (threadIdx.x >> 5)] = (static_cast<long long>(blockIdx.x) << 40) | (static_cast<long long>(slot + 1) << 32) |
                                            static_cast<unsigned>(item);
      }
      break;
    }
  }
  if ((threadIdx.x & 31) == 0 && blockIdx.x == 0) sh.prof[(threadIdx.x >> 5) * 8 + slot] += clock64() - t0;
}

// One softmax warp; all twelve run this one function (one copy of the unrolled code in the instruction cache).
// tile 0: rows 0..127 of every item; tile 1: rows 128..146 of the items whose parity is `parity`, placed in TMEM lanes /
// A-operand rows lane_base .. lane_base + 31. `half` selects key columns [80 half, 80 half + 80) of the score tile.
__device__ __forceinline__ void softmax_warp(const WaShared& sh, uint32_t tmem_base, int tile, int half, int lane_base,
                                          int parity, const bf16* __restrict__ bias_dense, bf16* __restrict__ out,
                                          const StageGeom& g, int n_items, int nwin, int T, int C, int u_lo, int n_my,
                                          float scale_log2e, bool shifted) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // the pointer must be derived from the __shared__ array itself, otherwise every access below compiles to generic LD/ST
  extern __shared__ __align__(128) uint8_t smem[];
  float* sM = reinterpret_cast<float*>(smem + WA_OFF_M);
  const int prow = lane_base + lane;            // A-operand row of the P tile == TMEM lane
  const int row = tile ? 128 + lane : prow;     // query row inside the window
  const int brow = min(row, WA_N - 1);
  const uint32_t lane_addr = static_cast<uint32_t>(lane_base) << 16;
  const uint32_t s_addr = tmem_base + lane_addr + (tile ? WA_TM_S1 : WA_TM_S0) + half * WA_HALF;
  const uint32_t o_addr = tmem_base + lane_addr + WA_TM_O + tile * 2 * WA_ON;  // O_A; O_B at + WA_ON
  uint8_t* p_row = smem + WA_OFF_P + tile * WA_P_TILE_BYTES + core_off(prow, half * (WA_HALF / 8), WA_KEYS / 8);
  const bf16* bias_row = reinterpret_cast<const bf16*>(smem + WA_OFF_BIAS) + brow * WA_BIAS_PITCH + half * WA_HALF;
  const float MASK_L2 = -100.0f * 1.4426950408889634f;
  // region of this row inside a bottom / right border window (shift 3 on a 7-wide window: the seam is at index 4)
  const bool rh_i = ((brow / 7) % 7) >= 4, rw_i = (brow % 7) >= 4;
  const int nw = g.W / g.ww, nh = g.H / g.wh;
  const int lw = 31 - __clz(nw);  // H/7 and W/7 are powers of two on this path (8, 4, 2, 1)
  // in-window coordinates of this thread's query row: its token (the output row) is one cyclic wrap per item
  const int row_d = brow / 49, row_h = (brow / 7) % 7, row_w = brow % 7;
  // index of this thread among the 384 softmax threads (warps 0-9, 12, 13)
  const int st = threadIdx.x - (warp < 10 ? 0 : 64);

  float m_prev = 0.f;
  int tok_prev = 0, seg_prev = 0, head_prev = 0;
  int head_loaded = -1;

  // epilogue of one item: merge the two half-products and scatter 32 bytes through the inverse remap
  auto store_o = [&](float m_mine, float m_other, int seg, int tok, int head) {
    uint32_t oa[16], ob[16];
    const uint32_t la_u = tmem_ld_32x1(o_addr + 32), lb_u = tmem_ld_32x1(o_addr + WA_ON + 32);
    tmem_ld_32x16(o_addr + half * 16, oa);
    tmem_ld_32x16(o_addr + WA_ON + half * 16, ob);
    tmem_ld_wait();
    const float la = __uint_as_float(la_u), lb = __uint_as_float(lb_u);
    const float m_a = half ? m_other : m_mine, m_b = half ? m_mine : m_other;
    const float m = fmaxf(m_a, m_b);
    float a = ex2_approx(m_a - m), b = ex2_approx(m_b - m);
    const float inv = 1.0f / fmaf(a, la, b * lb);
    a *= inv;
    b *= inv;
    float o[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) o[i] = fmaf(a, __uint_as_float(oa[i]), b * __uint_as_float(ob[i]));
    if (row < WA_N) {
      uint4 o0, o1;
      o0.x = pack_bf16x2(o[0], o[1]); o0.y = pack_bf16x2(o[2], o[3]); o0.z = pack_bf16x2(o[4], o[5]); o0.w = pack_bf16x2(o[6], o[7]);
      o1.x = pack_bf16x2(o[8], o[9]); o1.y = pack_bf16x2(o[10], o[11]); o1.z = pack_bf16x2(o[12], o[13]); o1.w = pack_bf16x2(o[14], o[15]);
      uint4* dst = reinterpret_cast<uint4*>(out + (static_cast<size_t>(seg) * T + tok) * C + head * 32 + half * 16);
      dst[0] = o0;
      dst[1] = o1;
    }
  };

  int head = u_lo / n_items, item = u_lo - head * n_items - 1;
  for (int j = 0; j < n_my; ++j) {
    if (++item == n_items) { item = 0; ++head; }
    const int seg = item / nwin, win = item - seg * nwin;
    if (head != head_loaded) {  // (re)load this head's dense bias; uniform across the 12 softmax warps
      asm volatile("bar.sync 7, 384;" ::: "memory");
      const uint4* src = reinterpret_cast<const uint4*>(bias_dense + static_cast<size_t>(head) * WA_N * WA_BIAS_PITCH);
      uint4* dst = reinterpret_cast<uint4*>(smem + WA_OFF_BIAS);
      for (int i = st; i < WA_N * WA_BIAS_PITCH / 8; i += 384) dst[i] = __ldg(src + i);
      asm volatile("bar.sync 7, 384;" ::: "memory");
      head_loaded = head;
    }
    if (tile == 1 && (j & 1) != parity) {
      // not this warp's item: its only duty is to confirm that its previous O tile has been drained (it has: the
      // epilogue at the end of the previous iteration) before P v of this item overwrites those TMEM columns
      __syncwarp();
      if (lane == 0) mbar_arrive(sh.p_full);
      // Stay within one phase of the barriers this warp waits on (a parity wait must never lag two phases behind):
      // o_full(j) cannot be followed by o_full(j+1) before this warp's own arrival on p_full(j+1), and because the tensor
      // pipe retires S(j+1) before P v(j), its completion also proves that s_full(j+1) — the next barrier this warp
      // waits on — has completed. (Waiting on s_full(j) here would be wrong: S(j+1) does not depend on this warp and
      // could complete first, leaving the parity wait two phases behind.)
      timed_wait(sh, sh.o_full, j & 1, 3, j);
      continue;
    }
    // shift mask of this item: keys are grouped by class (remap.cuh key_slot_377), so it is one additive constant per
    // 8-column group: -100 for the classes on the far side of the seam of a border window (video_swin_ori.py:357-358)
    float gm[WA_HALF / 8];
    {
      float madd[4] = {0.f, 0.f, 0.f, 0.f};
      if (shifted) {
        const bool use_w = ((win % nw) == nw - 1) && g.sw, use_h = (((win / nw) % nh) == nh - 1) && g.sh;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const bool ch = (k >> 1) != 0, cw = (k & 1) != 0;
          madd[k] = ((use_h && ch != rh_i) || (use_w && cw != rw_i)) ? MASK_L2 : 0.f;
        }
      }
#pragma unroll
      for (int q8 = 0; q8 < WA_HALF / 8; ++q8) {
        const int grp = half * (WA_HALF / 8) + q8;
        gm[q8] = grp < 6 ? madd[0] : (grp < 11 ? madd[1] : (grp < 16 ? madd[2] : madd[3]));
      }
    }
    // the maxima ring is 4 deep: a thread's partner reads slot(j) only in its epilogue of item j, which can run while this
    // thread is already two items further (never four: P v(j+1) needs the partner's arrival after that epilogue)
    const int mslot = tile ? ((j >> 1) & 3) : (j & 3);
    timed_wait(sh, sh.s_full, j & 1, 0, j);
    tcgen05_fence_after();
    int tok;
    {
      int y = ((win >> lw) & (nh - 1)) * 7 + g.sh + row_h, x = (win & (nw - 1)) * 7 + g.sw + row_w;
      if (y >= g.H) y -= g.H;
      if (x >= g.W) x -= g.W;
      tok = (row_d * g.H + y) * g.W + x;  // == window_source_token_377(g, hW, wW, brow)
    }
    // ---- pass 1: t = s * scale * log2e + bias (+ mask), maximum of this thread's 80 columns; t stays in registers.
    // Pad columns carry a bias of -inf (lrce_window_bias_pack), so they need no special case.
    const bool timing = sh.prof != nullptr && blockIdx.x == 0 && lane == 0;
    long long tc0 = timing ? clock64() : 0;
    float t[WA_HALF];
    float mx = -INFINITY;
    {
      // all five TMEM loads are in flight together (one exposed latency); the raw scores land in t's own registers
      uint32_t* raw = reinterpret_cast<uint32_t*>(t);
#pragma unroll
      for (int c = 0; c < WA_HALF; c += 16) tmem_ld_32x16(s_addr + c, raw + c);
      tmem_ld_wait();
    }
    // S(j) now lives in registers: release the accumulator at once, so that S(j+1) is computed under this item's softmax
    tcgen05_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(sh.s_free);
    if (timing) { const long long tc = clock64(); sh.prof[warp * 8 + 4] += tc - tc0; tc0 = tc; }  // [4] TMEM load of S
#pragma unroll
    for (int c = 0; c < WA_HALF; c += 8) {
      const uint4 b4 = *reinterpret_cast<const uint4*>(bias_row + c);
      const float2 b01 = unpack_bf16x2(b4.x), b23 = unpack_bf16x2(b4.y), b45 = unpack_bf16x2(b4.z), b67 = unpack_bf16x2(b4.w);
      const float bb[8] = {b01.x, b01.y, b23.x, b23.y, b45.x, b45.y, b67.x, b67.y};
      const float ma = gm[c / 8];
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float v = fmaf(t[c + e], scale_log2e, bb[e]) + ma;
        t[c + e] = v;
        mx = fmaxf(mx, v);
      }
    }
    sM[(mslot * 2 + half) * 160 + row] = mx;  // published before this warp's arrival on p_full(j)
    if (timing) { const long long tc = clock64(); sh.prof[warp * 8 + 5] += tc - tc0; tc0 = tc; }  // [5] pass 1
    // P v of the previous item must have finished reading the P tile before it is refilled
    if (tile == 0 && j > 0) timed_wait(sh, sh.o_full, (j - 1) & 1, 1, j);
    if (timing) tc0 = clock64();
    // ---- pass 2: p = exp2(t - max) -> bf16 A-operand tile (fp32 exponent: the packed bf16x2 MUFU path was measured and
    // is not faster here — the pass is issue-bound, not MUFU-bound — so the exact-argument form is kept)
#pragma unroll
    for (int c = 0; c < WA_HALF; c += 8) {
      float p[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) p[e] = ex2_approx(t[c + e] - mx);
      uint4 u;
      u.x = pack_bf16x2(p[0], p[1]); u.y = pack_bf16x2(p[2], p[3]);
      u.z = pack_bf16x2(p[4], p[5]); u.w = pack_bf16x2(p[6], p[7]);
      *reinterpret_cast<uint4*>(p_row + (c / 8) * 128) = u;
    }
    if (timing) { const long long tc = clock64(); sh.prof[warp * 8 + 6] += tc - tc0; tc0 = tc; }  // [6] pass 2
    // ---- epilogue of the previous item of the main tile: must drain O before P v of THIS item overwrites it
    if (tile == 0 && j > 0) {
      // the partner published its maximum of item j-1 before arriving on p_full(j-1), and o_full(j-1) (waited above)
      // completed after that
      tcgen05_fence_after();
      store_o(m_prev, sM[((((j - 1) & 3)) * 2 + (half ^ 1)) * 160 + row], seg_prev, tok_prev, head_prev);
      tcgen05_fence_before();
    }
    fence_proxy_async_smem();  // P writes -> visible to the tensor core
    __syncwarp();
    if (lane == 0) mbar_arrive(sh.p_full);
    if (timing) { const long long tc = clock64(); sh.prof[warp * 8 + 7] += tc - tc0; tc0 = tc; }  // [7] epilogue + fence
    if (tile == 0) {
      m_prev = mx; tok_prev = tok; seg_prev = seg; head_prev = head;
    } else {
      // leftover rows: finish this item right away (this warp idles during the next item anyway)
      timed_wait(sh, sh.o_full, j & 1, 4, j);
      tcgen05_fence_after();
      store_o(mx, sM[(mslot * 2 + (half ^ 1)) * 160 + row], seg, tok, head);
      tcgen05_fence_before();
    }
  }
  if (tile == 0 && n_my > 0) {
    timed_wait(sh, sh.o_full, (n_my - 1) & 1, 5, n_my);
    tcgen05_fence_after();
    store_o(m_prev, sM[(((n_my - 1) & 3) * 2 + (half ^ 1)) * 160 + row], seg_prev, tok_prev, head_prev);
    tcgen05_fence_before();
  }
}

// registers are granted per group of four warps, so 14 warps get the same 128 registers per thread as 16 would
__global__ void __launch_bounds__(WA_THREADS, 1)
window_attention_kernel(const bf16* __restrict__ qkv, bf16* __restrict__ out, const bf16* __restrict__ bias_dense,
                        StageGeom g, int n_seg, int C, int n_heads, float scale_log2e, long long* prof) {
  // NOTE: pointers must stay derived from the __shared__ array itself (no integer round trip), otherwise the compiler
  // falls back to generic LD/ST for every shared-memory access
  extern __shared__ __align__(128) uint8_t smem[];
  WaShared sh;
  sh.smem = smem;
  sh.prof = prof;
  sh.qk_full = reinterpret_cast<uint64_t*>(smem + WA_OFF_BAR);  // [2]
  sh.qk_empty = sh.qk_full + 2;                                 // [2]
  sh.v_full = sh.qk_empty + 2;                                  // [2]
  sh.v_empty = sh.v_full + 2;                                   // [2]
  sh.s_full = sh.v_empty + 2;
  sh.s_free = sh.s_full + 1;
  sh.p_full = sh.s_free + 1;
  sh.o_full = sh.p_full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sh.o_full + 1);

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
  if (prof != nullptr && blockIdx.x == 0 && tid == 0) prof[16 * 8 + 2] = clock64();

  // ---- one-time setup: zero the q/k/v staging (pad rows stay zero forever), the ones column of v, barriers, TMEM
  for (int i = tid; i < 2 * WA_QKV_BYTES / 16; i += WA_THREADS) reinterpret_cast<uint4*>(smem + WA_OFF_QKV)[i] = make_uint4(0, 0, 0, 0);
  for (int i = tid; i < 2 * WA_P_TILE_BYTES / 16; i += WA_THREADS) reinterpret_cast<uint4*>(smem + WA_OFF_P)[i] = make_uint4(0, 0, 0, 0);
  __syncthreads();
  for (int i = tid; i < 2 * WA_KEYS; i += WA_THREADS) {
    const int buf = i / WA_KEYS, r = i - buf * WA_KEYS;
    *reinterpret_cast<uint4*>(smem + WA_OFF_QKV + buf * WA_QKV_BYTES + WA_Q_BYTES + WA_K_BYTES + core_off(r, 4, WA_VCH)) =
        make_uint4(0x00003F80u, 0, 0, 0);  // bf16 1.0 in dim 32, zeros in 33..39
  }
  if (warp == 11 && lane == 0) {
    *reinterpret_cast<volatile int*>(smem + WA_OFF_BAR + 200) = 0;  // watchdog abort flag (profiling hook only)
    for (int b = 0; b < 2; ++b) {
      mbar_init(&sh.qk_full[b], 1);
      mbar_init(&sh.qk_empty[b], 1);
      mbar_init(&sh.v_full[b], 1);
      mbar_init(&sh.v_empty[b], 2);
    }
    mbar_init(sh.s_full, 1);
    mbar_init(sh.s_free, 10);  // 8 main warps + the 2 leftover warps of the item
    mbar_init(sh.p_full, WA_SOFTMAX_ARRIVALS);
    mbar_init(sh.o_full, 2);  // one tcgen05.commit from each of the two MMA issuers
    fence_barrier_init();
  }
  if (warp == 10) tmem_alloc(tmem_slot, WA_TM_COLS);
  fence_proxy_async_smem();  // the fills above must be visible to the tensor core's operand reads
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 10) {
    // ===================================================================== loader (q, k, v)
    const int sub = lane >> 2, ch = lane & 3;  // 4 lanes fetch the 64 contiguous bytes of one row
    // A lane always serves the same window rows r = sub + 8k: their in-window coordinates and staging offsets are fixed
    // for the whole kernel, so the per-item work is the cyclic wrap of 19 (y, x) pairs and 3 x 19 cp.async.
    constexpr int NK = (WA_N + 7) / 8;  // 19
    uint32_t r_hw[NK], r_dq[NK], r_dk[NK], r_dv[NK];
    int r_tok0[NK];
#pragma unroll
    for (int k = 0; k < NK; ++k) {
      const int r = min(sub + 8 * k, WA_N - 1);
      const int d = r / 49, rem = r - d * 49, h = rem / 7, w = rem - h * 7;
      r_hw[k] = static_cast<uint32_t>(h << 8 | w);
      r_tok0[k] = d * g.H * g.W;
      // q rows 128..146 go to row tile 1 (lane quarter 0; + 32 rows for odd items, added per item); k / v rows go to their
      // class-grouped slot (remap.cuh key_slot_377)
      const int slot = key_slot_377(r);
      r_dq[k] = core_off(r, ch, 4);
      r_dk[k] = WA_Q_BYTES + core_off(slot, ch, 4);
      r_dv[k] = WA_Q_BYTES + WA_K_BYTES + core_off(slot, ch, WA_VCH);
    }
    const int lw = 31 - __clz(g.W / 7);  // H/7 and W/7 are powers of two on this path (8, 4, 2, 1)
    const int nh_mask = g.H / 7 - 1, nw_mask = g.W / 7 - 1;
    int head = u_lo / n_items, item = u_lo - head * n_items;
    for (int j = 0; j < n_my; ++j) {
      const int buf = j & 1;
      const int seg = item / nwin, win = item - seg * nwin;
      const int ybase = ((win >> lw) & nh_mask) * 7 + g.sh, xbase = (win & nw_mask) * 7 + g.sw;
      const bf16* base = qkv + static_cast<size_t>(seg) * T * 3 * C + head * 32 + ch * 8;
      const uint32_t sbuf = smem_u32(smem + WA_OFF_QKV + buf * WA_QKV_BYTES);
      const uint32_t left_off = buf ? 32 * 64 : 0;  // odd items: 32 rows further down (TMEM lane quarter 1)
      int tok[NK];
#pragma unroll
      for (int k = 0; k < NK; ++k) {
        int y = ybase + static_cast<int>(r_hw[k] >> 8), x = xbase + static_cast<int>(r_hw[k] & 0xff);
        if (y >= g.H) y -= g.H;
        if (x >= g.W) x -= g.W;
        tok[k] = (r_tok0[k] + y * g.W + x) * 3 * C;
      }
      timed_wait(sh, &sh.qk_empty[buf], ((j >> 1) & 1) ^ 1, 0, j);
      const long long tl0 = clock64();
#pragma unroll
      for (int k = 0; k < NK; ++k)
        if (sub + 8 * k < WA_N) cp_async_16(sbuf + r_dq[k] + (k >= 16 ? left_off : 0u), base + tok[k]);
#pragma unroll
      for (int k = 0; k < NK; ++k)
        if (sub + 8 * k < WA_N) cp_async_16(sbuf + r_dk[k], base + C + tok[k]);
      asm volatile("cp.async.commit_group;" ::: "memory");
      timed_wait(sh, &sh.v_empty[buf], ((j >> 1) & 1) ^ 1, 3, j);
#pragma unroll
      for (int k = 0; k < NK; ++k)
        if (sub + 8 * k < WA_N) cp_async_16(sbuf + r_dv[k], base + 2 * C + tok[k]);
      asm volatile("cp.async.commit_group;" ::: "memory");
      const long long tl1 = clock64();
      asm volatile("cp.async.wait_group 1;" ::: "memory");
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&sh.qk_full[buf]);
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&sh.v_full[buf]);
      if (sh.prof != nullptr && lane == 0 && blockIdx.x == 0) {
        sh.prof[warp * 8 + 1] += tl1 - tl0;        // cp.async issue (incl. the wait for the v buffer)
        sh.prof[warp * 8 + 2] += clock64() - tl1;  // waiting for the data
      }
      if (++item == n_items) { item = 0; ++head; }
    }
  } else if (warp == 11) {
    // ===================================================================== MMA issuer
    if (lane == 0 && n_my > 0) {
      constexpr uint32_t idesc_s = umma_idesc_bf16(128, WA_KEYS);            // A, B K-major
      constexpr uint32_t idesc_o = umma_idesc_bf16(128, WA_ON) | (1u << 16);  // B (= v) MN-major
      // Descriptors of one operand differ only in their start-address field (bits 0..13, in 16-byte units), so every
      // further MMA of a sequence costs one 64-bit add instead of a fresh encode.
      const uint32_t smem0 = smem_u32(smem);
      auto issue_s = [&](int j) {
        const uint32_t b = smem0 + WA_OFF_QKV + (j & 1) * WA_QKV_BYTES;
        const uint64_t dq = umma_desc_nosw(b, 128, 512), dk = umma_desc_nosw(b + WA_Q_BYTES, 128, 512);
#pragma unroll
        for (int tile = 0; tile < 2; ++tile)
#pragma unroll
          for (int kk = 0; kk < 2; ++kk)
            umma_bf16_ss(tmem_base + (tile ? WA_TM_S1 : WA_TM_S0), dq + ((tile * (16 * 512) + kk * 256) >> 4), dk + ((kk * 256) >> 4),
                         idesc_s, kk);
        umma_commit(sh.s_full);
        umma_commit(&sh.qk_empty[j & 1]);
      };
      mbar_wait(&sh.qk_full[0], 0);
      tcgen05_fence_after();
      issue_s(0);
      for (int j = 0; j < n_my; ++j) {
        if (j + 1 < n_my) {
          timed_wait(sh, sh.s_free, j & 1, 0, j);  // every softmax warp holds S(j) in registers
          timed_wait(sh, &sh.qk_full[(j + 1) & 1], ((j + 1) >> 1) & 1, 1, j);
          tcgen05_fence_after();
          issue_s(j + 1);
        }
        timed_wait(sh, sh.p_full, j & 1, 2, j);  // P(j) in smem, O(j-1) drained
        timed_wait(sh, &sh.v_full[j & 1], (j >> 1) & 1, 3, j);
        tcgen05_fence_after();
        const uint64_t dv = umma_desc_nosw(smem0 + WA_OFF_QKV + (j & 1) * WA_QKV_BYTES + WA_Q_BYTES + WA_K_BYTES,
                                           /*lbo: key groups*/ WA_VCH * 128, /*sbo: dim groups*/ 128);
        const uint64_t dp = umma_desc_nosw(smem0 + WA_OFF_P, 128, (WA_KEYS / 8) * 128);
#pragma unroll
        for (int kk = 0; kk < WA_KEYS / 16; ++kk)  // row tile 0; row tile 1 is issued by warp 14
          umma_bf16_ss(tmem_base + WA_TM_O + (kk >= WA_HALF / 16 ? 1 : 0) * WA_ON, dp + ((kk * 256) >> 4),
                       dv + ((kk * (2 * WA_VCH * 128)) >> 4), idesc_o, (kk % (WA_HALF / 16)) != 0);
        umma_commit(sh.o_full);
        umma_commit(&sh.v_empty[j & 1]);
      }
    }
  } else if (warp == 14) {
    // ===================================================================== second MMA issuer: P v of row tile 1
    if (lane == 0) {
      constexpr uint32_t idesc_o = umma_idesc_bf16(128, WA_ON) | (1u << 16);  // B (= v) MN-major
      const uint32_t smem0 = smem_u32(smem);
      for (int j = 0; j < n_my; ++j) {
        timed_wait(sh, sh.p_full, j & 1, 2, j);
        timed_wait(sh, &sh.v_full[j & 1], (j >> 1) & 1, 3, j);
        tcgen05_fence_after();
        const uint64_t dv = umma_desc_nosw(smem0 + WA_OFF_QKV + (j & 1) * WA_QKV_BYTES + WA_Q_BYTES + WA_K_BYTES,
                                           /*lbo: key groups*/ WA_VCH * 128, /*sbo: dim groups*/ 128);
        const uint64_t dp = umma_desc_nosw(smem0 + WA_OFF_P + WA_P_TILE_BYTES, 128, (WA_KEYS / 8) * 128);
#pragma unroll
        for (int kk = 0; kk < WA_KEYS / 16; ++kk)
          umma_bf16_ss(tmem_base + WA_TM_O + (2 + (kk >= WA_HALF / 16 ? 1 : 0)) * WA_ON, dp + ((kk * 256) >> 4),
                       dv + ((kk * (2 * WA_VCH * 128)) >> 4), idesc_o, (kk % (WA_HALF / 16)) != 0);
        umma_commit(sh.o_full);
        umma_commit(&sh.v_empty[j & 1]);
      }
    }
  } else {
    // ===================================================================== softmax + epilogue (12 warps)
    const int tile = warp >= 8;
    const int half = tile ? (warp >= 12) : (warp >> 2);
    const int parity = warp & 1;  // warps 8, 12: even items (TMEM lanes 0..31); 9, 13: odd items (lanes 32..63)
    softmax_warp(sh, tmem_base, tile, half, tile ? parity * 32 : (warp & 3) * 32, parity, bias_dense, out, g, n_items, nwin, T,
                 C, u_lo, n_my, scale_log2e, shifted);
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 10) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, WA_TM_COLS);
  }
  if (prof != nullptr && blockIdx.x == 0 && tid == 0) {
    prof[16 * 8] = n_my;
    prof[16 * 8 + 1] = clock64() - prof[16 * 8 + 2];
  }
}

// bias_dense[h][i][key_slot(j)] = table[rel_index(i, j)][h] * log2(e); the 13 pad columns of the score tile hold -inf
__global__ void build_dense_bias_kernel(const float* __restrict__ table, bf16* __restrict__ dense, StageGeom g,
                                        int n_heads) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int total = n_heads * WA_N * WA_BIAS_PITCH;
  if (idx >= total) return;
  const int j = idx % WA_BIAS_PITCH, i = (idx / WA_BIAS_PITCH) % WA_N, h = idx / (WA_BIAS_PITCH * WA_N);
  if (j < WA_N) {
    const int rel = rel_pos_offset(g, i) - rel_pos_offset(g, j) + REL_POS_CENTER;
    dense[idx - j + key_slot_377(j)] = __float2bfloat16(table[static_cast<size_t>(rel) * n_heads + h] * 1.4426950408889634f);
  } else {
    // thread j = 147 + p fills the p-th pad column: 4 after class 1, 4 after class 2, 5 after class 3
    const int p = j - WA_N;
    const int col = p < 4 ? 84 + p : (p < 8 ? 124 + (p - 4) : 155 + (p - 8));
    dense[idx - j + col] = __float2bfloat16(-INFINITY);
  }
}

}  // namespace lrce

using namespace lrce;

static long long* g_attn_prof = nullptr;
// Profiling hook, not part of the product path: when buf (device, 16*8+3 zeroed int64) is non-NULL, CTA 0 of the following
// lrce_window_attention_bf16 launches (buf: 16*8+3+1+16 zeroed int64) accumulates, per warp, the cycles spent in each kind of mbarrier wait
// (softmax warps: [0] S ready, [1] P v done, [2] foreign-item waits; loaders: [0] buffer free, [1] issue, [2] data landed;
// MMA: [0] S released, [1] q/k landed, [2] P ready, [3] v landed) plus [128] items and [129] total cycles of CTA 0.
extern "C" int lrce_debug_attention_timing(long long* buf) {
  g_attn_prof = buf;
  return LRCE_OK;
}

static int geom_3x7x7(StageGeom* g, int D, int H, int W, int sh, int sw) {
  LRCE_REQUIRE(D == 3 && H % 7 == 0 && W % 7 == 0 && H > 0 && W > 0,
               "window attention is specialised for the clamped (3,7,7) window of LRCE's 5-frame segments; got grid "
               "(%d,%d,%d)", D, H, W);
  LRCE_REQUIRE((sh == 0 || sh == 3) && (sw == 0 || sw == 3),
               "window attention is specialised for Swin's shift = window // 2 = 3 (or 0); got (%d,%d)", sh, sw);
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
      n_seg, C, n_heads, scale_log2e, g_attn_prof);
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
