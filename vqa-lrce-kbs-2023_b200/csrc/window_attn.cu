// window_attn.cu — (shifted-)window multi-head attention core of Video Swin (video_swin_ori.py:166-186) on tcgen05
// tensor cores, with the cyclic shift / window_partition / window_reverse remap (video_swin_ori.py:262-276) fused into
// its TMA loads and its stores.
//
// Input  : qkv  bf16 [n_seg * D*H*W, 3C] in NATURAL token order (the qkv Linear is per token, so it runs before any
//          partition); column layout [q | k | v][head][32] (video_swin_ori.py:165).
// Output : out  bf16 [n_seg * D*H*W, C] in natural token order, heads merged (video_swin_ori.py:186), i.e. exactly
//          roll(window_reverse(attn @ v), +shift) — the proj GEMM + residual then runs with no remap at all.
//
// Remap as TMA boxes. A (3,7,7) window of the rolled frame splits at the shift seam (h = 4, w = 4) into four boxes
// (4|3 rows) x (4|3 columns) x 3 frames, none of which wraps around the frame border; each box is one
// cp.async.bulk.tensor.4d over the view (channel, w, h, frame*segment) of qkv, for q, k and v alike, and lands as dense
// 64-byte rows (64B-swizzled) in shared memory. The four boxes are at the same time the four classes of the shift mask
// (video_swin_ori.py:346-358): rows / keys are therefore held in "slot" order (remap.cuh key_slot_377: class 0 -> slots
// [0,48), 1 -> [48,84) + 4 pads, 2 -> [88,124) + 4 pads, 3 -> [128,155) + 5 pads), the mask is one additive constant per
// 8-column group of the score tile, and window_source_token_377() is the token a slot maps to (bit-exact test through
// lrce_remap_index). Pad slots are never written: they stay zero from the one-time fill.
//
// One persistent CTA per SM walks a contiguous range of (head, segment, window) work units; 24 warps:
//   warp 20      loader   : one thread, 15 TMA boxes per unit (q, k, v x 4 classes; the class-3 q box four times, see
//                           below), double-buffered, completion on mbarriers.
//   warps 21,22  MMA      : one thread each, row tile 0 (slots 0..127) / row tile 1 (slots 128..159):
//                           S = q k^T [M=128, N=160, K=32] (both operands K-major, 64B swizzle), then O = P v
//                           [M=128, N=48, K=160] with B = [v | 1]: v consumed MN-major exactly as TMA delivered it, plus
//                           a second N atom of ones, so that columns 32..47 are the row sums L = P 1 of the bf16
//                           probabilities: they come from the tensor pipe, which idles, with the A reads P v pays anyway,
//                           not from the softmax warps, which are the bottleneck.
//                           Accumulators in TMEM: S0, S1 (2 x 160 columns), O | L double-buffered per tile (4 x 48).
//   warps 0-15   softmax  : row tile 0, four warps per TMEM lane quarter (thread = row, warp = 40 key columns).
//   warps 16-19  softmax  : row tile 1. Its 32 slots are loaded FOUR times into the q tile, so every TMEM lane quarter of
//                           S1 holds the same 32 rows and warp 16+i (lane quarter i) takes key columns [40 i, 40 i + 40):
//                           all 20 softmax warps run the same code on 40 scores per thread, five per scheduler.
// Softmax without a dependent pass over the scores: the stabiliser is an upper bound of the row maximum,
// m = scale * max_j s_ij + max_j bias_ij (raw-score maxima exchanged between the four warps of a row through shared
// memory and a 128-thread named barrier; max_j bias_ij precomputed per row by lrce_window_bias_pack) — softmax is
// invariant to the shift, the shift mask only lowers scores, and bf16 probabilities keep their relative precision under
// a bound that is loose by a few units. Then p = exp2(s * scale*log2e + bias + mask - m) straight into bf16 A-operand
// tiles (double-buffered, so no warp ever waits for P v of the previous unit). The epilogue of unit j-1 (O and L from
// TMEM, 1 / row sum, 16-byte stores through the inverse remap) runs after the probabilities of unit j are handed to the
// tensor core.
// The S accumulator is released as soon as a warp has its 40 scores in registers, so S(j+1) is computed under the
// softmax of unit j.
#include <stdlib.h>

#include "host_common.h"
#include "lrce_common.cuh"
#include "remap.cuh"

namespace lrce {

constexpr int WA_N = 147;           // tokens per (3,7,7) window
constexpr int WA_KEYS = 160;        // row / key slots of a window (class-grouped, 13 pads)
constexpr int WA_QCOLS = 40;        // key columns per softmax thread
constexpr int WA_BIAS_PITCH = 160;  // dense bias row pitch (bf16) = key columns of the score tile
constexpr int WA_ON = 48;           // TMEM columns of one output buffer: 32 dims + 16 row-sum columns (L = P 1)
constexpr int WA_THREADS = 24 * 32;
constexpr int WA_SOFTMAX_THREADS = 20 * 32;
constexpr int WA_WARP_LOADER = 20, WA_WARP_MMA0 = 21, WA_WARP_MMA1 = 22, WA_WARP_TMEM = 23;

// shared memory map (bytes). Every TMA destination / swizzled UMMA operand start is a multiple of 512 B (64B-swizzle
// pattern period); the dynamic shared memory window itself starts 1024-byte aligned (checked at kernel start).
constexpr int WA_Q_BYTES = 256 * 64;  // slots 0..127 (row tile 0) + 4 copies of slots 128..159 (row tile 1)
constexpr int WA_K_BYTES = WA_KEYS * 64;
constexpr int WA_V_BYTES = WA_KEYS * 64;
constexpr int WA_STAGE_BYTES = WA_Q_BYTES + WA_K_BYTES + WA_V_BYTES;  // 36864
constexpr int WA_P0_BYTES = 128 * WA_KEYS * 2;                        // 40960: P of row tile 0
constexpr int WA_P1_BYTES = 32 * WA_KEYS * 2;  // 10240: P of row tile 1 (the MMA's rows 32..127 alias what follows: unused lanes)
constexpr int WA_BIAS_ROWS = 155;                                    // last valid slot + 1
constexpr int WA_BIAS_HEAD_BYTES = WA_KEYS * WA_BIAS_PITCH * 2;      // 51200 per head in global memory
constexpr int WA_BIAS_COPY_BYTES = (WA_BIAS_ROWS + 1) * WA_BIAS_PITCH * 2;  // rows 0..154 + bf16 bmax[160] (row 155)
constexpr int WA_OFF_STAGE = 0;
constexpr int WA_OFF_P0 = 2 * WA_STAGE_BYTES;
constexpr int WA_OFF_P1 = WA_OFF_P0 + 2 * WA_P0_BYTES;
constexpr int WA_OFF_BIAS = WA_OFF_P1 + 2 * WA_P1_BYTES;
constexpr int WA_OFF_BMAX = WA_OFF_BIAS + WA_BIAS_ROWS * WA_BIAS_PITCH * 2;
constexpr int WA_OFF_BAR = WA_OFF_BIAS + WA_BIAS_COPY_BYTES;   // mbarriers, TMEM slot, watchdog flag: 192 B
constexpr int WA_OFF_ONES = WA_OFF_BAR + 192;  // 16 keys x 64 B of bf16 ones: the second N atom of the P v operand (row sums L = P 1)
constexpr int WA_OFF_M = WA_OFF_BIAS + WA_BIAS_HEAD_BYTES;     // float [2][4][128] + [2][4][32]: partial raw-score maxima
constexpr int WA_SMEM = WA_OFF_M + (2 * 4 * 128 + 2 * 4 * 32) * 4;
static_assert(WA_OFF_ONES + 1024 <= WA_OFF_M && WA_OFF_ONES % 64 == 0, "barrier block and ones tile must fit behind the bias rows");
static_assert(WA_SMEM <= 227 * 1024, "window attention shared-memory budget");
static_assert(WA_OFF_P1 + WA_P1_BYTES + 128 * WA_KEYS * 2 <= WA_SMEM, "row tile 1's A operand must stay inside shared memory");

// TMEM columns
constexpr int WA_TM_S0 = 0, WA_TM_S1 = 160, WA_TM_O0 = 320, WA_TM_O1 = 320 + 2 * WA_ON, WA_TM_COLS = 512;
static_assert(WA_TM_O1 + 2 * WA_ON <= WA_TM_COLS, "TMEM budget");

constexpr uint32_t WA_QK_TX_BYTES = (WA_N + 3 * 27) * 64 + WA_N * 64;  // q (class 3 four times) + k
constexpr uint32_t WA_V_TX_BYTES = WA_N * 64;

struct WaMaps {
  CUtensorMap m[4];  // boxes (32 channels, bw, bh, 3 frames) with (bw, bh) = (4,4), (3,4), (4,3), (3,3) = mask class 0..3
};

// UMMA shared-memory descriptors. No swizzle: operands are 8-row x 16-byte "core matrices" (128 contiguous bytes); lbo /
// sbo are the byte distances between core matrices along K and along M/N. 64B swizzle: dense 64-byte rows, 16-byte chunk
// index XOR ((row >> 1) & 3), groups of 8 rows 512 B apart — what TMA writes with CU_TENSOR_MAP_SWIZZLE_64B; the same
// bytes serve as K-major operand (q, k: row = M/N index, 64 B = 32 K elements) and as MN-major operand (v: row = K
// index, 64 B = 32 N elements; sbo = distance between groups of 8 K rows).
__device__ __forceinline__ uint64_t umma_desc_nosw(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(lbo_bytes >> 4) << 16;
  d |= static_cast<uint64_t>(sbo_bytes >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;  // descriptor version (Blackwell)
  return d;                             // layout type 0 = no swizzle
}
__device__ __forceinline__ uint64_t umma_desc_sw64(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(1) << 16;          // lbo: unused (one swizzle atom along the contiguous direction)
  d |= static_cast<uint64_t>(512 >> 4) << 32;   // sbo: 8 rows x 64 B
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(4) << 61;          // layout type 4 = SWIZZLE_64B
  return d;
}

// MN-major 64B-swizzle operand wider than one 32-element atom: atom a of the N direction starts lbo_bytes * a behind the first
__device__ __forceinline__ uint64_t umma_desc_sw64_lbo(uint32_t smem_addr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(lbo_bytes >> 4) << 16;
  d |= static_cast<uint64_t>(512 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(4) << 61;
  return d;
}

// q / k / v slices are read once (by this CTA and, at the same moment, by its partner on the other head of the line) and are dead
// afterwards: `policy` = L2 evict_first keeps them from displacing the attention output the projection GEMM reads next
__device__ __forceinline__ void tma_load_4d(uint32_t smem_dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1, int c2,
                                            int c3, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4, %5, %6}], [%2], "
      "%7;" ::"r"(smem_dst),
      "l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "l"(policy)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_32x8(uint32_t taddr, uint32_t* v) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ uint32_t tmem_ld_32x1(uint32_t taddr) {
  uint32_t v;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(v) : "r"(taddr) : "memory");
  return v;
}
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// byte offset of the 16-byte chunk (row, chunk) inside a no-swizzle operand with `cores_per_group` cores along K
__device__ __forceinline__ uint32_t core_off(int row, int chunk, int cores_per_group) {
  return static_cast<uint32_t>(((row >> 3) * cores_per_group + chunk) * 128 + (row & 7) * 16);
}

struct WaShared {
  uint8_t* smem;
  uint64_t *qk_full, *qk_empty, *v_full, *v_empty;  // [2] each: staging buffers
  uint64_t *s_full, *s_free;                        // [2]: per row tile
  uint64_t *p_full, *o_full;                        // [4]: [row tile][buffer]
  long long* prof;  // lrce_window_attention_profile: [24 warps][8] cycle counters of CTA 0, or nullptr
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

// mbarrier wait. PROF (profiling hook armed): accounts the stall to counter `slot` of this warp and doubles as a watchdog —
// a wait longer than ~50 ms records (warp, slot, item) in prof[196..] and raises a CTA-wide abort flag that makes every
// later wait fall through, so a protocol deadlock ends with a report, not a hang. Production: parked wait, no spinning.
template <bool PROF>
__device__ __forceinline__ void timed_wait(const WaShared& sh, uint64_t* bar, uint32_t parity, int slot, int item = -1) {
  if (!PROF) {
    // Non-blocking test_wait in a tight loop, not the parked try_wait of the other kernels: every wait of this kernel sits on the
    // per-unit dependency chain (S ready -> softmax -> P ready -> P v -> O ready), and the wake-up of a parked warp costs more
    // than the issue slots the spinning warps take (forward, same box: 15.40 -> 15.20 ms; the walk kernel is the opposite case:
    // 818 -> 850 us with spinning waits, its eight compute warps need the slots)
    uint32_t done = 0;
    while (!done)
      asm volatile("{\n.reg .pred p;\nmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                   : "=r"(done)
                   : "r"(smem_u32(bar)), "r"(parity)
                   : "memory");
    return;
  }
  volatile int* abort_flag = reinterpret_cast<volatile int*>(sh.smem + WA_OFF_BAR + 184);
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (*abort_flag) break;
    if (clock64() - t0 > 100000000LL) {
      *abort_flag = 1;
      sh.prof[196 + (threadIdx.x >> 5)] = (static_cast<long long>(blockIdx.x) << 40) | (static_cast<long long>(slot + 1) << 32) |
                                          static_cast<unsigned>(item);
      break;
    }
  }
  if ((threadIdx.x & 31) == 0 && blockIdx.x == 0) sh.prof[(threadIdx.x >> 5) * 8 + slot] += clock64() - t0;
}

struct WaItemCtx {
  const bf16* bias_dense;
  bf16* out;
  StageGeom g;
  int n_items, nwin, T, C, head0, item0, head_step, n_my;
  float scale_log2e;
  bool shifted;
};

// (re)load the dense bias rows + row maxima of `head` when the work range crosses a head boundary; all 20 softmax warps
__device__ __forceinline__ void reload_bias(uint8_t* smem, const bf16* bias_dense, int head) {
  asm volatile("bar.sync 7, 640;" ::: "memory");
  const uint4* src = reinterpret_cast<const uint4*>(reinterpret_cast<const uint8_t*>(bias_dense) +
                                                    static_cast<size_t>(head) * WA_BIAS_HEAD_BYTES);
  uint4* dst = reinterpret_cast<uint4*>(smem + WA_OFF_BIAS);
  for (int i = threadIdx.x; i < WA_BIAS_COPY_BYTES / 16; i += WA_SOFTMAX_THREADS) dst[i] = __ldg(src + i);
  asm volatile("bar.sync 7, 640;" ::: "memory");
}

// Softmax warp. TILE 0: row slots 32 q + lane of every unit, key columns [40 c, 40 c + 40), q = warp & 3, c = warp >> 2.
// TILE 1: row slots 128 + lane (replicated in every TMEM lane quarter), key columns [40 q, 40 q + 40), q = c = warp & 3.
template <int TILE, bool PROF>
__device__ __forceinline__ void softmax_warp(const WaShared& sh, uint32_t tmem_base, int q, int c, const WaItemCtx& cx) {
  // the pointer must be derived from the __shared__ array itself, otherwise every access below compiles to generic LD/ST
  extern __shared__ __align__(1024) uint8_t smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const StageGeom& g = cx.g;
  const int slot = TILE == 0 ? q * 32 + lane : 128 + lane;  // row slot of this thread
  const int prow = TILE == 0 ? slot : lane;                 // row inside the tile's P operand
  float* sMax = reinterpret_cast<float*>(smem + WA_OFF_M) + (TILE == 0 ? 0 : 2 * 4 * 128);
  constexpr int mrows = TILE == 0 ? 128 : 32;
  const uint32_t lane_addr = static_cast<uint32_t>(q * 32) << 16;
  const uint32_t s_addr = tmem_base + lane_addr + (TILE == 0 ? WA_TM_S0 : WA_TM_S1) + c * WA_QCOLS;
  const uint32_t o_base = tmem_base + lane_addr + (TILE == 0 ? WA_TM_O0 : WA_TM_O1);
  uint8_t* p_base = smem + (TILE == 0 ? WA_OFF_P0 : WA_OFF_P1) + core_off(prow, c * (WA_QCOLS / 8), WA_KEYS / 8);
  constexpr int p_stride = TILE == 0 ? WA_P0_BYTES : WA_P1_BYTES;
  const int tokw = slot_token_377(slot);  // window token of this row, -1 for a pad slot
  const bool valid = tokw >= 0;
  const int brow = valid ? slot : 0;
  // bias rows are stored with their 16-byte chunks XOR-swizzled by ((row >> 1) & 3) (lrce_window_bias_pack): the eight
  // lanes of an LDS.128 phase read the same logical chunk of eight consecutive 320-byte rows, which would otherwise be a
  // 4-way bank conflict
  const bf16* bias_row = reinterpret_cast<const bf16*>(smem + WA_OFF_BIAS) + brow * WA_BIAS_PITCH;
  const int bias_sw = (brow >> 1) & 3;
  const bf16* bmax = reinterpret_cast<const bf16*>(smem + WA_OFF_BMAX) + brow;
  // shift mask (video_swin_ori.py:346-358) of a bottom / right border window: key group k of this warp is masked for this
  // row iff they lie on different sides of the h seam (bit k of mh) or of the w seam (bit k of mw)
  const int row_cls = (slot >= 48) + (slot >= 88) + (slot >= 128);
  uint32_t mh = 0, mw = 0;
#pragma unroll
  for (int k = 0; k < WA_QCOLS / 8; ++k) {
    const int kc = key_group_class_377(c * (WA_QCOLS / 8) + k);
    mh |= static_cast<uint32_t>((kc >> 1) != (row_cls >> 1)) << k;
    mw |= static_cast<uint32_t>((kc & 1) != (row_cls & 1)) << k;
  }
  const int nw = g.W / g.ww, nh = g.H / g.wh;
  const int lw = 31 - __clz(nw);  // H/7 and W/7 are powers of two on this path (8, 4, 2, 1)
  const int tw = valid ? tokw : 0;
  const int row_d = tw / 49, row_h = (tw / 7) % 7, row_w = tw % 7;
  // who stores: tile 0 — every warp its 8 of the 32 dims; tile 1 — lane quarter 0 holds the rows, all 32 dims
  const bool stores = TILE == 0 || q == 0;
  const int bar_id = TILE == 0 ? 1 + q : 5;
  uint64_t* s_full = sh.s_full + TILE;
  uint64_t* s_free = sh.s_free + TILE;
  uint64_t* p_full = sh.p_full + 2 * TILE;
  uint64_t* o_full = sh.o_full + 2 * TILE;
  const bool timing = PROF && blockIdx.x == 0 && lane == 0;
  const float MASK_L2 = -100.0f * 1.4426950408889634f;

  bf16* dst_prev = cx.out;     // output row of the previous unit

  // epilogue of unit jp: normalise by the row sum (column 32 of the output buffer, L = P 1) and scatter through the
  // inverse remap
  auto store_o = [&](int jp) {
    timed_wait<PROF>(sh, o_full + (jp & 1), (jp >> 1) & 1, 1, jp);
    if (!stores) return;
    tcgen05_fence_after();
    const uint32_t o_addr = o_base + (jp & 1) * WA_ON;
    if (TILE == 0) {
      uint32_t o8[8];
      const uint32_t l_u = tmem_ld_32x1(o_addr + 32);
      tmem_ld_32x8(o_addr + c * 8, o8);
      tmem_ld_wait();
      tcgen05_fence_before();
      const float inv = 1.0f / __uint_as_float(l_u);
      uint4 o;
      o.x = pack_bf16x2(__uint_as_float(o8[0]) * inv, __uint_as_float(o8[1]) * inv);
      o.y = pack_bf16x2(__uint_as_float(o8[2]) * inv, __uint_as_float(o8[3]) * inv);
      o.z = pack_bf16x2(__uint_as_float(o8[4]) * inv, __uint_as_float(o8[5]) * inv);
      o.w = pack_bf16x2(__uint_as_float(o8[6]) * inv, __uint_as_float(o8[7]) * inv);
      if (valid) *reinterpret_cast<uint4*>(dst_prev + c * 8) = o;
    } else {
      uint32_t o32[32];
      tmem_ld_32x32(o_addr, o32);
      const float inv = 1.0f / __uint_as_float(tmem_ld_32x1(o_addr + 32));
      tmem_ld_wait();
      tcgen05_fence_before();
      if (valid) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          uint4 o;
          o.x = pack_bf16x2(__uint_as_float(o32[8 * k + 0]) * inv, __uint_as_float(o32[8 * k + 1]) * inv);
          o.y = pack_bf16x2(__uint_as_float(o32[8 * k + 2]) * inv, __uint_as_float(o32[8 * k + 3]) * inv);
          o.z = pack_bf16x2(__uint_as_float(o32[8 * k + 4]) * inv, __uint_as_float(o32[8 * k + 5]) * inv);
          o.w = pack_bf16x2(__uint_as_float(o32[8 * k + 6]) * inv, __uint_as_float(o32[8 * k + 7]) * inv);
          reinterpret_cast<uint4*>(dst_prev)[k] = o;
        }
      }
    }
  };

  int head = cx.head0, item = cx.item0;
  int seg = item / cx.nwin, win = item - seg * cx.nwin;
  int head_loaded = -1;
  for (int j = 0; j < cx.n_my; ++j) {
    if (head != head_loaded) {
      reload_bias(smem, cx.bias_dense, head);
      head_loaded = head;
    }
    // per-unit geometry of this row: mask bits of the five key groups, output row
    const int wy = (win >> lw) & (nh - 1), wx = win & (nw - 1);
    uint32_t mbits = 0;
    if (cx.shifted) mbits = ((wy == nh - 1 && g.sh) ? mh : 0u) | ((wx == nw - 1 && g.sw) ? mw : 0u);
    bf16* dst;
    {
      int y = wy * 7 + g.sh + row_h, x = wx * 7 + g.sw + row_w;
      if (y >= g.H) y -= g.H;
      if (x >= g.W) x -= g.W;
      const int tok = (row_d * g.H + y) * g.W + x;  // == window_source_token_377(g, wy, wx, tokw)
      dst = cx.out + (static_cast<size_t>(seg) * cx.T + tok) * cx.C + head * 32;
    }
    timed_wait<PROF>(sh, s_full, j & 1, 0, j);
    tcgen05_fence_after();
    long long tc0 = timing ? clock64() : 0;
    float s[WA_QCOLS];
    {
      uint32_t* raw = reinterpret_cast<uint32_t*>(s);
      tmem_ld_32x32(s_addr, raw);
      tmem_ld_32x8(s_addr + 32, raw + 32);
      tmem_ld_wait();
    }
    // S(j) now lives in registers: release the accumulator at once, so that S(j+1) is computed under this unit's softmax
    tcgen05_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(s_free);
    if (timing) { const long long tc = clock64(); sh.prof[warp * 8 + 4] += tc - tc0; tc0 = tc; }  // [4] TMEM load of S
    // ---- raw-score maximum of the row: own 40 columns, then the four warps of the row through shared memory
    float mx = s[0];
#pragma unroll
    for (int e = 1; e < WA_QCOLS; ++e) mx = fmaxf(mx, s[e]);
    sMax[((j & 1) * 4 + c) * mrows + prow] = mx;
    asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
#pragma unroll
    for (int k = 0; k < 4; ++k) mx = fmaxf(mx, sMax[((j & 1) * 4 + k) * mrows + prow]);
    const float nbound = -fmaf(mx, cx.scale_log2e, __bfloat162float(*bmax));  // -(upper bound of every t_ij of the row)
    if (timing) { const long long tc = clock64(); sh.prof[warp * 8 + 5] += tc - tc0; tc0 = tc; }  // [5] maximum + exchange
    // ---- p = exp2(s * scale*log2e + bias + mask - bound) -> bf16 A-operand tile (buffer j & 1: P v(j-2) has completed,
    // observed at the epilogue of unit j-2); pad columns carry a bias of -inf
    uint8_t* p_row = p_base + (j & 1) * p_stride;
    const float2 sc2 = make_float2(cx.scale_log2e, cx.scale_log2e);
#pragma unroll
    for (int cc = 0; cc < WA_QCOLS; cc += 8) {
      const uint4 b4 = *reinterpret_cast<const uint4*>(bias_row + ((c * (WA_QCOLS / 8) + cc / 8) ^ bias_sw) * 8);
      const float cg = ((mbits >> (cc / 8)) & 1u) ? nbound + MASK_L2 : nbound;
      const float2 cg2 = make_float2(cg, cg);
      const uint32_t bw[4] = {b4.x, b4.y, b4.z, b4.w};
      float2 p[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 t = ffma2(make_float2(s[cc + 2 * e], s[cc + 2 * e + 1]), sc2, fadd2(bf16x2_to_f32x2(bw[e]), cg2));
        p[e] = make_float2(ex2_approx(t.x), ex2_approx(t.y));
      }
      uint4 u;
      u.x = pack_bf16x2(p[0].x, p[0].y); u.y = pack_bf16x2(p[1].x, p[1].y);
      u.z = pack_bf16x2(p[2].x, p[2].y); u.w = pack_bf16x2(p[3].x, p[3].y);
      *reinterpret_cast<uint4*>(p_row + (cc / 8) * 128) = u;
    }
    fence_proxy_async_smem();  // P writes -> visible to the tensor core
    __syncwarp();
    if (lane == 0) mbar_arrive(p_full + (j & 1));
    if (timing) { const long long tc = clock64(); sh.prof[warp * 8 + 6] += tc - tc0; tc0 = tc; }  // [6] probabilities
    // ---- epilogue of the previous unit (its P v was issued a whole softmax ago). Program order puts these TMEM reads
    // before this warp's next arrival on p_full, i.e. before P v(j+1) overwrites the same output buffer.
    if (j > 0) store_o(j - 1);
    if (timing) { const long long tc = clock64(); sh.prof[warp * 8 + 7] += tc - tc0; tc0 = tc; }  // [7] epilogue
    dst_prev = dst;
    if (++win == cx.nwin) {
      win = 0;
      if (++seg * cx.nwin == cx.n_items) { seg = 0; head += cx.head_step; }
    }
  }
  if (cx.n_my > 0) store_o(cx.n_my - 1);
}

template <bool PROF>
__global__ void __launch_bounds__(WA_THREADS, 1)
window_attention_kernel(const __grid_constant__ WaMaps maps, bf16* __restrict__ out, const bf16* __restrict__ bias_dense,
                        StageGeom g, int n_seg, int C, int n_heads, int paired, int in_evict_first, float scale_log2e, long long* prof) {
  // NOTE: pointers must stay derived from the __shared__ array itself (no integer round trip), otherwise the compiler
  // falls back to generic LD/ST for every shared-memory access
  extern __shared__ __align__(1024) uint8_t smem[];
  WaShared sh;
  sh.smem = smem;
  sh.prof = prof;
  sh.qk_full = reinterpret_cast<uint64_t*>(smem + WA_OFF_BAR);  // [2]
  sh.qk_empty = sh.qk_full + 2;                                 // [2]
  sh.v_full = sh.qk_empty + 2;                                  // [2]
  sh.v_empty = sh.v_full + 2;                                   // [2]
  sh.s_full = sh.v_empty + 2;                                   // [2]
  sh.s_free = sh.s_full + 2;                                    // [2]
  sh.p_full = sh.s_free + 2;                                    // [4]
  sh.o_full = sh.p_full + 4;                                    // [4]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sh.o_full + 4);  // byte 160; watchdog flag at byte 184

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nwin = windows_per_segment(g);
  const int T = g.D * g.H * g.W;
  const int n_items = n_seg * nwin;
  const bool shifted = (g.sd | g.sh | g.sw) != 0;
  // work units are (head, segment, window) triples in head-major order; every CTA takes one contiguous, equally sized
  // range, so a CTA changes head (and reloads the 50 KB bias table) at most ceil(heads / CTAs) + 1 times.
  // paired: two heads share every 128-byte line of a q / k / v row (64 B per head), so CTAs 2k and 2k + 1 walk the SAME
  // (head pair, segment, window) range, one with the even and one with the odd head of each pair: the second half of every line
  // is an L2 hit a few hundred nanoseconds after the first instead of a second DRAM read tens of microseconds later (with plain
  // head-major ranges over 148 CTAs the two heads of a line drift apart and stages 2-4 read their input twice)
  const int n_groups = paired ? n_heads / 2 : n_heads, head_step = paired ? 2 : 1;
  const int worker = paired ? static_cast<int>(blockIdx.x >> 1) : static_cast<int>(blockIdx.x);
  const int n_workers = paired ? static_cast<int>(gridDim.x >> 1) : static_cast<int>(gridDim.x);
  const long long n_units = static_cast<long long>(n_groups) * n_items;
  const int u_lo = static_cast<int>(n_units * worker / n_workers);
  const int u_hi = static_cast<int>(n_units * (worker + 1) / n_workers);
  const int n_my = u_hi - u_lo;
  const int head0 = (u_lo / n_items) * head_step + (paired ? static_cast<int>(blockIdx.x & 1) : 0), item0 = u_lo % n_items;
  if (PROF && blockIdx.x == 0 && tid == 0) prof[24 * 8 + 2] = clock64();
  if ((smem_u32(smem) & 1023u) != 0) __trap();  // swizzled TMA / UMMA tiles assume an aligned window

  // ---- one-time setup: zero the staging and P buffers (pad slots stay zero forever), barriers, TMEM
  for (int i = tid; i < WA_OFF_BIAS / 16; i += WA_THREADS) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (tid < 64) reinterpret_cast<uint4*>(smem + WA_OFF_ONES)[tid] = make_uint4(0x3F803F80u, 0x3F803F80u, 0x3F803F80u, 0x3F803F80u);
  if (warp == WA_WARP_MMA0 && lane == 0) {
    *reinterpret_cast<volatile int*>(smem + WA_OFF_BAR + 184) = 0;  // watchdog abort flag (profiling hook only)
    for (int b = 0; b < 2; ++b) {
      mbar_init(&sh.qk_full[b], 1);
      mbar_init(&sh.qk_empty[b], 2);  // S0 (warp 21) and S1 (warp 22) both read q / k
      mbar_init(&sh.v_full[b], 1);
      mbar_init(&sh.v_empty[b], 2);
    }
    for (int t = 0; t < 2; ++t) {
      mbar_init(&sh.s_full[t], 1);
      mbar_init(&sh.s_free[t], t == 0 ? 16 : 4);
      for (int b = 0; b < 2; ++b) {
        mbar_init(&sh.p_full[2 * t + b], t == 0 ? 16 : 4);
        mbar_init(&sh.o_full[2 * t + b], 1);
      }
    }
    fence_barrier_init();
  }
  if (warp == WA_WARP_LOADER && lane == 0) {
#pragma unroll
    for (int k = 0; k < 4; ++k) tma_prefetch_desc(&maps.m[k]);
  }
  if (warp == WA_WARP_TMEM) tmem_alloc(tmem_slot, WA_TM_COLS);
  fence_proxy_async_smem();  // the fills above must be visible to the tensor core's operand reads and ordered before TMA writes
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // the set-up above (176 KB of shared-memory fills, barriers, TMEM) overlaps the tail of the previous kernel of the stream
  // (programmatic dependent launch); global memory is touched only below
  griddep_wait();
  griddep_launch();

  if (warp == WA_WARP_LOADER) {
    // ===================================================================== loader: 15 TMA boxes per unit
    if (lane == 0) {
      const int lw = 31 - __clz(g.W / 7);  // H/7 and W/7 are powers of two on this path (8, 4, 2, 1)
      const int nh_mask = g.H / 7 - 1, nw_mask = g.W / 7 - 1;
      const uint32_t smem0 = smem_u32(smem);
      const uint64_t pol_in = l2_policy(in_evict_first ? 1 : 0);
      // unit -> TMA coordinates: channel of the head's q slice, (x, y) of the parts before / behind the seam (only the
      // part behind the seam can wrap), first frame row of the segment
      struct Coord { int cq, xa, xb, ya, yb, ds; };
      int head = head0, item = item0;
      int seg = item / nwin, win = item - seg * nwin;
      auto coord_next = [&]() {
        Coord k;
        const WindowBoxes377 bx = window_boxes_377(g, (win >> lw) & nh_mask, win & nw_mask);  // pinned by tests/test_remap_cpu.py
        k.ya = bx.ya; k.xa = bx.xa; k.yb = bx.yb; k.xb = bx.xb;
        k.cq = head * 32; k.ds = seg * 3;
        if (++win == nwin) {
          win = 0;
          if (++seg * nwin == n_items) { seg = 0; head += head_step; }
        }
        return k;
      };
      auto load_qk = [&](int j, const Coord& k) {
        const int buf = j & 1;
        const uint32_t sq = smem0 + WA_OFF_STAGE + buf * WA_STAGE_BYTES, sk = sq + WA_Q_BYTES;
        uint64_t* bar = &sh.qk_full[buf];
        timed_wait<PROF>(sh, &sh.qk_empty[buf], ((j >> 1) & 1) ^ 1, 0, j);
        mbar_expect_tx(bar, WA_QK_TX_BYTES);
        tma_load_4d(sk + 0 * 64, &maps.m[0], bar, C + k.cq, k.xa, k.ya, k.ds, pol_in);
        tma_load_4d(sk + 48 * 64, &maps.m[1], bar, C + k.cq, k.xb, k.ya, k.ds, pol_in);
        tma_load_4d(sk + 88 * 64, &maps.m[2], bar, C + k.cq, k.xa, k.yb, k.ds, pol_in);
        tma_load_4d(sk + 128 * 64, &maps.m[3], bar, C + k.cq, k.xb, k.yb, k.ds, pol_in);
        tma_load_4d(sq + 0 * 64, &maps.m[0], bar, k.cq, k.xa, k.ya, k.ds, pol_in);
        tma_load_4d(sq + 48 * 64, &maps.m[1], bar, k.cq, k.xb, k.ya, k.ds, pol_in);
        tma_load_4d(sq + 88 * 64, &maps.m[2], bar, k.cq, k.xa, k.yb, k.ds, pol_in);
#pragma unroll
        for (int r = 0; r < 4; ++r) tma_load_4d(sq + (128 + 32 * r) * 64, &maps.m[3], bar, k.cq, k.xb, k.yb, k.ds, pol_in);
      };
      // q / k of unit j+1 are requested BEFORE v of unit j: their buffer is free as soon as S(j-1) has been issued, a whole
      // unit earlier than the v buffer (P v(j-2))
      Coord cur = coord_next();
      load_qk(0, cur);
      for (int j = 0; j < n_my; ++j) {
        Coord nxt = cur;
        if (j + 1 < n_my) {
          nxt = coord_next();
          load_qk(j + 1, nxt);
        }
        const int buf = j & 1;
        const uint32_t sv = smem0 + WA_OFF_STAGE + buf * WA_STAGE_BYTES + WA_Q_BYTES + WA_K_BYTES;
        uint64_t* bar = &sh.v_full[buf];
        timed_wait<PROF>(sh, &sh.v_empty[buf], ((j >> 1) & 1) ^ 1, 3, j);
        mbar_expect_tx(bar, WA_V_TX_BYTES);
        tma_load_4d(sv + 0 * 64, &maps.m[0], bar, 2 * C + cur.cq, cur.xa, cur.ya, cur.ds, pol_in);
        tma_load_4d(sv + 48 * 64, &maps.m[1], bar, 2 * C + cur.cq, cur.xb, cur.ya, cur.ds, pol_in);
        tma_load_4d(sv + 88 * 64, &maps.m[2], bar, 2 * C + cur.cq, cur.xa, cur.yb, cur.ds, pol_in);
        tma_load_4d(sv + 128 * 64, &maps.m[3], bar, 2 * C + cur.cq, cur.xb, cur.yb, cur.ds, pol_in);
        cur = nxt;
      }
    }
  } else if (warp == WA_WARP_MMA0 || warp == WA_WARP_MMA1) {
    // ===================================================================== MMA issuers: warp 21 -> row tile 0, 22 -> tile 1
    // the whole warp walks the loop and one elected lane issues: with the warp converged the descriptors live in uniform registers
    // and an MMA costs a few instructions (behind a lane-0 branch every operand is a per-MMA R2UR waterfall, ~60 cycles x 22 / unit)
    if (n_my > 0) {
      const int tile = warp == WA_WARP_MMA1;
      constexpr uint32_t idesc_s = umma_idesc_bf16(128, WA_KEYS);          // A, B K-major
      constexpr uint32_t idesc_o = umma_idesc_bf16(128, WA_ON) | (1u << 16);  // B (= [v | 1]) MN-major
      uint64_t* s_full = sh.s_full + tile;
      uint64_t* s_free = sh.s_free + tile;
      uint64_t* p_full = sh.p_full + 2 * tile;
      uint64_t* o_full = sh.o_full + 2 * tile;
      const uint32_t tm_s = tmem_base + (tile ? WA_TM_S1 : WA_TM_S0), tm_o = tmem_base + (tile ? WA_TM_O1 : WA_TM_O0);
      // Descriptors of one operand differ only in their start-address field (bits 0..13, in 16-byte units), so every
      // further MMA of a sequence costs one 64-bit add instead of a fresh encode.
      const uint32_t smem0 = smem_u32(smem);
      auto issue_s = [&](int j) {
        const uint32_t b = smem0 + WA_OFF_STAGE + (j & 1) * WA_STAGE_BYTES;
        const uint64_t dq = umma_desc_sw64(b + tile * (128 * 64)), dk = umma_desc_sw64(b + WA_Q_BYTES);
        if (elect_one()) {
#pragma unroll
          for (int kk = 0; kk < 2; ++kk) umma_bf16_ss(tm_s, dq + ((kk * 32) >> 4), dk + ((kk * 32) >> 4), idesc_s, kk);
          umma_commit(s_full);
          umma_commit(&sh.qk_empty[j & 1]);
        }
        __syncwarp();
      };
      timed_wait<PROF>(sh, &sh.qk_full[0], 0, 1, 0);
      tcgen05_fence_after();
      issue_s(0);
      for (int j = 0; j < n_my; ++j) {
        if (j + 1 < n_my) {
          timed_wait<PROF>(sh, s_free, j & 1, 0, j);  // every softmax warp of this row tile has S(j) in registers
          timed_wait<PROF>(sh, &sh.qk_full[(j + 1) & 1], ((j + 1) >> 1) & 1, 1, j);
          tcgen05_fence_after();
          issue_s(j + 1);
        }
        timed_wait<PROF>(sh, p_full + (j & 1), (j >> 1) & 1, 2, j);  // P(j) in smem, output buffer j & 1 drained (unit j-2)
        timed_wait<PROF>(sh, &sh.v_full[j & 1], (j >> 1) & 1, 3, j);
        tcgen05_fence_after();
        // B = [v | 1]: N = 48 = the 32 dims of v (one 64-byte swizzle atom per key) + a second atom of ones (columns 32..47 = the row
        // sums L = P 1 of the bf16 probabilities, for free with the same A reads). The ones are ONE 16-key slab: every K step
        // moves the start address 1024 B ahead and the atom stride (lbo) 1024 B back, so the second atom never moves.
        const uint32_t v_addr = smem0 + WA_OFF_STAGE + (j & 1) * WA_STAGE_BYTES + WA_Q_BYTES + WA_K_BYTES;
        const uint64_t dv = umma_desc_sw64_lbo(v_addr, smem0 + WA_OFF_ONES - v_addr);
        const uint64_t dp = umma_desc_nosw(smem0 + (tile ? WA_OFF_P1 + (j & 1) * WA_P1_BYTES : WA_OFF_P0 + (j & 1) * WA_P0_BYTES),
                                           128, (WA_KEYS / 8) * 128);
        const uint32_t tm_oj = tm_o + (j & 1) * WA_ON;
        if (elect_one()) {
#pragma unroll
          for (int kk = 0; kk < WA_KEYS / 16; ++kk)  // 16 keys per step: 2 cores of P (256 B), 2 row groups of v (1024 B)
            umma_bf16_ss(tm_oj, dp + ((kk * 256) >> 4), dv + ((kk * 1024) >> 4) - (static_cast<uint64_t>((kk * 1024) >> 4) << 16), idesc_o,
                         kk);
          umma_commit(&sh.v_empty[j & 1]);
          umma_commit(o_full + (j & 1));
        }
        __syncwarp();
      }
    }
  } else if (warp < 20) {
    // ===================================================================== softmax + epilogue (20 warps)
    WaItemCtx cx;
    cx.bias_dense = bias_dense; cx.out = out; cx.g = g; cx.n_items = n_items; cx.nwin = nwin; cx.T = T; cx.C = C;
    cx.head0 = head0; cx.item0 = item0; cx.head_step = head_step; cx.n_my = n_my; cx.scale_log2e = scale_log2e; cx.shifted = shifted;
    if (warp < 16) softmax_warp<0, PROF>(sh, tmem_base, warp & 3, warp >> 2, cx);
    else softmax_warp<1, PROF>(sh, tmem_base, warp & 3, warp & 3, cx);
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == WA_WARP_TMEM) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, WA_TM_COLS);
  }
  if (PROF && blockIdx.x == 0 && tid == 0) {
    prof[24 * 8] = n_my;
    prof[24 * 8 + 1] = clock64() - prof[24 * 8 + 2];
  }
}

// dense[h][slot_i][slot_j] = table[rel_index(tok_i, tok_j)][h] * log2(e) for rows 0..154 (pad key columns -inf, pad rows 0),
// the 16-byte chunks of row i stored at chunk index (j / 8) ^ ((i >> 1) & 3) (bank-conflict-free row-per-lane reads);
// row 155 of every head holds bf16 bmax[160] = max_j dense[h][slot_i][.] (0 for pad rows; exact: a maximum of bf16 values)
__global__ void build_dense_bias_kernel(const float* __restrict__ table, bf16* __restrict__ dense, StageGeom g,
                                        int n_heads) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n_heads * WA_KEYS) return;
  const int si = idx % WA_KEYS, h = idx / WA_KEYS;
  bf16* head = dense + static_cast<size_t>(h) * WA_KEYS * WA_BIAS_PITCH;
  bf16* bmax = head + WA_BIAS_ROWS * WA_BIAS_PITCH;
  const int ti = slot_token_377(si);
  float mx = -INFINITY;
  for (int sj = 0; sj < WA_KEYS; ++sj) {
    const int tj = slot_token_377(sj);
    bf16 v = __float2bfloat16(-INFINITY);
    if (tj >= 0) {
      if (ti >= 0) {
        const int rel = rel_pos_offset(g, ti) - rel_pos_offset(g, tj) + REL_POS_CENTER;
        v = __float2bfloat16(table[static_cast<size_t>(rel) * n_heads + h] * 1.4426950408889634f);
      } else {
        v = __float2bfloat16(0.f);
      }
      mx = fmaxf(mx, __bfloat162float(v));
    }
    if (si < WA_BIAS_ROWS) head[si * WA_BIAS_PITCH + (((sj >> 3) ^ ((si >> 1) & 3)) << 3) + (sj & 7)] = v;
  }
  bmax[si] = __float2bfloat16(mx);
}

}  // namespace lrce

using namespace lrce;

// Instrumented instantiation (lrce_window_attention_profile, tools only; a per-call argument, the library keeps no profiling
// state): CTA 0 accumulates into buf (device, 224 zeroed int64), per warp w at buf[8 w ..], the cycles spent in each kind of mbarrier wait
// (softmax warps: [0] S ready, [1] P v done, [4] TMEM load, [5] maximum, [6] probabilities, [7] epilogue; loader: [0] q/k
// buffer free, [3] v buffer free; MMA: [0] S released, [1] q/k landed, [2] P ready, [3] v landed), plus buf[192] units and
// buf[193] total cycles of CTA 0; buf[196 + w] != 0 reports a watchdog hit of warp w.
static int geom_3x7x7(StageGeom* g, int D, int H, int W, int sh, int sw) {
  LRCE_REQUIRE(D == 3 && H % 7 == 0 && W % 7 == 0 && H > 0 && W > 0,
               "window attention is specialised for the clamped (3,7,7) window of LRCE's 5-frame segments; got grid "
               "(%d,%d,%d)", D, H, W);
  LRCE_REQUIRE(((H / 7) & (H / 7 - 1)) == 0 && ((W / 7) & (W / 7 - 1)) == 0,
               "window attention expects a power-of-two number of windows per axis; got grid (%d,%d,%d)", D, H, W);
  LRCE_REQUIRE((sh == 0 || sh == 3) && (sw == 0 || sw == 3),
               "window attention is specialised for Swin's shift = window // 2 = 3 (or 0); got (%d,%d)", sh, sw);
  g->D = D; g->H = H; g->W = W; g->wd = 3; g->wh = 7; g->ww = 7; g->sd = 0; g->sh = sh; g->sw = sw;
  return LRCE_OK;
}

// tensor maps of the four class boxes over qkv viewed as (channel 3C, w, h, frame*segment); cached per (pointer, geometry)
struct WaMapEntry {
  const void* qkv;
  int dev, n_seg, H, W, C;
  WaMaps maps;
};
static int get_maps(const WaMaps** out, const void* qkv, int n_seg, int H, int W, int C) {
  static thread_local WaMapEntry cache[16];
  static thread_local int n_cached = 0, next = 0;
  const int dev = current_device();  // the same virtual address can name different memory on another device
  for (int i = 0; i < n_cached; ++i) {
    const WaMapEntry& e = cache[i];
    if (e.qkv == qkv && e.dev == dev && e.n_seg == n_seg && e.H == H && e.W == W && e.C == C) {
      *out = &e.maps;
      return LRCE_OK;
    }
  }
  WaMapEntry& e = cache[next];
  const uint64_t row = static_cast<uint64_t>(3 * C) * 2;
  const uint64_t dims[4] = {static_cast<uint64_t>(3 * C), static_cast<uint64_t>(W), static_cast<uint64_t>(H),
                            static_cast<uint64_t>(3) * n_seg};
  const uint64_t strides[3] = {row, row * W, row * W * H};
  for (int k = 0; k < 4; ++k) {
    const uint32_t box[4] = {32, (k & 1) ? 3u : 4u, (k & 2) ? 3u : 4u, 3};
    const int rc = make_tmap_nd_bf16(&e.maps.m[k], qkv, 4, dims, strides, box, 64, 128);
    if (rc != LRCE_OK) return rc;
  }
  e.qkv = qkv; e.dev = dev; e.n_seg = n_seg; e.H = H; e.W = W; e.C = C;
  *out = &e.maps;
  next = (next + 1) % 16;
  if (n_cached < 16) ++n_cached;
  return LRCE_OK;
}

static int window_attention_launch(const void* qkv, void* out, const void* bias_dense, int n_seg, int D, int H, int W, int C,
                                   int n_heads, int shift_h, int shift_w, void* stream, long long* prof) {
  int rc = require_sm100();
  if (rc != LRCE_OK) return rc;
  StageGeom g;
  rc = geom_3x7x7(&g, D, H, W, shift_h, shift_w);
  if (rc != LRCE_OK) return rc;
  LRCE_REQUIRE(qkv && out && bias_dense && n_seg > 0, "lrce_window_attention_bf16: null operand");
  LRCE_REQUIRE(n_heads > 0 && C == n_heads * 32, "lrce_window_attention_bf16: head_dim must be 32 (C=%d heads=%d)", C, n_heads);
  static thread_local uint64_t configured = 0;  // one bit per device
  if (needs_device_setup(&configured)) {
    cudaError_t e = cudaFuncSetAttribute(window_attention_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, WA_SMEM);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(window_attention_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, WA_SMEM);
    if (e != cudaSuccess) {
      set_error("cudaFuncSetAttribute(window_attention_kernel): %s", cudaGetErrorString(e));
      return LRCE_ECUDA;
    }
    mark_device_setup(&configured);
  }
  const WaMaps* maps = nullptr;
  rc = get_maps(&maps, qkv, n_seg, H, W, C);
  if (rc != LRCE_OK) return rc;
  const long long n_units = static_cast<long long>(n_seg) * windows_per_segment(g) * n_heads;
  int grid = sm_count();  // one persistent CTA per SM (it owns all 512 TMEM columns)
  // pairs of CTAs walk the two heads of a 128-byte q / k / v line together (see the kernel); LRCE_B200_ATTN_ALIGN=0: plain
  // head-major ranges (A/B runs of tools/)
  static const bool align_heads = [] {
    const char* e = getenv("LRCE_B200_ATTN_ALIGN");
    return !(e && e[0] == '0');
  }();
  const int paired = (align_heads && n_heads % 2 == 0 && grid >= 2) ? 1 : 0;
  static const int in_hint = [] {
    const char* e = getenv("LRCE_B200_ATTN_IN_HINT");  // A/B runs of tools/
    return e ? atoi(e) : 1;
  }();
  if (paired) grid &= ~1;
  if (grid > n_units) grid = static_cast<int>(paired ? (n_units & ~1LL) : n_units);
  const float scale_log2e = 0.17677669529663687f * 1.4426950408889634f;  // 32^-0.5 * log2(e)
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  cudaError_t e;
  if (prof != nullptr)  // instrumented instantiation
    e = launch_pdl(window_attention_kernel<true>, dim3(grid), dim3(WA_THREADS), WA_SMEM, s, *maps, reinterpret_cast<bf16*>(out),
                   reinterpret_cast<const bf16*>(bias_dense), g, n_seg, C, n_heads, paired, in_hint, scale_log2e, prof);
  else
    e = launch_pdl(window_attention_kernel<false>, dim3(grid), dim3(WA_THREADS), WA_SMEM, s, *maps, reinterpret_cast<bf16*>(out),
                   reinterpret_cast<const bf16*>(bias_dense), g, n_seg, C, n_heads, paired, in_hint, scale_log2e, static_cast<long long*>(nullptr));
  if (e != cudaSuccess) {
    set_error("cudaLaunchKernelEx(window_attention_kernel): %s", cudaGetErrorString(e));
    return LRCE_ECUDA;
  }
  return check_launch("window_attention_kernel");
}

extern "C" int lrce_window_attention_bf16(const void* qkv, void* out, const void* bias_dense, int n_seg, int D, int H,
                                          int W, int C, int n_heads, int shift_h, int shift_w, void* stream) {
  return window_attention_launch(qkv, out, bias_dense, n_seg, D, H, W, C, n_heads, shift_h, shift_w, stream, nullptr);
}

extern "C" int lrce_window_attention_profile(const void* qkv, void* out, const void* bias_dense, int n_seg, int D, int H,
                                             int W, int C, int n_heads, int shift_h, int shift_w, void* stream, long long* prof) {
  LRCE_REQUIRE(prof != nullptr, "lrce_window_attention_profile: prof buffer required");
  return window_attention_launch(qkv, out, bias_dense, n_seg, D, H, W, C, n_heads, shift_h, shift_w, stream, prof);
}

extern "C" int lrce_window_bias_pack(const float* table, void* bias_dense, int n_heads, void* stream) {
  int rc = require_sm100();
  if (rc != LRCE_OK) return rc;
  LRCE_REQUIRE(table && bias_dense && n_heads > 0, "lrce_window_bias_pack: bad arguments");
  StageGeom g;
  g.D = 3; g.H = 7; g.W = 7; g.wd = 3; g.wh = 7; g.ww = 7; g.sd = g.sh = g.sw = 0;
  const int total = n_heads * WA_KEYS;
  build_dense_bias_kernel<<<(total + 127) / 128, 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      table, reinterpret_cast<bf16*>(bias_dense), g, n_heads);
  return check_launch("build_dense_bias_kernel");
}
