// encoder_walk.cu — the summarisation token's walk through the recurrent cross-modal encoder (fusionv3.py:41-51
// FusionTransformer.forward + :195 final_fc, nn.TransformerDecoderLayer post-norm) as ONE kernel of row-sharded 16-CTA
// clusters: TMA-streamed weights, swap-AB tcgen05 products, cluster-local exchanges. No grid barrier exists.
//
// The walk is S segments x 12 layers of six dependent sub-steps on a (rows <= 160) x 768 state:
//   P1  y1 = x + W_sa x + b_sa           length-1 self-attention == out_proj(v_proj(x)), one folded matrix
//   P2  h1 = LN1(y1) ; q = W_q h1 + b_q  (1/8 scale folded into W_q)
//   P3  ctx = softmax(q K^T) V           per (row, head) over [video segment s ; text] from the precomputed K/V GEMM
//   P4  y2 = h1 + W_o ctx + b_o
//   P5  h2 = LN2(y2) ; hdn = gelu(W_1 h2 + b_1)
//   P6  y3 = h2 + W_2 hdn + b_2 ; x' = LN3(y3)          after layer 12: tok' = LN_f(tok + x'), then the next segment
//   end logits = act(W_fc tok_S + b_fc)
// Every row (clip / candidate) walks independently of every other row, so rows — not output features — are what is
// sharded across the chip: a cluster of 16 CTAs owns <= 8 rows for the whole walk and never talks to another cluster.
// Inside a cluster the OUTPUT FEATURES of every product are split 16 ways (48 of 768, 192 of 3072), which makes the
// weights the M operand of the MMA and the <= 8 token rows its N operand (swap-AB):
//   * weights: lrce_encoder_walk_pack re-tiles them once per weight version into the exact order CTA `rank` consumes them
//     (24 KB ring slots = 192 rows of 64 bf16; LayerNorm gamma/beta of the preceding norm folded into W and b), so the
//     whole walk is ONE linear TMA stream per CTA (cp.async.bulk.tensor, 128B swizzle) that never waits for the token
//     state; K / V head slices of the CTA's (row, head) attention units travel through the same ring in consumption order;
//   * products: one thread issues tcgen05.mma (M = 64, N = 8, bf16, fp32 accumulators in TMEM) straight on the
//     TMA-delivered tiles; fc2 is split-K over the 192 hidden features the CTA itself just produced (no exchange of the
//     hidden state), followed by a reduce-scatter of fp32 partials;
//   * exchanges (6 per layer-step): every CTA sends its slice to all 16 CTAs with cp.async.bulk shared::cta ->
//     shared::cluster copies that complete_tx on the RECEIVER's mbarrier — one-sided, no flags, no cluster barrier on the
//     path; LayerNorm is then recomputed redundantly by every CTA (8 warps, one row each).
// What bounds it (tools/probes/walk_probe.cu, profiles/): per-SM L2 -> shared-memory ingest of the weight stream
// (0.9 MB per layer-step and CTA at ~150 GB/s) next to a dependency chain of about the same length.
#include "encoder_common.cuh"
#include "host_common.h"

namespace lrce {

struct EncLayerF32 {  // pack input: device pointers of one decoder layer, fp32 master weights (16 pointers per layer)
  const float *sa_w, *q_w, *o_w, *w1, *w2;
  const float *sa_b, *q_b, *o_b, *b1, *b2, *n1g, *n1b, *n2g, *n2b, *n3g, *n3b;
};
static_assert(sizeof(EncLayerF32) == 16 * sizeof(void*), "layer table layout is part of the C ABI (16 pointers per layer)");

constexpr int WK_CL = 16;                  // CTAs per cluster
constexpr int WK_F = ENC_D / WK_CL;        // 48 output features of a 768-wide product per CTA
constexpr int WK_H = 4 * ENC_D / WK_CL;    // 192 hidden features per CTA
constexpr int WK_NPAD = 8;                 // token rows per cluster = N of the MMA
constexpr int WK_SLOT_ROWS = 192;          // ring slot: 192 rows of 64 bf16
constexpr int WK_SLOT_BYTES = WK_SLOT_ROWS * 128;
constexpr int WK_W_ROWS = 3 * 12 * WK_F + 12 * WK_H + 12 * 3 * 64;  // stream rows per (layer, rank): 1728 + 2304 + 2304
constexpr int WK_W_SLOTS = WK_W_ROWS / WK_SLOT_ROWS;                // 33
constexpr int WK_PB = 10 * WK_F + WK_H;    // floats per (layer, rank) parameter block
enum { PB_SA_B = 0, PB_Q_B = 48, PB_O_B = 96, PB_B2 = 144, PB_G1 = 192, PB_BE1 = 240, PB_G2 = 288, PB_BE2 = 336, PB_G3 = 384,
       PB_BE3 = 432, PB_B1 = 480 };
constexpr int WK_THREADS = 384;            // warps 0-7 compute, 8 TMA producer, 9 and 11 MMA issuers, 10 TMEM allocator
constexpr int WK_CWARPS = 8;
constexpr int WK_CTHREADS = WK_CWARPS * 32;
constexpr int WK_MAX_KEYS = 192;           // one K (or V) head slice of a unit = one ring slot
constexpr int WK_MAX_SLOTS = 8;
constexpr int WK_ATT_FLOATS = WK_CWARPS * 64 + 2 * WK_CWARPS + WK_CWARPS * 24;
constexpr int WK_TMEM_COLS = 256;
// accumulator columns (8 per 64-row tile); products whose K runs over several slots keep one partial sum per MMA issuer
constexpr int WK_TM_SA = 0, WK_TM_Q = 16, WK_TM_O = 32, WK_TM_FC1 = 48, WK_TM_FC2 = 96, WK_TM_HEAD = 192;
constexpr int WK_MAX_HEAD_TILES = 4;       // 64-feature head tiles per CTA: n_out <= 4096
constexpr int WK_SMEM_MAX = 232448;
static_assert(WK_W_ROWS % WK_SLOT_ROWS == 0 && WK_PB * 4 % 16 == 0, "stream layout");
enum { ACT_NONE = 0, ACT_GELU = 1, ACT_RELU = 2 };

__host__ __device__ inline int walk_head_tiles(int n_out) { return ((n_out + WK_CL - 1) / WK_CL + 63) / 64; }

// packed buffer (bytes): [weight stream][head stream][parameter blocks][gamma3 / beta3 of the last layer][head bias]
struct WalkPackLayout {
  size_t off_head, off_params, off_tail, off_hbias, total;
  long long w_rows, head_rows;
};
__host__ __device__ inline WalkPackLayout walk_pack_layout(int n_layers, int n_out) {
  WalkPackLayout l;
  const int mt = walk_head_tiles(n_out);
  l.w_rows = static_cast<long long>(n_layers) * WK_CL * WK_W_ROWS;
  l.head_rows = static_cast<long long>(WK_CL) * mt * 12 * 64;
  l.off_head = static_cast<size_t>(l.w_rows) * 128;
  l.off_params = l.off_head + static_cast<size_t>(l.head_rows) * 128;
  l.off_tail = l.off_params + static_cast<size_t>(n_layers) * WK_CL * WK_PB * 4;
  l.off_hbias = l.off_tail + 2 * ENC_D * 4;
  l.total = l.off_hbias + static_cast<size_t>(WK_CL) * 64 * mt * 4;
  return l;
}

// shared-memory map of the walk kernel for `rpc` rows per cluster and `ns` ring slots (same function on host and device)
struct WalkSmem {
  int xb, hdn, G, tok, res, stg, ctx, par, att, bar, total;
};
__host__ __device__ inline WalkSmem walk_smem(int rpc, int ns) {
  WalkSmem m;
  m.xb = ns * WK_SLOT_BYTES + 2048;          // ring + slack: an M = 64 read of the last 48-row tile of a slot runs 2 KB past it
  m.hdn = m.xb + WK_NPAD * ENC_D * 2;        // xb: bf16 [8][768] B operand (128B-swizzled k-blocks of 1 KB)
  m.G = m.hdn + WK_NPAD * WK_H * 2;          // hdn: bf16 [8][192] B operand of the split-K fc2
  m.tok = m.G + 2 * WK_CL * rpc * WK_F * 4;  // G: fp32 [2][16 sources][rpc][48] exchange buffers
  m.res = m.tok + rpc * ENC_D * 4;           // tok: fp32 [rpc][768] token at the start of the segment
  m.stg = m.res + rpc * WK_F * 4;            // res: fp32 [rpc][48] residual slice of the running sub-step
  m.ctx = m.stg + 2 * rpc * WK_F * 4;        // stg: fp32 [2][rpc][48] outgoing slices
  m.par = m.ctx + 2048;                      // ctx: outgoing attention outputs, an image of this CTA's address range of xb
  m.att = m.par + 2 * WK_PB * 4;             // par: fp32 [2][672] parameter blocks
  m.bar = m.att + 2 * WK_ATT_FLOATS * 4;     // att: per-warp attention partials (outputs, maxima, sums, probabilities), double-buffered
  m.total = m.bar + 512;  // barriers, TMEM slot; the upper 256 bytes are the profile build's cycle buckets
  return m;
}

struct WalkParams {
  const float* params;     // [n_layers][16][WK_PB]
  const float* tail;       // [2][768]: gamma3, beta3 of the last layer
  const float* head_bias;  // [16 * 64 * mt]
  const float *tok0, *f_g, *f_b;
  float* out;         // [R, n_out]
  float* tokens_tap;  // nullptr or [S, R, 768]: the token after every segment (tests)
  int n_layers, n_out, act, R, S, Tv, Lt, n_cand, rpc, n_groups, ns, mt_head;
  int head_row0;  // first row of the head stream inside the packed weight tensor
  float eps;
  long long* prof;  // lrce_encoder_walk_profile only: [CTA][32] cycles per sub-step of the compute warps' chain
};

// ------------------------------------------------------------------------------------------------------------------
// small device helpers
// ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
// shared::cta -> shared::cluster bulk copy; completes `bytes` on the mbarrier at `bar_cluster` (an address in the
// destination CTA's window)
__device__ __forceinline__ void bulk_s2c(uint32_t dst_cluster, uint32_t src_cta, uint32_t bytes, uint32_t bar_cluster) {
  asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_cluster),
               "r"(src_cta), "r"(bytes), "r"(bar_cluster)
               : "memory");
}
// plain (non-tensor) global -> shared bulk copy
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
               "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tma_prefetch_l2_2d(const CUtensorMap* tm, int c0, int c1) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global [%0, {%1, %2}];" ::"l"(reinterpret_cast<uint64_t>(tm)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tmem_ld_8(uint32_t taddr, uint32_t (&v)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(taddr)
               : "memory");
}
// position in the slot ring: slot index and the parity of its current use (no divisions on the per-slot paths)
struct RingPos {
  uint32_t s, ph;
};
__device__ __forceinline__ void ring_adv(RingPos& r, uint32_t n, uint32_t ns) {
  r.s += n;
  while (r.s >= ns) {
    r.s -= ns;
    r.ph ^= 1;
  }
}
__device__ __forceinline__ void cbar_sync() { asm volatile("bar.sync 1, %0;" ::"n"(WK_CTHREADS) : "memory"); }
__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// ------------------------------------------------------------------------------------------------------------------
// the kernel
// ------------------------------------------------------------------------------------------------------------------
template <bool PROF, int ATTN>
__global__ void __launch_bounds__(WK_THREADS, 1)
encoder_walk_kernel(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmKVv,
                    const __grid_constant__ CUtensorMap tmKVt, const WalkParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int rank = static_cast<int>(cluster_ctarank());
  const int cluster_id = blockIdx.x / WK_CL, n_clusters = gridDim.x / WK_CL;
  const int rpc = p.rpc, ns = p.ns;
  const WalkSmem sm = walk_smem(rpc, ns);
  uint8_t* ring = smem;
  uint8_t* xb = smem + sm.xb;
  uint8_t* hdn = smem + sm.hdn;
  float* G = reinterpret_cast<float*>(smem + sm.G);
  float* tok = reinterpret_cast<float*>(smem + sm.tok);
  float* res = reinterpret_cast<float*>(smem + sm.res);
  float* stg = reinterpret_cast<float*>(smem + sm.stg);
  uint8_t* ctx_stg = smem + sm.ctx;
  float* par_base = reinterpret_cast<float*>(smem + sm.par);
  float* att = reinterpret_cast<float*>(smem + sm.att);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + sm.bar);  // [8]
  uint64_t* empty = full + WK_MAX_SLOTS;                        // [8]
  uint64_t* acc_full = empty + WK_MAX_SLOTS;
  uint64_t* b_ready = acc_full + 1;
  uint64_t* gbar = b_ready + 1;    // [2]
  uint64_t* ctx_bar = gbar + 2;
  uint64_t* par_bar = ctx_bar + 1;  // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(par_bar + 2);

  if ((smem_u32(smem) & 1023u) != 0) __trap();  // swizzled TMA / UMMA tiles assume an aligned window
  // ---- one-time setup
  for (int i = tid; i < (sm.G - sm.xb) / 16; i += WK_THREADS) reinterpret_cast<uint4*>(xb)[i] = make_uint4(0, 0, 0, 0);
  for (int i = tid; i < 2048 / 16; i += WK_THREADS) {
    reinterpret_cast<uint4*>(ring + ns * WK_SLOT_BYTES)[i] = make_uint4(0, 0, 0, 0);
    reinterpret_cast<uint4*>(ctx_stg)[i] = make_uint4(0, 0, 0, 0);
  }
  if (tid == 0) {
    for (int s = 0; s < WK_MAX_SLOTS; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(acc_full, 2);  // one commit per MMA issuer
    mbar_init(b_ready, 1);
    mbar_init(&gbar[0], 1);
    mbar_init(&gbar[1], 1);
    mbar_init(ctx_bar, 1);
    mbar_init(&par_bar[0], 1);
    mbar_init(&par_bar[1], 1);
    fence_barrier_init();
  }
  if (warp == 8 && lane == 0) {
    tma_prefetch_desc(&tmW);
    tma_prefetch_desc(&tmKVv);
    tma_prefetch_desc(&tmKVt);
  }
  if (warp == 10) tmem_alloc(tmem_slot, WK_TMEM_COLS);
  fence_proxy_async_smem();
  tcgen05_fence_before();
  cluster_sync_all();  // every CTA's barriers exist before the first remote copy can target them
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  griddep_wait();  // programmatic dependent launch: the set-up above overlaps the previous kernel's tail, global memory only below
  griddep_launch();

  const int n_keys = p.Tv + p.Lt;
  // (row, head) attention units in head-major order u = head * rpc + row: unit u's output is the 128-byte row at byte
  // (u / rpc) * 1024 + (u % rpc) * 128 of xb (k-block = head). CTA `rank` owns the contiguous units [u0, u0 + nu), so its
  // outputs are ONE address range of xb (the pad rows between two heads travel along as zeros).
  const int n_units = rpc * 12;
  const int nu = n_units / WK_CL + (rank < n_units % WK_CL ? 1 : 0);
  const int u0 = rank * (n_units / WK_CL) + min(rank, n_units % WK_CL);
  auto unit_off = [&](int u) { return (u / rpc) * 1024 + (u % rpc) * 128; };
  const int ctx_off = unit_off(u0), ctx_bytes = nu > 0 ? unit_off(u0 + nu - 1) + 128 - ctx_off : 0;
  uint32_t ctx_total = 0;  // bytes every CTA receives per layer-step
  for (int r = 0; r < WK_CL; ++r) {
    const int nr = n_units / WK_CL + (r < n_units % WK_CL ? 1 : 0), ur = r * (n_units / WK_CL) + min(r, n_units % WK_CL);
    if (nr > 0) ctx_total += unit_off(ur + nr - 1) + 128 - unit_off(ur);
  }
  const int L = p.n_layers, S = p.S;
  const int head_slots = 4 * p.mt_head;

  if (warp == 8) {
    // =================================================================== TMA producer: one linear stream per CTA
    if (lane == 0) {
      uint32_t k = 0;
      RingPos rp = {0, 0};
      auto acquire = [&]() {
        const uint32_t s = rp.s;
        mbar_wait_short(&empty[s], rp.ph ^ 1);
        ring_adv(rp, 1, ns);
        return s;
      };
      auto load_params = [&](uint32_t step, int layer) {
        mbar_expect_tx(&par_bar[step & 1], WK_PB * 4);
        bulk_g2s(smem_u32(par_base + (step & 1) * WK_PB), p.params + (static_cast<size_t>(layer) * WK_CL + rank) * WK_PB,
                 WK_PB * 4, &par_bar[step & 1]);
      };
      load_params(0, 0);
      for (int g = cluster_id; g < p.n_groups; g += n_clusters) {
        const int r0 = g * rpc;
        for (int s = 0; s < S; ++s) {
          for (int n = 0; n < L; ++n, ++k) {
            const int w_row0 = (n * WK_CL + rank) * WK_W_ROWS;
            int wj = 0;
            auto wslots = [&](int count) {
              for (int c = 0; c < count; ++c, ++wj) {
                const uint32_t sl = acquire();
                mbar_expect_tx(&full[sl], WK_SLOT_BYTES);
                tma_load_2d(ring + sl * WK_SLOT_BYTES, &tmW, &full[sl], 0, w_row0 + wj * WK_SLOT_ROWS);
                // Every cluster streams the same bytes at about the same time, so every load would otherwise see HBM latency:
                // the clusters take turns pulling the NEXT layer-step's slots into L2 a whole layer-step ahead.
                if ((k * WK_W_SLOTS + wj) % n_clusters == static_cast<uint32_t>(cluster_id))
                  tma_prefetch_l2_2d(&tmW, 0, (((n + 1) % L) * WK_CL + rank) * WK_W_ROWS + wj * WK_SLOT_ROWS);
              }
            };
            wslots(6);  // sa, q
            for (int ui = 0; ui < nu; ++ui) {
              const int u = u0 + ui, head = u / rpc, lr = u - head * rpc;
              const int b = min(r0 + lr, p.R - 1);
              const int col = n * 2 * ENC_D + head * 64;
              const int vrow = ((b / p.n_cand) * S + s) * p.Tv, trow = b * p.Lt;
#pragma unroll
              for (int part = 0; part < 2; ++part) {  // K, then V: one ring slot each
                const uint32_t sl = acquire();
                mbar_expect_tx(&full[sl], n_keys * 128);
                tma_load_2d(ring + sl * WK_SLOT_BYTES, &tmKVv, &full[sl], col + part * ENC_D, vrow);
                tma_load_2d(ring + sl * WK_SLOT_BYTES + p.Tv * 128, &tmKVt, &full[sl], col + part * ENC_D, trow);
              }
            }
            wslots(3);  // o
            // The parameter block of the NEXT layer-step goes into the buffer layer-step k-1 used. The slot just acquired
            // was released by a consumer of layer-step k, so the compute warps have left layer-step k-1 for good.
            {
              const bool last = (n + 1 == L) && (s + 1 == S) && (g + n_clusters >= p.n_groups);
              if (!last) load_params(k + 1, (n + 1) % L);
            }
            wslots(24);  // fc1, fc2
          }
        }
        for (int j = 0; j < head_slots; ++j) {
          const uint32_t sl = acquire();
          mbar_expect_tx(&full[sl], WK_SLOT_BYTES);
          tma_load_2d(ring + sl * WK_SLOT_BYTES, &tmW, &full[sl], 0, p.head_row0 + rank * p.mt_head * 768 + j * WK_SLOT_ROWS);
        }
      }
    }
  } else if (warp == 9 || warp == 11) {
    // =================================================================== two MMA issuers (each warp converged, one elected lane)
    // The issuers take alternate weight slots of every product, each into its own accumulators (the epilogues add the two
    // partial sums): one warp's barrier wait / descriptor setup hides behind the other's MMAs on the shared tensor pipe.
    const int mw = warp == 9 ? 0 : 1;
    constexpr uint32_t idesc = umma_idesc_bf16(64, WK_NPAD);
    const uint32_t ring_lo = desc_lo(smem_u32(ring)), xb_lo = desc_lo(smem_u32(xb)), hdn_lo = desc_lo(smem_u32(hdn));
    uint32_t br = 0, kstep = 0;
    RingPos rp = {0, 0};
    long long mwait[4] = {0, 0, 0, 0};  // PROF: cycles blocked on the weight stream in [lin48, fc1, fc2] phases, [3] = waiting for B
    int mph = 0;
    // next ring slot: returns its index; waits for the data only if this warp is the one that consumes it
    auto next_slot = [&](bool mine) {
      const uint32_t s = rp.s;
      if (mine) {
        const long long t0 = PROF ? clock64() : 0;
        mbar_wait_short(&full[s], rp.ph);
        if (PROF) mwait[mph] += clock64() - t0;
        tcgen05_fence_after();
      }
      ring_adv(rp, 1, ns);
      return s;
    };
    auto wait_b = [&]() {
      const long long t0 = PROF ? clock64() : 0;
      mbar_wait_short(b_ready, br & 1);
      if (PROF) mwait[3] += clock64() - t0;
      ++br;
      tcgen05_fence_after();
    };
    auto phase_done = [&]() {
      if (elect_one()) umma_commit(acc_full);
      __syncwarp();
    };
    // 48 output features x K = 768: 3 slots of four 48-row k-block tiles; partial sums at tm_col + 8 mw
    auto lin48 = [&](uint32_t tm_col) {
#pragma unroll 1
      for (int j = 0; j < 3; ++j) {
        const bool mine = (j & 1) == mw;
        const uint32_t s = next_slot(mine);
        if (!mine) continue;
        const uint32_t a_lo = ring_lo + s * (WK_SLOT_BYTES >> 4), b_lo = xb_lo + j * (4 * 1024 >> 4);
        const uint32_t tm_d = tmem_base + tm_col + mw * WK_NPAD;
        if (elect_one()) {
          if (j < 2) umma_lo<false>(tm_d, a_lo, b_lo, idesc);
          else umma_lo<true>(tm_d, a_lo, b_lo, idesc);
#pragma unroll
          for (int m = 1; m < 16; ++m)  // m = 4 kbl + ks
            umma_lo<true>(tm_d, a_lo + (((m >> 2) * (WK_F * 128) + (m & 3) * 32) >> 4), b_lo + (((m >> 2) * 1024 + (m & 3) * 32) >> 4),
                          idesc);
          umma_commit(&empty[s]);
        }
        __syncwarp();
      }
      phase_done();
    };
    for (int g = cluster_id; g < p.n_groups; g += n_clusters) {
#pragma unroll 1
      for (int sl = 0; sl < S * L; ++sl, ++kstep) {
        mph = 0;
        wait_b();
        lin48(WK_TM_SA);
        wait_b();
        lin48(WK_TM_Q);
        ring_adv(rp, 2 * nu, ns);  // K / V slots are consumed by the compute warps
        mbar_wait_short(ctx_bar, kstep & 1);  // every unit's ctx slice has landed in xb
        tcgen05_fence_after();
        lin48(WK_TM_O);
        wait_b();
        mph = 1;
#pragma unroll 1
        for (int kb = 0; kb < 12; ++kb) {  // fc1: slot = k-block kb of the CTA's three 64-row hidden tiles
          const bool mine = (kb & 1) == mw;
          const uint32_t s = next_slot(mine);
          if (!mine) continue;
          const uint32_t a_lo = ring_lo + s * (WK_SLOT_BYTES >> 4), b_lo = xb_lo + kb * (1024 >> 4);
          const uint32_t tm_d = tmem_base + WK_TM_FC1 + mw * WK_NPAD;  // tile t, issuer w: column 16 t + 8 w
          if (elect_one()) {
#pragma unroll
            for (int m = 0; m < 12; ++m) {  // m = 3 ks + t: consecutive MMAs go to different accumulators
              const int t = m % 3, ks = m / 3;
              if (kb < 2 && ks == 0) umma_lo<false>(tm_d + t * 2 * WK_NPAD, a_lo + ((t * 8192 + ks * 32) >> 4), b_lo + ((ks * 32) >> 4), idesc);
              else umma_lo<true>(tm_d + t * 2 * WK_NPAD, a_lo + ((t * 8192 + ks * 32) >> 4), b_lo + ((ks * 32) >> 4), idesc);
            }
            umma_commit(&empty[s]);
          }
          __syncwarp();
        }
        phase_done();
        wait_b();
        mph = 2;
#pragma unroll 1
        for (int mt = 0; mt < 12; ++mt) {  // fc2, split-K: slot = the three k-blocks of 64-row output tile mt
          const bool mine = (mt & 1) == mw;
          const uint32_t s = next_slot(mine);
          if (!mine) continue;
          const uint32_t a_lo = ring_lo + s * (WK_SLOT_BYTES >> 4);
          const uint32_t tm_d = tmem_base + WK_TM_FC2 + mt * WK_NPAD;
          if (elect_one()) {
            umma_lo<false>(tm_d, a_lo, hdn_lo, idesc);
#pragma unroll
            for (int m = 1; m < 12; ++m)  // m = 4 kbi + ks
              umma_lo<true>(tm_d, a_lo + (((m >> 2) * 8192 + (m & 3) * 32) >> 4), hdn_lo + (((m >> 2) * 1024 + (m & 3) * 32) >> 4), idesc);
            umma_commit(&empty[s]);
          }
          __syncwarp();
        }
        phase_done();
      }
      wait_b();
#pragma unroll 1
      for (int j = 0; j < head_slots; ++j) {  // answer head: slot = three k-blocks of 64-row tile j / 4
        const bool mine = (j & 1) == mw;
        const uint32_t s = next_slot(mine);
        if (!mine) continue;
        const uint32_t a_lo = ring_lo + s * (WK_SLOT_BYTES >> 4), b_lo = xb_lo + (j & 3) * (3 * 1024 >> 4);
        const uint32_t tm_d = tmem_base + WK_TM_HEAD + ((j >> 2) * 2 + mw) * WK_NPAD;
        if (elect_one()) {
          if ((j & 3) < 2) umma_lo<false>(tm_d, a_lo, b_lo, idesc);
          else umma_lo<true>(tm_d, a_lo, b_lo, idesc);
#pragma unroll
          for (int m = 1; m < 12; ++m)
            umma_lo<true>(tm_d, a_lo + (((m >> 2) * 8192 + (m & 3) * 32) >> 4), b_lo + (((m >> 2) * 1024 + (m & 3) * 32) >> 4), idesc);
          umma_commit(&empty[s]);
        }
        __syncwarp();
      }
      phase_done();
    }
    if (PROF && lane == 0 && mw == 0) {
#pragma unroll
      for (int t = 0; t < 4; ++t) p.prof[blockIdx.x * 32 + 24 + t] = mwait[t];
    }
  } else if (warp < WK_CWARPS) {
    // =================================================================== compute warps: epilogues, exchanges, LayerNorm, attention
    const int sub = warp & 3;  // TMEM lane quarter this warp may read; an M = 64 tile keeps rows 16 sub .. 16 sub + 15 in its lanes 0..15
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(sub * 32) << 16);
    const uint32_t xbytes = static_cast<uint32_t>(rpc) * WK_F * 4;  // one CTA's slice of an fp32 exchange
    uint32_t e = 0, k = 0, af = 0;
    RingPos rc = {0, 0};  // ring position of the running layer-step's first slot
    long long* pacc = reinterpret_cast<long long*>(smem + sm.bar + 256);  // PROF: [24] cycle buckets of thread 0
    long long pt = 0;
    if (PROF && tid == 0) {
      for (int t = 0; t < 24; ++t) pacc[t] = 0;
      pt = clock64();
    }
    auto stamp = [&](int slot) {  // time since the previous stamp goes to bucket `slot`
      if (PROF && tid == 0) {
        const long long t = clock64();
        pacc[slot] += t - pt;
        pt = t;
      }
    };

    auto wait_acc = [&]() {
      mbar_wait_short(acc_full, af & 1);
      ++af;
      tcgen05_fence_after();
    };
    auto signal_b = [&]() {  // the B operand (xb / hdn) written by the generic proxy is complete
      fence_proxy_async_smem();
      cbar_sync();
      if (tid == 0) mbar_arrive(b_ready);
    };
    // all-to-all of one fp32 slice per CTA: slice for destination d is read at src + d * src_stride
    auto exchange = [&](const float* src, uint32_t src_stride_bytes) {
      const uint32_t buf = e & 1;
      fence_proxy_async_smem();
      cbar_sync();
      if (warp == 0 && lane < WK_CL) {
        if (lane == 0) mbar_expect_tx(&gbar[buf], WK_CL * xbytes);
        bulk_s2c(mapa_u32(smem_u32(G + (buf * WK_CL + rank) * rpc * WK_F), lane), smem_u32(src) + lane * src_stride_bytes, xbytes,
                 mapa_u32(smem_u32(&gbar[buf]), lane));
      }
      mbar_wait_short(&gbar[buf], (e >> 1) & 1);
      ++e;
      return buf;
    };
    // epilogue of a 48-feature product: y = acc + bias (+ residual slice) -> outgoing slice
    auto epi48 = [&](int tm_col, const float* bias48, bool add_res, int pslot) {
      wait_acc();
      stamp(pslot);
      uint32_t v[8], w[8];
      if (warp < 3) {
        tmem_ld_8(t_lane + tm_col, v);
        tmem_ld_8(t_lane + tm_col + WK_NPAD, w);
        tmem_ld_wait();
      }
      tcgen05_fence_before();
      float* o = stg + (e & 1) * rpc * WK_F;
      if (warp < 3 && lane < 16) {
        const int f = 16 * warp + lane;
        const float b = bias48[f];
#pragma unroll
        for (int r = 0; r < WK_NPAD; ++r)
          if (r < rpc) o[r * WK_F + f] = (__uint_as_float(v[r]) + __uint_as_float(w[r])) + b + (add_res ? res[r * WK_F + f] : 0.f);
      }
      return o;
    };
    // row `row` of exchange buffer `buf` -> this lane's 24 values: 16-byte chunks lane, lane + 32, lane + 64 of the 96
    auto load_row = [&](uint32_t buf, int row, float (&v)[24]) {
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        const int col = (lane + 32 * j) * 8, src = col / WK_F, kk = col - src * WK_F;
        const float4* ptr = reinterpret_cast<const float4*>(G + ((buf * WK_CL + src) * rpc + row) * WK_F + kk);
        const float4 a = ptr[0], b = ptr[1];
        v[8 * j + 0] = a.x; v[8 * j + 1] = a.y; v[8 * j + 2] = a.z; v[8 * j + 3] = a.w;
        v[8 * j + 4] = b.x; v[8 * j + 5] = b.y; v[8 * j + 6] = b.z; v[8 * j + 7] = b.w;
      }
    };
    auto normalise = [&](float (&v)[24]) {  // (v - mean) * rstd, no affine
      // one butterfly for both moments, around a per-lane pivot-free shift: the inputs are O(1..10) residual-stream values,
      // so E[x^2] - mean^2 in fp32 keeps ~5 digits of the variance — far below the bf16 rounding of the result
      float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;
#pragma unroll
      for (int t = 0; t < 24; t += 2) {
        s0 += v[t]; s1 += v[t + 1];
        q0 = fmaf(v[t], v[t], q0); q1 = fmaf(v[t + 1], v[t + 1], q1);
      }
      float s = s0 + s1, q = q0 + q1;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        s += __shfl_xor_sync(0xffffffffu, s, o);
        q += __shfl_xor_sync(0xffffffffu, q, o);
      }
      const float mean = s * (1.0f / ENC_D);
      const float rstd = rsqrtf(fmaxf(q * (1.0f / ENC_D) - mean * mean, 0.f) + p.eps);
      const float nm = -mean * rstd;
#pragma unroll
      for (int t = 0; t < 24; ++t) v[t] = fmaf(v[t], rstd, nm);
    };
    // this lane's 24 values of row `row` -> bf16 B operand xb (k-block = 64 columns = 8 chunks; 128B swizzle by row)
    auto store_xb = [&](int row, const float (&v)[24]) {
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        const int ch = lane + 32 * j;
        uint4 u;
        u.x = pack_bf16x2(v[8 * j + 0], v[8 * j + 1]); u.y = pack_bf16x2(v[8 * j + 2], v[8 * j + 3]);
        u.z = pack_bf16x2(v[8 * j + 4], v[8 * j + 5]); u.w = pack_bf16x2(v[8 * j + 6], v[8 * j + 7]);
        *reinterpret_cast<uint4*>(xb + (ch >> 3) * 1024 + row * 128 + (((ch & 7) ^ row) << 4)) = u;
      }
    };
    // LayerNorm of exchange buffer `buf` between two sub-steps of a layer: the normalised rows (gamma / beta are folded
    // into the next product's weights) become the B operand; the affine result of this CTA's 48 columns is the residual
    auto ln_mid = [&](uint32_t buf, const float* g48, const float* b48) {
      if (warp < rpc) {
        float v[24];
        load_row(buf, warp, v);
        normalise(v);
        store_xb(warp, v);
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          const int c = lane + 32 * j - 6 * rank;  // chunk index inside this CTA's column slice
          if (c >= 0 && c < 6) {
#pragma unroll
            for (int t = 0; t < 8; ++t) res[warp * WK_F + c * 8 + t] = fmaf(g48[c * 8 + t], v[8 * j + t], b48[c * 8 + t]);
          }
        }
      }
    };
    // token of a (new) segment: fp32 copy for the segment's final residual, bf16 B operand with the affine applied (layer
    // 0's self-attention weights are not folded), residual slice, optional tap
    auto emit_token = [&](int row, const float (&v)[24], float* tap_row) {
      store_xb(row, v);
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        const int ch = lane + 32 * j, c = ch - 6 * rank;
        float4* t4 = reinterpret_cast<float4*>(tok + row * ENC_D + ch * 8);
        t4[0] = make_float4(v[8 * j + 0], v[8 * j + 1], v[8 * j + 2], v[8 * j + 3]);
        t4[1] = make_float4(v[8 * j + 4], v[8 * j + 5], v[8 * j + 6], v[8 * j + 7]);
        if (c >= 0 && c < 6) {
#pragma unroll
          for (int t = 0; t < 8; ++t) res[row * WK_F + c * 8 + t] = v[8 * j + t];
        }
        if (tap_row != nullptr) {
          float4* o4 = reinterpret_cast<float4*>(tap_row + ch * 8);
          o4[0] = t4[0];
          o4[1] = t4[1];
        }
      }
    };

    for (int g = cluster_id; g < p.n_groups; g += n_clusters) {
      const int r0 = g * rpc;
      // ---- segment 0 starts from the summarisation token itself
      if (warp < rpc) {
        float v[24];
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          const float4* t4 = reinterpret_cast<const float4*>(p.tok0 + (lane + 32 * j) * 8);
          const float4 a = __ldg(t4), b = __ldg(t4 + 1);
          v[8 * j + 0] = a.x; v[8 * j + 1] = a.y; v[8 * j + 2] = a.z; v[8 * j + 3] = a.w;
          v[8 * j + 4] = b.x; v[8 * j + 5] = b.y; v[8 * j + 6] = b.z; v[8 * j + 7] = b.w;
        }
        emit_token(warp, v, nullptr);
      }
      signal_b();
      for (int s = 0; s < S; ++s) {
        for (int n = 0; n < L; ++n, ++k) {
          mbar_wait_short(&par_bar[k & 1], (k >> 1) & 1);
          const float* par = par_base + (k & 1) * WK_PB;
          stamp(0);
          // ---- P1: y1 = x + W_sa x + b_sa ; h1 = LN1(y1)
          {
            const float* o = epi48(WK_TM_SA, par + PB_SA_B, true, 1);
            const uint32_t buf = exchange(o, 0);
            stamp(2);
            ln_mid(buf, par + PB_G1, par + PB_BE1);
            signal_b();
            stamp(3);
          }
          // ---- P2: q = W_q h1 + b_q  (stays in its exchange buffer)
          uint32_t qbuf;
          {
            const float* o = epi48(WK_TM_Q, par + PB_Q_B, false, 4);
            qbuf = exchange(o, 0);
            stamp(5);
          }
          // ---- P3: cross attention of this CTA's (row, head) units; the outputs go straight into every CTA's xb
          if (tid == 0) mbar_expect_tx(ctx_bar, ctx_total);
          for (int ui = 0; ui < nu; ++ui) {
            const int u = u0 + ui, head = u / rpc, lr = u - head * rpc;
            RingPos rk = rc;
            ring_adv(rk, 6 + 2 * ui, ns);
            RingPos rv = rk;
            ring_adv(rv, 1, ns);
            const uint32_t sk = rk.s, sv = rv.s;
            const uint8_t* Ks = ring + sk * WK_SLOT_BYTES;
            const uint8_t* Vs = ring + sv * WK_SLOT_BYTES;
            // ---- scores: warp w owns keys [24 w, 24 w + 24), 4 lanes per key (lane c takes the 16-byte chunks c and c + 4 of the
            // 128-byte K row; odd keys read them in the opposite order so that a quarter-warp covers all 32 banks)
            const int c4 = lane & 3, kq = lane >> 2;
            float qa[8], qb[8];
            {
              const int cola = head * 64 + c4 * 8, srca = cola / WK_F, colb = cola + 32, srcb = colb / WK_F;
              const float4* pa4 = reinterpret_cast<const float4*>(G + ((qbuf * WK_CL + srca) * rpc + lr) * WK_F + (cola - srca * WK_F));
              const float4* pb4 = reinterpret_cast<const float4*>(G + ((qbuf * WK_CL + srcb) * rpc + lr) * WK_F + (colb - srcb * WK_F));
              const float4 a0 = pa4[0], a1 = pa4[1], b0 = pb4[0], b1 = pb4[1];
              qa[0] = a0.x; qa[1] = a0.y; qa[2] = a0.z; qa[3] = a0.w; qa[4] = a1.x; qa[5] = a1.y; qa[6] = a1.z; qa[7] = a1.w;
              qb[0] = b0.x; qb[1] = b0.y; qb[2] = b0.z; qb[3] = b0.w; qb[4] = b1.x; qb[5] = b1.y; qb[6] = b1.z; qb[7] = b1.w;
            }
            mbar_wait_short(&full[sk], rk.ph);
            stamp(6);
            const int k_lo = warp * 24;
            uint4 ka[3], kb4[3];
#pragma unroll
            for (int it = 0; it < 3; ++it) {
              const uint8_t* row = Ks + min(k_lo + it * 8 + kq, n_keys - 1) * 128;
              const uint4 x = *reinterpret_cast<const uint4*>(row + ((c4 + 4 * (kq & 1)) & 7) * 16);
              const uint4 y = *reinterpret_cast<const uint4*>(row + ((c4 + 4 * ((kq & 1) ^ 1)) & 7) * 16);
              ka[it] = (kq & 1) ? y : x;   // chunk c
              kb4[it] = (kq & 1) ? x : y;  // chunk c + 4
            }
            float sc[3];
#pragma unroll
            for (int it = 0; it < 3; ++it) {
              const uint32_t wa[4] = {ka[it].x, ka[it].y, ka[it].z, ka[it].w}, wb[4] = {kb4[it].x, kb4[it].y, kb4[it].z, kb4[it].w};
              float d0 = 0.f, d1 = 0.f;
#pragma unroll
              for (int t = 0; t < 4; ++t) {
                const float2 fa = unpack_bf16x2(wa[t]), fb = unpack_bf16x2(wb[t]);
                d0 = fmaf(qa[2 * t], fa.x, d0); d0 = fmaf(qa[2 * t + 1], fa.y, d0);
                d1 = fmaf(qb[2 * t], fb.x, d1); d1 = fmaf(qb[2 * t + 1], fb.y, d1);
              }
              sc[it] = d0 + d1;
            }
#pragma unroll
            for (int it = 0; it < 3; ++it) sc[it] += __shfl_xor_sync(0xffffffffu, sc[it], 1);
#pragma unroll
            for (int it = 0; it < 3; ++it) sc[it] += __shfl_xor_sync(0xffffffffu, sc[it], 2);
#pragma unroll
            for (int it = 0; it < 3; ++it)
              if (k_lo + it * 8 + kq >= n_keys) sc[it] = -INFINITY;
            float mx = fmaxf(fmaxf(sc[0], sc[1]), sc[2]);
            mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 4));
            mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 8));
            mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 16));
            const float mref = mx == -INFINITY ? 0.f : mx;  // a warp without keys contributes exact zeros
            // p = exp(s - warp max) of the warp's 24 keys -> this warp's strip of shared memory (read back by the same warp)
            float* pa = att + (ui & 1) * WK_ATT_FLOATS;
            float* sP = pa + WK_CWARPS * 64 + 2 * WK_CWARPS + warp * 24;
            float lsum = 0.f;
#pragma unroll
            for (int it = 0; it < 3; ++it) {
              const float pj = ex2f((sc[it] - mref) * 1.4426950408889634f);
              lsum += pj;
              if (c4 == 0) sP[it * 8 + kq] = pj;
            }
            lsum += __shfl_xor_sync(0xffffffffu, lsum, 4);
            lsum += __shfl_xor_sync(0xffffffffu, lsum, 8);
            lsum += __shfl_xor_sync(0xffffffffu, lsum, 16);  // every key was counted by one lane per c4: this IS the sum over keys
            __syncwarp();
            stamp(7);
            // ---- partial P V over the same 24 keys, two output dims per lane: a V row is one conflict-free 128-byte read per
            // warp, the probabilities are shared-memory broadcasts, nothing is shuffled
            mbar_wait_short(&full[sv], rv.ph);
            stamp(8);
            float2 o2 = make_float2(0.f, 0.f);
#pragma unroll
            for (int j4 = 0; j4 < 6; ++j4) {
              const float4 p4 = *reinterpret_cast<const float4*>(sP + j4 * 4);
              const float pj[4] = {p4.x, p4.y, p4.z, p4.w};
#pragma unroll
              for (int t = 0; t < 4; ++t) {
                const float2 f = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(Vs + min(k_lo + j4 * 4 + t, n_keys - 1) * 128 + lane * 4));
                o2.x = fmaf(pj[t], f.x, o2.x);
                o2.y = fmaf(pj[t], f.y, o2.y);
              }
            }
            *reinterpret_cast<float2*>(pa + warp * 64 + lane * 2) = o2;
            if (lane == 0) {
              pa[WK_CWARPS * 64 + warp] = mx;
              pa[WK_CWARPS * 64 + WK_CWARPS + warp] = lsum;
            }
            cbar_sync();
            if (tid == 0) {  // every warp is done with the unit's K and V slots
              mbar_arrive(&empty[sk]);
              mbar_arrive(&empty[sv]);
            }
            if (tid < 64) {  // combine the eight partials (flash-style rescale by the warp maxima)
              float M = -INFINITY;
#pragma unroll
              for (int w = 0; w < WK_CWARPS; ++w) M = fmaxf(M, pa[WK_CWARPS * 64 + w]);
              float num = 0.f, den = 0.f;
#pragma unroll
              for (int w = 0; w < WK_CWARPS; ++w) {
                const float mw = pa[WK_CWARPS * 64 + w];
                const float sc_w = mw == -INFINITY ? 0.f : ex2f((mw - M) * 1.4426950408889634f);
                num = fmaf(pa[w * 64 + tid], sc_w, num);
                den = fmaf(pa[WK_CWARPS * 64 + WK_CWARPS + w], sc_w, den);
              }
              // column tid of row lr in k-block `head` of xb: chunk (tid / 8) ^ lr (128B swizzle) of the 128-byte row
              *reinterpret_cast<bf16*>(ctx_stg + (unit_off(u) - ctx_off) + ((((tid >> 3) ^ lr) & 7) << 4) + (tid & 7) * 2) =
                  __float2bfloat16(num / den);
            }
            stamp(9);
          }
          fence_proxy_async_smem();
          cbar_sync();
          if (warp == 0 && lane < WK_CL && nu > 0)
            bulk_s2c(mapa_u32(smem_u32(xb + ctx_off), lane), smem_u32(ctx_stg), ctx_bytes, mapa_u32(smem_u32(ctx_bar), lane));
          // ---- P4: y2 = h1 + W_o ctx + b_o ; h2 = LN2(y2)
          {
            stamp(10);
            const float* o = epi48(WK_TM_O, par + PB_O_B, true, 11);
            const uint32_t buf = exchange(o, 0);
            stamp(12);
            ln_mid(buf, par + PB_G2, par + PB_BE2);
            signal_b();
            stamp(13);
          }
          // ---- P5: hdn = gelu(W_1 h2 + b_1) for this CTA's 192 hidden features -> bf16 B operand of the split-K fc2
          wait_acc();
          stamp(14);
          for (int t = (warp >> 2); t < 3; t += 2) {  // warps 0-3: tiles 0 and 2, warps 4-7: tile 1
            uint32_t v[8], w[8];
            tmem_ld_8(t_lane + WK_TM_FC1 + t * 2 * WK_NPAD, v);
            tmem_ld_8(t_lane + WK_TM_FC1 + t * 2 * WK_NPAD + WK_NPAD, w);
            tmem_ld_wait();
            if (lane < 16) {
              const float b = par[PB_B1 + 64 * t + 16 * sub + lane];
              const int c = 2 * sub + (lane >> 3);
#pragma unroll
              for (int r = 0; r < WK_NPAD; ++r)
                if (r < rpc)
                  *reinterpret_cast<bf16*>(hdn + t * 1024 + r * 128 + (((c ^ r) & 7) << 4) + (lane & 7) * 2) =
                      __float2bfloat16(gelu_erf((__uint_as_float(v[r]) + __uint_as_float(w[r])) + b));
            }
          }
          tcgen05_fence_before();
          signal_b();
          stamp(15);
          // ---- P6: fc2 partials over this CTA's K slice -> reduce-scatter -> y3 = h2 + sum + b_2 -> all-gather -> LN3
          wait_acc();
          stamp(16);
          {
            // staging = the exchange buffer that is NOT received into next (its last content, y2, was consumed by LN2), laid
            // out [destination][row][48]. It is also where y3 arrives two exchanges later: source j's slice overwrites region
            // j only after j has received this CTA's partials, i.e. after the copy out of region j has completed.
            float* part = G + ((e & 1) ^ 1) * WK_CL * rpc * WK_F;
            for (int t = (warp >> 2) * 6; t < (warp >> 2) * 6 + 6; ++t) {
              uint32_t v[8];
              tmem_ld_8(t_lane + WK_TM_FC2 + t * WK_NPAD, v);
              tmem_ld_wait();
              if (lane < 16) {
                const int f = 64 * t + 16 * sub + lane, d = f / WK_F, kk = f - d * WK_F;
#pragma unroll
                for (int r = 0; r < WK_NPAD; ++r)
                  if (r < rpc) part[(d * rpc + r) * WK_F + kk] = __uint_as_float(v[r]);
              }
            }
            tcgen05_fence_before();
            stamp(17);
            const uint32_t pbuf = exchange(part, xbytes);
            stamp(18);
            float* o = stg + (e & 1) * rpc * WK_F;
            for (int t = tid; t < rpc * WK_F; t += WK_CTHREADS) {
              const int r = t / WK_F, kk = t - r * WK_F;
              float a = par[PB_B2 + kk] + res[t];
#pragma unroll
              for (int src = 0; src < WK_CL; ++src) a += G[((pbuf * WK_CL + src) * rpc + r) * WK_F + kk];
              o[t] = a;
            }
            stamp(19);
            const uint32_t buf = exchange(o, 0);
            stamp(20);
            if (n + 1 < L) {
              ln_mid(buf, par + PB_G3, par + PB_BE3);
            } else if (warp < rpc) {
              // end of the segment: tok' = LN_f(tok + LN3(y3)) (fusionv3.py:47-48), full affine on both norms
              float v[24];
              load_row(buf, warp, v);
              normalise(v);
#pragma unroll
              for (int j = 0; j < 3; ++j) {
                const int col = (lane + 32 * j) * 8;
#pragma unroll
                for (int t = 0; t < 8; ++t)
                  v[8 * j + t] = fmaf(__ldg(p.tail + col + t), v[8 * j + t], __ldg(p.tail + ENC_D + col + t)) + tok[warp * ENC_D + col + t];
              }
              normalise(v);
#pragma unroll
              for (int j = 0; j < 3; ++j) {
                const int col = (lane + 32 * j) * 8;
#pragma unroll
                for (int t = 0; t < 8; ++t) v[8 * j + t] = fmaf(__ldg(p.f_g + col + t), v[8 * j + t], __ldg(p.f_b + col + t));
              }
              float* tap = nullptr;
              if (p.tokens_tap != nullptr && rank == 0 && r0 + warp < p.R)
                tap = p.tokens_tap + (static_cast<size_t>(s) * p.R + r0 + warp) * ENC_D;
              emit_token(warp, v, tap);
            }
            signal_b();
            stamp(21);
          }
          ring_adv(rc, WK_W_SLOTS + 2 * nu, ns);
        }
      }
      // ---- answer head on the final token: this CTA's 64 * mt features, straight to global memory
      wait_acc();
      if (warp < 4) {
        for (int t = 0; t < p.mt_head; ++t) {
          uint32_t v[8], w[8];
          tmem_ld_8(t_lane + WK_TM_HEAD + t * 2 * WK_NPAD, v);
          tmem_ld_8(t_lane + WK_TM_HEAD + t * 2 * WK_NPAD + WK_NPAD, w);
          tmem_ld_wait();
          const int f = (rank * p.mt_head + t) * 64 + 16 * sub + lane;
          if (lane < 16 && f < p.n_out) {
            const float b = __ldg(p.head_bias + f);
#pragma unroll
            for (int r = 0; r < WK_NPAD; ++r) {
              if (r < rpc && r0 + r < p.R) {
                float y = (__uint_as_float(v[r]) + __uint_as_float(w[r])) + b;
                if (p.act == ACT_GELU) y = gelu_erf(y);
                else if (p.act == ACT_RELU) y = fmaxf(y, 0.f);
                p.out[static_cast<size_t>(r0 + r) * p.n_out + f] = y;
              }
            }
          }
        }
      }
      tcgen05_fence_before();
      cbar_sync();  // the next group's token must not overwrite xb / tok while a warp is still in this group's head
      ring_adv(rc, head_slots, ns);
      stamp(22);
    }
    if (PROF && tid == 0) {
      for (int t = 0; t < 24; ++t) p.prof[blockIdx.x * 32 + t] = pacc[t];
    }
  }

  tcgen05_fence_before();
  cluster_sync_all();  // no CTA exits while a peer can still copy into its shared memory
  if (warp == 10) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, WK_TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// one-time repack of the decoder weights into the per-CTA streaming order
// ------------------------------------------------------------------------------------------------------------------
// One thread per 16-byte chunk (8 bf16 along K) of the weight and head streams.
__global__ void walk_pack_weights_kernel(const EncLayerF32* __restrict__ layers, int n_layers, const float* __restrict__ fc_w,
                                         int n_out, bf16* __restrict__ dst) {
  const WalkPackLayout lay = walk_pack_layout(n_layers, n_out);
  const long long total = (lay.w_rows + lay.head_rows) * 8;
  const int mt_head = walk_head_tiles(n_out);
  for (long long c = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; c < total;
       c += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long row = c >> 3;
    const int cc = static_cast<int>(c & 7);
    const float* src = nullptr;   // 8 consecutive fp32 of the source matrix
    const float* fold = nullptr;  // gamma of the LayerNorm in front of this product (same columns), or nullptr
    if (row < lay.w_rows) {
      const int n = static_cast<int>(row / (WK_CL * WK_W_ROWS));
      const int rem = static_cast<int>(row - static_cast<long long>(n) * (WK_CL * WK_W_ROWS));
      const int r = rem / WK_W_ROWS, q = rem - r * WK_W_ROWS;
      const EncLayerF32 Lw = layers[n];
      if (q < 3 * 12 * WK_F) {  // sa | q | o: [k-block][48 rows]
        const int ph = q / (12 * WK_F), qq = q - ph * (12 * WK_F), kb = qq / WK_F, fr = qq - kb * WK_F;
        const float* W = ph == 0 ? Lw.sa_w : ph == 1 ? Lw.q_w : Lw.o_w;
        src = W + static_cast<size_t>(WK_F * r + fr) * ENC_D + kb * 64 + cc * 8;
        if (ph == 0 && n > 0) fold = layers[n - 1].n3g + kb * 64 + cc * 8;
        if (ph == 1) fold = Lw.n1g + kb * 64 + cc * 8;
      } else if (q < 3 * 12 * WK_F + 12 * WK_H) {  // fc1: [k-block][3 tiles][64 rows]
        const int qq = q - 3 * 12 * WK_F, kb = qq / WK_H, hf = qq - kb * WK_H;
        src = Lw.w1 + static_cast<size_t>(WK_H * r + hf) * ENC_D + kb * 64 + cc * 8;
        fold = Lw.n2g + kb * 64 + cc * 8;
      } else {  // fc2, split-K: [64-row output tile][3 k-blocks of this CTA's hidden slice][64 rows]
        const int qq = q - 3 * 12 * WK_F - 12 * WK_H, mt = qq / 192, kbi = (qq - mt * 192) / 64, fr = qq & 63;
        src = Lw.w2 + static_cast<size_t>(64 * mt + fr) * (4 * ENC_D) + WK_H * r + kbi * 64 + cc * 8;
      }
    } else {  // head: [rank][tile][k-block][64 rows]
      const int hq = static_cast<int>(row - lay.w_rows);
      const int fr = hq & 63, kb = (hq >> 6) % 12, rt = (hq >> 6) / 12;  // rt = rank * mt_head + tile
      const int f = rt * 64 + fr;
      (void)mt_head;
      if (f < n_out) src = fc_w + static_cast<size_t>(f) * ENC_D + kb * 64 + cc * 8;
    }
    uint4 u = make_uint4(0, 0, 0, 0);
    if (src != nullptr) {
      float4 a = __ldg(reinterpret_cast<const float4*>(src)), b = __ldg(reinterpret_cast<const float4*>(src) + 1);
      if (fold != nullptr) {
        const float4 ga = __ldg(reinterpret_cast<const float4*>(fold)), gb = __ldg(reinterpret_cast<const float4*>(fold) + 1);
        a.x *= ga.x; a.y *= ga.y; a.z *= ga.z; a.w *= ga.w;
        b.x *= gb.x; b.y *= gb.y; b.z *= gb.z; b.w *= gb.w;
      }
      u.x = pack_bf16x2(a.x, a.y); u.y = pack_bf16x2(a.z, a.w); u.z = pack_bf16x2(b.x, b.y); u.w = pack_bf16x2(b.z, b.w);
    }
    reinterpret_cast<uint4*>(dst)[c] = u;
  }
}

// b' = b + W beta (beta of the LayerNorm folded into W), one warp per output feature
__device__ __forceinline__ float dot768(const float* __restrict__ w, const float* __restrict__ beta, int lane) {
  float s = 0.f;
  for (int k = lane; k < ENC_D; k += 32) s = fmaf(__ldg(w + k), __ldg(beta + k), s);
  return warp_sum(s);
}
// block (n, r): the parameter block of layer n for CTA rank r; block n == n_layers: tail and head bias
__global__ void __launch_bounds__(256) walk_pack_params_kernel(const EncLayerF32* __restrict__ layers, int n_layers,
                                                               const float* __restrict__ fc_b, int n_out, float* __restrict__ params,
                                                               float* __restrict__ tail, float* __restrict__ hbias) {
  const int n = blockIdx.x / WK_CL, r = blockIdx.x % WK_CL;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (n == n_layers) {
    const EncLayerF32 Ll = layers[n_layers - 1];
    for (int i = threadIdx.x + r * 256; i < ENC_D; i += 256 * WK_CL) {
      tail[i] = Ll.n3g[i];
      tail[ENC_D + i] = Ll.n3b[i];
    }
    const int nh = WK_CL * 64 * walk_head_tiles(n_out);
    for (int i = threadIdx.x + r * 256; i < nh; i += 256 * WK_CL) hbias[i] = i < n_out ? fc_b[i] : 0.f;
    return;
  }
  const EncLayerF32 Lw = layers[n];
  float* pb = params + (static_cast<size_t>(n) * WK_CL + r) * WK_PB;
  for (int i = threadIdx.x; i < WK_F; i += 256) {
    const int f = WK_F * r + i;
    pb[PB_O_B + i] = Lw.o_b[f];
    pb[PB_B2 + i] = Lw.b2[f];
    pb[PB_G1 + i] = Lw.n1g[f]; pb[PB_BE1 + i] = Lw.n1b[f];
    pb[PB_G2 + i] = Lw.n2g[f]; pb[PB_BE2 + i] = Lw.n2b[f];
    pb[PB_G3 + i] = Lw.n3g[f]; pb[PB_BE3 + i] = Lw.n3b[f];
  }
  for (int i = warp; i < 2 * WK_F + WK_H; i += 8) {
    float v;
    int slot;
    if (i < WK_F) {  // sa: the norm in front is LN3 of the previous layer (layer 0 takes the token with its affine applied)
      const int f = WK_F * r + i;
      v = Lw.sa_b[f] + (n > 0 ? dot768(Lw.sa_w + static_cast<size_t>(f) * ENC_D, layers[n - 1].n3b, lane) : 0.f);
      slot = PB_SA_B + i;
    } else if (i < 2 * WK_F) {
      const int f = WK_F * r + i - WK_F;
      v = Lw.q_b[f] + dot768(Lw.q_w + static_cast<size_t>(f) * ENC_D, Lw.n1b, lane);
      slot = PB_Q_B + i - WK_F;
    } else {
      const int hf = WK_H * r + i - 2 * WK_F;
      v = Lw.b1[hf] + dot768(Lw.w1 + static_cast<size_t>(hf) * ENC_D, Lw.n2b, lane);
      slot = PB_B1 + i - 2 * WK_F;
    }
    if (lane == 0) pb[slot] = v;
  }
}

}  // namespace lrce

using namespace lrce;

extern "C" size_t lrce_encoder_walk_pack_bytes(int n_layers, int n_out) {
  if (n_layers <= 0 || n_out <= 0) return 0;
  return walk_pack_layout(n_layers, n_out).total;
}

extern "C" int lrce_encoder_walk_pack(const void* layer_table, int n_layers, const float* fc_w, const float* fc_b, int n_out,
                                      void* packed, void* stream) {
  int rc = require_sm100();
  if (rc != LRCE_OK) return rc;
  LRCE_REQUIRE(layer_table && fc_w && fc_b && packed, "lrce_encoder_walk_pack: null argument");
  LRCE_REQUIRE(n_layers > 0 && n_out > 0 && walk_head_tiles(n_out) <= WK_MAX_HEAD_TILES,
               "lrce_encoder_walk_pack: bad sizes (layers=%d, n_out=%d; at most %d outputs)", n_layers, n_out,
               WK_CL * 64 * WK_MAX_HEAD_TILES);
  LRCE_REQUIRE((reinterpret_cast<uintptr_t>(packed) & 127) == 0, "lrce_encoder_walk_pack: packed buffer must be 128-byte aligned");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const WalkPackLayout lay = walk_pack_layout(n_layers, n_out);
  uint8_t* base = reinterpret_cast<uint8_t*>(packed);
  walk_pack_weights_kernel<<<sm_count() * 8, 256, 0, s>>>(reinterpret_cast<const EncLayerF32*>(layer_table), n_layers, fc_w, n_out,
                                                          reinterpret_cast<bf16*>(base));
  rc = check_launch("walk_pack_weights_kernel");
  if (rc != LRCE_OK) return rc;
  walk_pack_params_kernel<<<(n_layers + 1) * WK_CL, 256, 0, s>>>(reinterpret_cast<const EncLayerF32*>(layer_table), n_layers, fc_b,
                                                                 n_out, reinterpret_cast<float*>(base + lay.off_params),
                                                                 reinterpret_cast<float*>(base + lay.off_tail),
                                                                 reinterpret_cast<float*>(base + lay.off_hbias));
  return check_launch("walk_pack_params_kernel");
}

// co-resident 16-CTA clusters of the walk kernel on the current device (cached per device)
static int walk_max_clusters(int* out) {
  static thread_local int cached[64];
  static thread_local uint64_t configured = 0;
  const int dev = current_device();
  if (needs_device_setup(&configured)) {
    cudaError_t e = cudaSuccess;
    for (int v = 0; v < 4 && e == cudaSuccess; ++v) {
      const void* fn = v == 0   ? reinterpret_cast<const void*>(encoder_walk_kernel<false, 1>)
                       : v == 1 ? reinterpret_cast<const void*>(encoder_walk_kernel<true, 1>)
                       : v == 2 ? reinterpret_cast<const void*>(encoder_walk_kernel<true, 0>)
                                : reinterpret_cast<const void*>(encoder_walk_kernel<false, 0>);
      e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, WK_SMEM_MAX);
      if (e == cudaSuccess) e = cudaFuncSetAttribute(fn, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    }
    int ncl = 0;
    if (e == cudaSuccess) {
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(WK_CL);
      cfg.blockDim = dim3(WK_THREADS);
      cfg.dynamicSmemBytes = WK_SMEM_MAX;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeClusterDimension;
      attr[0].val.clusterDim.x = WK_CL;
      attr[0].val.clusterDim.y = 1;
      attr[0].val.clusterDim.z = 1;
      cfg.attrs = attr;
      cfg.numAttrs = 1;
      e = cudaOccupancyMaxActiveClusters(&ncl, encoder_walk_kernel<false, 1>, &cfg);
    }
    if (e != cudaSuccess || ncl < 1) {
      set_error("encoder_walk_kernel: no 16-CTA cluster can be made resident (%s, clusters=%d)", cudaGetErrorString(e), ncl);
      cudaGetLastError();
      return LRCE_ECUDA;
    }
    if (dev >= 0 && dev < 64) cached[dev] = ncl;
    mark_device_setup(&configured);
    *out = ncl;
    return LRCE_OK;
  }
  *out = cached[dev];
  return LRCE_OK;
}

// Host-side plan of a walk over `rows` rows on a device that can keep `max_clusters` 16-CTA clusters resident: rows per
// cluster (spread over every cluster, at most WK_NPAD rows per pass), row groups, clusters launched, ring slots that fit next
// to the per-row buffers, dynamic shared memory. Pure host arithmetic (no CUDA call): also what tests/ checks on the CPU.
extern "C" int lrce_encoder_walk_plan(int rows, int max_clusters, int* rows_per_cluster, int* n_groups, int* clusters,
                                      int* ring_slots, int* smem_bytes) {
  LRCE_REQUIRE(rows > 0 && max_clusters > 0 && rows_per_cluster && n_groups && clusters && ring_slots && smem_bytes,
               "lrce_encoder_walk_plan: bad arguments (rows=%d, max_clusters=%d)", rows, max_clusters);
  const int passes = (rows + WK_NPAD * max_clusters - 1) / (WK_NPAD * max_clusters);
  const int rpc = (rows + passes * max_clusters - 1) / (passes * max_clusters);
  const int groups = (rows + rpc - 1) / rpc;
  int ns = WK_MAX_SLOTS;
  while (ns > 0 && walk_smem(rpc, ns).total > WK_SMEM_MAX) --ns;
  LRCE_REQUIRE(ns >= 3, "lrce_encoder_walk: shared memory leaves only %d ring slots for %d rows per cluster", ns, rpc);
  *rows_per_cluster = rpc;
  *n_groups = groups;
  *clusters = groups < max_clusters ? groups : max_clusters;
  *ring_slots = ns;
  *smem_bytes = walk_smem(rpc, ns).total;
  return LRCE_OK;
}

static int walk_launch(const void* packed, int n_layers, const void* kv_video, const void* kv_text, int ld_kv, const float* tok0,
                       const float* f_gamma, const float* f_beta, float eps, int n_out, int act, float* out, float* tokens_tap,
                       int rows, int S, int Tv, int Lt, int n_cand, void* stream, long long* prof, int max_clusters, int variant) {
  int rc = require_sm100();
  if (rc != LRCE_OK) return rc;
  LRCE_REQUIRE(packed && kv_video && kv_text && tok0 && f_gamma && f_beta && out, "lrce_encoder_walk: null argument");
  LRCE_REQUIRE(n_layers > 0 && rows > 0 && S > 0 && Tv > 0 && Lt > 0 && n_cand > 0 && n_out > 0 && rows % n_cand == 0,
               "lrce_encoder_walk: bad sizes (layers=%d rows=%d S=%d Tv=%d Lt=%d cand=%d out=%d)", n_layers, rows, S, Tv, Lt,
               n_cand, n_out);
  LRCE_REQUIRE(Tv + Lt <= WK_MAX_KEYS && Tv <= 256 && Lt <= 256, "lrce_encoder_walk: %d memory tokens exceed the %d-key limit",
               Tv + Lt, WK_MAX_KEYS);
  LRCE_REQUIRE(walk_head_tiles(n_out) <= WK_MAX_HEAD_TILES, "lrce_encoder_walk: n_out=%d exceeds %d", n_out,
               WK_CL * 64 * WK_MAX_HEAD_TILES);
  LRCE_REQUIRE(ld_kv % 8 == 0 && ld_kv >= n_layers * 2 * ENC_D, "lrce_encoder_walk: K/V row pitch %d too small / unaligned", ld_kv);
  LRCE_REQUIRE(act >= 0 && act <= 2, "lrce_encoder_walk: unknown activation %d", act);
  LRCE_REQUIRE((reinterpret_cast<uintptr_t>(packed) & 127) == 0, "lrce_encoder_walk: packed buffer must be 128-byte aligned");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  int ncl = 0;
  rc = walk_max_clusters(&ncl);
  if (rc != LRCE_OK) return rc;
  if (max_clusters > 0 && max_clusters < ncl) ncl = max_clusters;
  int rpc = 0, n_groups = 0, clusters = 0, ns = 0, smem_bytes = 0;
  rc = lrce_encoder_walk_plan(rows, ncl, &rpc, &n_groups, &clusters, &ns, &smem_bytes);
  if (rc != LRCE_OK) return rc;
  const WalkPackLayout lay = walk_pack_layout(n_layers, n_out);
  const uint8_t* base = reinterpret_cast<const uint8_t*>(packed);
  WalkParams p;
  p.params = reinterpret_cast<const float*>(base + lay.off_params);
  p.tail = reinterpret_cast<const float*>(base + lay.off_tail);
  p.head_bias = reinterpret_cast<const float*>(base + lay.off_hbias);
  p.tok0 = tok0; p.f_g = f_gamma; p.f_b = f_beta;
  p.out = out; p.tokens_tap = tokens_tap;
  p.n_layers = n_layers; p.n_out = n_out; p.act = act; p.R = rows; p.S = S; p.Tv = Tv; p.Lt = Lt; p.n_cand = n_cand;
  p.rpc = rpc; p.n_groups = n_groups; p.ns = ns; p.mt_head = walk_head_tiles(n_out);
  p.head_row0 = static_cast<int>(lay.w_rows);
  p.eps = eps;
  p.prof = prof;
  CUtensorMap tmW, tmV, tmT;
  rc = make_tmap_2d_bf16(&tmW, packed, 64, static_cast<uint64_t>(lay.w_rows + lay.head_rows), 64, 64, WK_SLOT_ROWS);
  if (rc != LRCE_OK) return rc;
  {
    const uint64_t strides[1] = {static_cast<uint64_t>(ld_kv) * 2};
    const uint64_t dv[2] = {static_cast<uint64_t>(n_layers) * 2 * ENC_D, static_cast<uint64_t>(rows / n_cand) * S * Tv};
    const uint32_t bv[2] = {64, static_cast<uint32_t>(Tv)};
    rc = make_tmap_nd_bf16(&tmV, kv_video, 2, dv, strides, bv, 0, 128);
    if (rc != LRCE_OK) return rc;
    const uint64_t dt[2] = {static_cast<uint64_t>(n_layers) * 2 * ENC_D, static_cast<uint64_t>(rows) * Lt};
    const uint32_t bt[2] = {64, static_cast<uint32_t>(Lt)};
    rc = make_tmap_nd_bf16(&tmT, kv_text, 2, dt, strides, bt, 0, 128);
    if (rc != LRCE_OK) return rc;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(clusters * WK_CL);
  cfg.blockDim = dim3(WK_THREADS);
  cfg.dynamicSmemBytes = static_cast<size_t>(smem_bytes);
  cfg.stream = s;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = WK_CL;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1 + pdl_attr(attr + 1);
  cudaError_t e = prof == nullptr ? (variant == 0 ? cudaLaunchKernelEx(&cfg, encoder_walk_kernel<false, 0>, tmW, tmV, tmT, p)
                                                  : cudaLaunchKernelEx(&cfg, encoder_walk_kernel<false, 1>, tmW, tmV, tmT, p))
                  : variant == 0  ? cudaLaunchKernelEx(&cfg, encoder_walk_kernel<true, 0>, tmW, tmV, tmT, p)
                                  : cudaLaunchKernelEx(&cfg, encoder_walk_kernel<true, 1>, tmW, tmV, tmT, p);
  if (e != cudaSuccess) {
    set_error("cudaLaunchKernelEx(encoder_walk_kernel, %d clusters of %d): %s", clusters, WK_CL, cudaGetErrorString(e));
    return LRCE_ECUDA;
  }
  return check_launch("encoder_walk_kernel");
}

extern "C" int lrce_encoder_walk(const void* packed, int n_layers, const void* kv_video, const void* kv_text, int ld_kv,
                                 const float* tok0, const float* f_gamma, const float* f_beta, float eps, int n_out, int act,
                                 float* out, float* tokens_tap, int rows, int S, int Tv, int Lt, int n_cand, void* stream) {
  return walk_launch(packed, n_layers, kv_video, kv_text, ld_kv, tok0, f_gamma, f_beta, eps, n_out, act, out, tokens_tap, rows, S, Tv,
                     Lt, n_cand, stream, nullptr, 0, 1);
}

// Instrumented instantiation (tools only; everything is a per-call argument, the library keeps no profiling state): `prof`
// = NULL (uninstrumented kernel, for timing a variant) or device int64 [grid CTAs][32], receives the cycles thread 0 of every CTA's compute warps spent per sub-step of the chain;
// max_clusters > 0 caps the number of clusters (rows per cluster grow accordingly); variant selects a code variant kept
// for same-box A/B measurements (1 = the production code).
extern "C" int lrce_encoder_walk_profile(const void* packed, int n_layers, const void* kv_video, const void* kv_text, int ld_kv,
                                         const float* tok0, const float* f_gamma, const float* f_beta, float eps, int n_out,
                                         int act, float* out, float* tokens_tap, int rows, int S, int Tv, int Lt, int n_cand,
                                         void* stream, long long* prof, int max_clusters, int variant) {
  return walk_launch(packed, n_layers, kv_video, kv_text, ld_kv, tok0, f_gamma, f_beta, eps, n_out, act, out, tokens_tap, rows, S, Tv,
                     Lt, n_cand, stream, prof, max_clusters, variant);
}
