// window_attn_pp.cuh — "ping-pong" variant of the window attention kernel (included by window_attn.cu, shares its helpers):
// TWO work units in flight per SM, so that the serial per-unit chain of one unit's softmax warps (S wait, TMEM loads,
// maximum exchange, P hand-over, O load, stores — 75 % of the time of the one-unit kernel, profiles/README.md) is covered
// by the other unit's arithmetic.
//
// A CTA's units alternate between two independent streams (even / odd). Per stream: ONE 160-column S accumulator that
// holds row tile 0 and then row tile 1 of the unit, two 32-column O accumulators, 8 softmax warps (2 per TMEM lane
// quarter), one MMA thread, double-buffered q / k / v staging. TMEM: 2 x (160 + 64) = 448 columns.
//   row tile 0 : a warp owns 80 key columns of its 32 rows in two 40-column halves. Pass A reads them from TMEM for the
//                raw maximum (scores are not kept), the two warps of a row exchange maxima, pass B re-reads each half,
//                forms the probabilities and writes them as packed bf16 pairs INTO the S columns it has just consumed
//                (tcgen05.st): P never touches shared memory and P v runs with its A operand in TMEM.
//   row tile 1 : the 32 slots are replicated over the four lane quarters (as in the one-unit kernel); the eight warps
//                split the 160 columns 24 / 16 per quarter, P goes to a 10 KB shared-memory tile (A operand from smem).
// The epilogue of a tile runs after the next tile's probabilities have been handed over (its P v is long complete by then).
#pragma once

namespace lrce {

constexpr int W2_THREADS = 20 * 32;
constexpr int W2_WARP_LOADER = 16, W2_WARP_MMA = 17, W2_WARP_TMEM = 19;  // MMA threads: warp 17 (stream 0), 18 (stream 1)
constexpr int W2_OFF_STAGE = 0;                                          // [stream][buffer] x WA_STAGE_BYTES
constexpr int W2_OFF_P1 = 4 * WA_STAGE_BYTES;                            // [stream] x WA_P1_BYTES
constexpr int W2_OFF_BIAS = W2_OFF_P1 + 2 * WA_P1_BYTES;
constexpr int W2_OFF_BMAX = W2_OFF_BIAS + WA_BIAS_ROWS * WA_BIAS_PITCH * 2;
constexpr int W2_OFF_BAR = W2_OFF_BIAS + WA_BIAS_COPY_BYTES;  // 24 mbarriers, TMEM slot (+192), watchdog flag (+200)
constexpr int W2_OFF_X = W2_OFF_BIAS + WA_BIAS_HEAD_BYTES;    // float: max0 [2][2][128], sum0 [2][2][128], max1 [2][8][32], sum1 [2][8][32]
constexpr int W2_SMEM = W2_OFF_X + 4 * 2048;
static_assert(W2_SMEM <= 227 * 1024, "ping-pong window attention shared-memory budget");
static_assert(W2_OFF_P1 + WA_P1_BYTES + 128 * WA_KEYS * 2 <= W2_SMEM, "row tile 1's A operand must stay inside shared memory");
constexpr int W2_TM_O = 320;  // O[stream][tile] at W2_TM_O + 64 stream + 32 tile; S[stream] at 160 stream

struct W2Bars {
  uint64_t *qk_full, *qk_empty, *v_full, *v_empty;  // [stream * 2 + buffer]
  uint64_t *s_full, *p_full;                        // [stream]
  uint64_t* o_full;                                 // [stream * 2 + tile]
};

__device__ __forceinline__ void tmem_st_32x4(uint32_t taddr, const uint32_t* v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]),
               "r"(v[3])
               : "memory");
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
// D[tmem] (+)= A[tmem] * B[smem]: A = 128 lanes x 8 columns of packed bf16 pairs per K = 16 step
__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}

// mbarrier wait of the ping-pong kernel: parked in production; with the profiling hook armed a watchdog (see timed_wait)
template <bool PROF>
__device__ __forceinline__ void wait2(uint8_t* smem, long long* prof, uint64_t* bar, uint32_t parity, int slot, int item) {
  if (!PROF) {
    mbar_wait_parked(bar, parity);
    return;
  }
  volatile int* abort_flag = reinterpret_cast<volatile int*>(smem + W2_OFF_BAR + 200);
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (*abort_flag) break;
    if (clock64() - t0 > 100000000LL) {
      *abort_flag = 1;
      prof[196 + (threadIdx.x >> 5)] = (static_cast<long long>(blockIdx.x) << 40) | (static_cast<long long>(slot + 1) << 32) |
                                       static_cast<unsigned>(item);
      break;
    }
  }
  if ((threadIdx.x & 31) == 0 && blockIdx.x == 0) prof[(threadIdx.x >> 5) * 8 + slot] += clock64() - t0;
}

// eight probabilities of one 8-column key group: t = s * scale*log2e + bias + cg; returns the packed bf16 pairs and adds to l
__device__ __forceinline__ void prob_group(const float* s, uint4 b4, float cg, float2 sc2, uint32_t (&w)[4], float2& l) {
  const float2 cg2 = make_float2(cg, cg);
  const uint32_t bw[4] = {b4.x, b4.y, b4.z, b4.w};
  float2 p[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const float2 t = ffma2(make_float2(s[2 * e], s[2 * e + 1]), sc2, fadd2(bf16x2_to_f32x2(bw[e]), cg2));
    p[e] = make_float2(ex2_approx(t.x), ex2_approx(t.y));
    w[e] = pack_bf16x2(p[e].x, p[e].y);
  }
  l = fadd2(l, fadd2(fadd2(p[0], p[1]), fadd2(p[2], p[3])));
}

template <bool PROF>
__device__ __forceinline__ void softmax_stream(uint8_t* smem_, const W2Bars& br, uint32_t tmem_base, int s, int q, int c,
                                               const WaItemCtx& cx, long long* prof) {
  extern __shared__ __align__(1024) uint8_t smem[];  // same window as smem_; keeps the accesses below on LDS / STS
  const int lane = threadIdx.x & 31;
  const StageGeom& g = cx.g;
  const float MASK_L2 = -100.0f * 1.4426950408889634f;
  const float2 sc2 = make_float2(cx.scale_log2e, cx.scale_log2e);
  float* xch = reinterpret_cast<float*>(smem + W2_OFF_X);
  float* max0 = xch + s * 256;                // [c][128]
  float* sum0 = xch + 512 + s * 256;          // [c][128]
  float* max1 = xch + 1024 + s * 256 + lane;  // [8][32]
  float* sum1 = xch + 1536 + s * 256 + lane;  // [8][32]
  const uint32_t lane_addr = static_cast<uint32_t>(q * 32) << 16;
  const uint32_t tm_s = tmem_base + lane_addr + s * 160;
  const uint32_t tm_o = tmem_base + lane_addr + W2_TM_O + s * 64;
  const bf16* bias_tab = reinterpret_cast<const bf16*>(smem + W2_OFF_BIAS);
  const float* bmax_tab = reinterpret_cast<const float*>(smem + W2_OFF_BMAX);
  const int nw = g.W / g.ww, nh = g.H / g.wh;
  const int lw = 31 - __clz(nw);
  // ---- row tile 0: slot 32 q + lane, key groups [10 c, 10 c + 10)
  const int slot0 = q * 32 + lane;
  const int tok0 = slot_token_377(slot0);
  const bool valid0 = tok0 >= 0;
  const int brow0 = valid0 ? slot0 : 0;
  const int cls0 = (slot0 >= 48) + (slot0 >= 88);
  uint32_t mh0 = 0, mw0 = 0;
#pragma unroll
  for (int k = 0; k < 10; ++k) {
    const int kc = key_group_class_377(c * 10 + k);
    mh0 |= static_cast<uint32_t>((kc >> 1) != (cls0 >> 1)) << k;
    mw0 |= static_cast<uint32_t>((kc & 1) != (cls0 & 1)) << k;
  }
  const int t0w = valid0 ? tok0 : 0;
  const int d0 = t0w / 49, h0 = (t0w / 7) % 7, w0 = t0w % 7;
  // ---- row tile 1: slot 128 + lane (class 3), key groups [5 q + 3 c, + (c ? 2 : 3))
  const int slot1 = 128 + lane;
  const int tok1 = slot_token_377(slot1);
  const bool valid1 = tok1 >= 0;
  const int brow1 = valid1 ? slot1 : 0;
  const int g1 = q * 5 + (c ? 3 : 0), n1 = c ? 2 : 3;
  uint32_t mh1 = 0, mw1 = 0;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const int kc = key_group_class_377(min(g1 + k, 19));
    mh1 |= static_cast<uint32_t>((kc >> 1) != 1) << k;
    mw1 |= static_cast<uint32_t>((kc & 1) != 1) << k;
  }
  const int t1w = valid1 ? tok1 : 0;
  const int d1 = t1w / 49, h1 = (t1w / 7) % 7, w1 = t1w % 7;
  uint8_t* p1_row = smem + W2_OFF_P1 + s * WA_P1_BYTES + core_off(lane, g1, WA_KEYS / 8);
  const int bar0 = 1 + s * 4 + q, bar1 = 9 + s;

  uint32_t dst0 = 0, dst1_prev = 0;

  // epilogue of row tile 0 (all eight warps: 16 of the 32 dims each) / row tile 1 (lane quarter 0 only)
  auto epilogue0 = [&](int n) {
    wait2<PROF>(smem, prof, br.o_full + s * 2 + 0, n & 1, 1, n);
    tcgen05_fence_after();
    uint32_t o16[16];
    tmem_ld_32x16(tm_o + c * 16, o16);
    const float inv = 1.0f / (sum0[slot0] + sum0[128 + slot0]);
    tmem_ld_wait();
    tcgen05_fence_before();
    if (valid0) {
      uint4* dst = reinterpret_cast<uint4*>(cx.out + dst0 + c * 16);
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        uint4 o;
        o.x = pack_bf16x2(__uint_as_float(o16[8 * k + 0]) * inv, __uint_as_float(o16[8 * k + 1]) * inv);
        o.y = pack_bf16x2(__uint_as_float(o16[8 * k + 2]) * inv, __uint_as_float(o16[8 * k + 3]) * inv);
        o.z = pack_bf16x2(__uint_as_float(o16[8 * k + 4]) * inv, __uint_as_float(o16[8 * k + 5]) * inv);
        o.w = pack_bf16x2(__uint_as_float(o16[8 * k + 6]) * inv, __uint_as_float(o16[8 * k + 7]) * inv);
        dst[k] = o;
      }
    }
  };
  auto epilogue1 = [&](int n) {
    wait2<PROF>(smem, prof, br.o_full + s * 2 + 1, n & 1, 2, n);
    if (q != 0) return;
    tcgen05_fence_after();
    uint32_t o16[16];
    tmem_ld_32x16(tm_o + 32 + c * 16, o16);
    float l = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) l += sum1[k * 32];
    const float inv = 1.0f / l;
    tmem_ld_wait();
    tcgen05_fence_before();
    if (valid1) {
      uint4* dst = reinterpret_cast<uint4*>(cx.out + dst1_prev + c * 16);
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        uint4 o;
        o.x = pack_bf16x2(__uint_as_float(o16[8 * k + 0]) * inv, __uint_as_float(o16[8 * k + 1]) * inv);
        o.y = pack_bf16x2(__uint_as_float(o16[8 * k + 2]) * inv, __uint_as_float(o16[8 * k + 3]) * inv);
        o.z = pack_bf16x2(__uint_as_float(o16[8 * k + 4]) * inv, __uint_as_float(o16[8 * k + 5]) * inv);
        o.w = pack_bf16x2(__uint_as_float(o16[8 * k + 6]) * inv, __uint_as_float(o16[8 * k + 7]) * inv);
        dst[k] = o;
      }
    }
  };

  int head = cx.u_lo / cx.n_items, item = cx.u_lo - head * cx.n_items;
  int seg = item / cx.nwin, win = item - seg * cx.nwin;
  int head_loaded = -1, n_done = 0;
  for (int j = 0; j < cx.n_my; ++j) {
    // both streams walk the whole unit list, so that a head change reloads the bias table at the same point for all 16 warps
    if (head != head_loaded) {
      asm volatile("bar.sync 11, 512;" ::: "memory");
      const uint4* src = reinterpret_cast<const uint4*>(reinterpret_cast<const uint8_t*>(cx.bias_dense) +
                                                        static_cast<size_t>(head) * WA_BIAS_HEAD_BYTES);
      uint4* dstb = reinterpret_cast<uint4*>(smem + W2_OFF_BIAS);
      for (int i = threadIdx.x; i < WA_BIAS_COPY_BYTES / 16; i += 512) dstb[i] = __ldg(src + i);
      asm volatile("bar.sync 11, 512;" ::: "memory");
      head_loaded = head;
    }
    if ((j & 1) == s) {
      const int n = j >> 1;  // unit index inside the stream
      const int wy = (win >> lw) & (nh - 1), wx = win & (nw - 1);
      const bool bh = cx.shifted && wy == nh - 1 && g.sh, bw_ = cx.shifted && wx == nw - 1 && g.sw;
      const int row_base = (seg * cx.T) * cx.C + head * 32;
      // ================= row tile 0
      {
        const uint32_t mbits = (bh ? mh0 : 0u) | (bw_ ? mw0 : 0u);
        {
          int y = wy * 7 + g.sh + h0, x = wx * 7 + g.sw + w0;
          if (y >= g.H) y -= g.H;
          if (x >= g.W) x -= g.W;
          dst0 = static_cast<uint32_t>(row_base + ((d0 * g.H + y) * g.W + x) * cx.C);
        }
        wait2<PROF>(smem, prof, br.s_full + s, 0, 0, j);
        tcgen05_fence_after();
        float sc[WA_QCOLS];
        uint32_t* raw = reinterpret_cast<uint32_t*>(sc);
        float mx = -INFINITY;
#pragma unroll
        for (int k = 0; k < 2; ++k) {  // pass A: raw maximum of the own 80 columns
          tmem_ld_32x32(tm_s + c * 80 + k * 40, raw);
          tmem_ld_32x8(tm_s + c * 80 + k * 40 + 32, raw + 32);
          tmem_ld_wait();
#pragma unroll
          for (int e = 0; e < WA_QCOLS; ++e) mx = fmaxf(mx, sc[e]);
        }
        max0[c * 128 + slot0] = mx;
        asm volatile("bar.sync %0, 64;" ::"r"(bar0) : "memory");
        mx = fmaxf(mx, max0[(c ^ 1) * 128 + slot0]);
        const float nbound = -fmaf(mx, cx.scale_log2e, bmax_tab[brow0]);
        const bf16* bias_row = bias_tab + brow0 * WA_BIAS_PITCH;
        const int bsw = (brow0 >> 1) & 3;
        float2 l = make_float2(0.f, 0.f);
#pragma unroll
        for (int k = 0; k < 2; ++k) {  // pass B: probabilities, written over the consumed score columns
          tmem_ld_32x32(tm_s + c * 80 + k * 40, raw);
          tmem_ld_32x8(tm_s + c * 80 + k * 40 + 32, raw + 32);
          tmem_ld_wait();
#pragma unroll
          for (int gq = 0; gq < 5; ++gq) {
            const int lg = k * 5 + gq;
            const uint4 b4 = *reinterpret_cast<const uint4*>(bias_row + ((c * 10 + lg) ^ bsw) * 8);
            const float cg = ((mbits >> lg) & 1u) ? nbound + MASK_L2 : nbound;
            uint32_t w[4];
            prob_group(sc + gq * 8, b4, cg, sc2, w, l);
            tmem_st_32x4(tm_s + c * 80 + lg * 4, w);
          }
        }
        sum0[c * 128 + slot0] = l.x + l.y;
        tmem_st_wait();
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(br.p_full + s);
        if (n > 0) epilogue1(n - 1);  // row tile 1 of the stream's previous unit
      }
      // ================= row tile 1
      {
        const uint32_t mbits = (bh ? mh1 : 0u) | (bw_ ? mw1 : 0u);
        uint32_t dst1;
        {
          int y = wy * 7 + g.sh + h1, x = wx * 7 + g.sw + w1;
          if (y >= g.H) y -= g.H;
          if (x >= g.W) x -= g.W;
          dst1 = static_cast<uint32_t>(row_base + ((d1 * g.H + y) * g.W + x) * cx.C);
        }
        wait2<PROF>(smem, prof, br.s_full + s, 1, 3, j);
        tcgen05_fence_after();
        float sc[24];
        uint32_t* raw = reinterpret_cast<uint32_t*>(sc);
        tmem_ld_32x16(tm_s + g1 * 8, raw);
        if (c == 0) tmem_ld_32x8(tm_s + g1 * 8 + 16, raw + 16);
        tmem_ld_wait();
        tcgen05_fence_before();
        float mx = sc[0];
#pragma unroll
        for (int e = 1; e < 16; ++e) mx = fmaxf(mx, sc[e]);
        if (c == 0) {
#pragma unroll
          for (int e = 16; e < 24; ++e) mx = fmaxf(mx, sc[e]);
        }
        max1[(q * 2 + c) * 32] = mx;
        asm volatile("bar.sync %0, 256;" ::"r"(bar1) : "memory");
#pragma unroll
        for (int k = 0; k < 8; ++k) mx = fmaxf(mx, max1[k * 32]);
        const float nbound = -fmaf(mx, cx.scale_log2e, bmax_tab[brow1]);
        const bf16* bias_row = bias_tab + brow1 * WA_BIAS_PITCH;
        const int bsw = (brow1 >> 1) & 3;
        float2 l = make_float2(0.f, 0.f);
#pragma unroll
        for (int gq = 0; gq < 3; ++gq) {
          if (gq < n1) {
            const uint4 b4 = *reinterpret_cast<const uint4*>(bias_row + ((g1 + gq) ^ bsw) * 8);
            const float cg = ((mbits >> gq) & 1u) ? nbound + MASK_L2 : nbound;
            uint32_t w[4];
            prob_group(sc + gq * 8, b4, cg, sc2, w, l);
            *reinterpret_cast<uint4*>(p1_row + gq * 128) = make_uint4(w[0], w[1], w[2], w[3]);
          }
        }
        sum1[(q * 2 + c) * 32] = l.x + l.y;
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(br.p_full + s);
        epilogue0(n);  // row tile 0 of this unit
        dst1_prev = dst1;
      }
      n_done = n + 1;
    }
    if (++win == cx.nwin) {
      win = 0;
      if (++seg * cx.nwin == cx.n_items) { seg = 0; ++head; }
    }
  }
  if (n_done > 0) epilogue1(n_done - 1);
}

template <bool PROF>
__global__ void __launch_bounds__(W2_THREADS, 1)
window_attention_pp_kernel(const __grid_constant__ WaMaps maps, bf16* __restrict__ out, const bf16* __restrict__ bias_dense,
                           StageGeom g, int n_seg, int C, int n_heads, float scale_log2e, long long* prof) {
  extern __shared__ __align__(1024) uint8_t smem[];
  W2Bars br;
  br.qk_full = reinterpret_cast<uint64_t*>(smem + W2_OFF_BAR);  // [4]
  br.qk_empty = br.qk_full + 4;
  br.v_full = br.qk_empty + 4;
  br.v_empty = br.v_full + 4;
  br.s_full = br.v_empty + 4;  // [2]
  br.p_full = br.s_full + 2;   // [2]
  br.o_full = br.p_full + 2;   // [4]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(br.o_full + 4);  // byte 192; watchdog flag at byte 200

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nwin = windows_per_segment(g);
  const int T = g.D * g.H * g.W;
  const int n_items = n_seg * nwin;
  const bool shifted = (g.sd | g.sh | g.sw) != 0;
  const long long n_units = static_cast<long long>(n_heads) * n_items;
  const int u_lo = static_cast<int>(n_units * blockIdx.x / gridDim.x);
  const int u_hi = static_cast<int>(n_units * (blockIdx.x + 1) / gridDim.x);
  const int n_my = u_hi - u_lo;
  if (PROF && blockIdx.x == 0 && tid == 0) prof[24 * 8 + 2] = clock64();
  if ((smem_u32(smem) & 1023u) != 0) __trap();

  for (int i = tid; i < W2_OFF_BIAS / 16; i += W2_THREADS) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (warp == W2_WARP_MMA && lane == 0) {
    *reinterpret_cast<volatile int*>(smem + W2_OFF_BAR + 200) = 0;
    for (int k = 0; k < 4; ++k) {
      mbar_init(&br.qk_full[k], 1);
      mbar_init(&br.qk_empty[k], 1);
      mbar_init(&br.v_full[k], 1);
      mbar_init(&br.v_empty[k], 1);
      mbar_init(&br.o_full[k], 1);
    }
    for (int k = 0; k < 2; ++k) {
      mbar_init(&br.s_full[k], 1);
      mbar_init(&br.p_full[k], 8);
    }
    fence_barrier_init();
  }
  if (warp == W2_WARP_LOADER && lane == 0) {
#pragma unroll
    for (int k = 0; k < 4; ++k) tma_prefetch_desc(&maps.m[k]);
  }
  if (warp == W2_WARP_TMEM) tmem_alloc(tmem_slot, 512);
  fence_proxy_async_smem();
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == W2_WARP_LOADER) {
    if (lane == 0) {
      const int lw = 31 - __clz(g.W / 7);
      const int nh_mask = g.H / 7 - 1, nw_mask = g.W / 7 - 1;
      const uint32_t smem0 = smem_u32(smem);
      int head = u_lo / n_items, item = u_lo - head * n_items;
      int seg = item / nwin, win = item - seg * nwin;
      for (int j = 0; j < n_my; ++j) {
        const int s = j & 1, n = j >> 1, k = s * 2 + (n & 1);
        const uint32_t par = ((n >> 1) & 1) ^ 1;
        int ya = ((win >> lw) & nh_mask) * 7 + g.sh, xa = (win & nw_mask) * 7 + g.sw;
        int yb = ya + 4, xb = xa + 4;
        if (yb >= g.H) yb -= g.H;
        if (xb >= g.W) xb -= g.W;
        const int cq = head * 32, ds = seg * 3;
        const uint32_t sq = smem0 + W2_OFF_STAGE + k * WA_STAGE_BYTES, sk = sq + WA_Q_BYTES, sv = sk + WA_K_BYTES;
        wait2<PROF>(smem, prof, &br.qk_empty[k], par, 0, j);
        mbar_expect_tx(&br.qk_full[k], WA_QK_TX_BYTES);
        tma_load_4d(sk + 0 * 64, &maps.m[0], &br.qk_full[k], C + cq, xa, ya, ds);
        tma_load_4d(sk + 48 * 64, &maps.m[1], &br.qk_full[k], C + cq, xb, ya, ds);
        tma_load_4d(sk + 88 * 64, &maps.m[2], &br.qk_full[k], C + cq, xa, yb, ds);
        tma_load_4d(sk + 128 * 64, &maps.m[3], &br.qk_full[k], C + cq, xb, yb, ds);
        tma_load_4d(sq + 0 * 64, &maps.m[0], &br.qk_full[k], cq, xa, ya, ds);
        tma_load_4d(sq + 48 * 64, &maps.m[1], &br.qk_full[k], cq, xb, ya, ds);
        tma_load_4d(sq + 88 * 64, &maps.m[2], &br.qk_full[k], cq, xa, yb, ds);
#pragma unroll
        for (int r = 0; r < 4; ++r) tma_load_4d(sq + (128 + 32 * r) * 64, &maps.m[3], &br.qk_full[k], cq, xb, yb, ds);
        wait2<PROF>(smem, prof, &br.v_empty[k], par, 3, j);
        mbar_expect_tx(&br.v_full[k], WA_V_TX_BYTES);
        tma_load_4d(sv + 0 * 64, &maps.m[0], &br.v_full[k], 2 * C + cq, xa, ya, ds);
        tma_load_4d(sv + 48 * 64, &maps.m[1], &br.v_full[k], 2 * C + cq, xb, ya, ds);
        tma_load_4d(sv + 88 * 64, &maps.m[2], &br.v_full[k], 2 * C + cq, xa, yb, ds);
        tma_load_4d(sv + 128 * 64, &maps.m[3], &br.v_full[k], 2 * C + cq, xb, yb, ds);
        if (++win == nwin) {
          win = 0;
          if (++seg * nwin == n_items) { seg = 0; ++head; }
        }
      }
    }
  } else if (warp == W2_WARP_MMA || warp == W2_WARP_MMA + 1) {
    const int s = warp - W2_WARP_MMA;
    const int n_s = (n_my + 1 - s) >> 1;  // units of this stream
    if (lane == 0 && n_s > 0) {
      constexpr uint32_t idesc_s = umma_idesc_bf16(128, WA_KEYS);
      constexpr uint32_t idesc_o = umma_idesc_bf16(128, 32) | (1u << 16);  // B (= v) MN-major
      const uint32_t smem0 = smem_u32(smem);
      const uint32_t tm_s = tmem_base + s * 160, tm_o = tmem_base + W2_TM_O + s * 64;
      const uint64_t dp1 = umma_desc_nosw(smem0 + W2_OFF_P1 + s * WA_P1_BYTES, 128, (WA_KEYS / 8) * 128);
      auto issue_s = [&](int t) {  // S of tile t: both row tiles of a unit read the same staged q / k
        const int n = t >> 1, k = s * 2 + (n & 1);
        const uint32_t b = smem0 + W2_OFF_STAGE + k * WA_STAGE_BYTES;
        if ((t & 1) == 0) {
          wait2<PROF>(smem, prof, &br.qk_full[k], (n >> 1) & 1, 1, t);
          tcgen05_fence_after();
        }
        const uint64_t dq = umma_desc_sw64(b + (t & 1) * (128 * 64)), dk = umma_desc_sw64(b + WA_Q_BYTES);
#pragma unroll
        for (int kk = 0; kk < 2; ++kk) umma_bf16_ss(tm_s, dq + ((kk * 32) >> 4), dk + ((kk * 32) >> 4), idesc_s, kk);
        umma_commit(br.s_full + s);
        if (t & 1) umma_commit(&br.qk_empty[k]);
      };
      issue_s(0);
      for (int t = 0; t < 2 * n_s; ++t) {
        const int n = t >> 1, k = s * 2 + (n & 1), tile = t & 1;
        wait2<PROF>(smem, prof, br.p_full + s, tile, 2, t);
        if (tile == 0) wait2<PROF>(smem, prof, &br.v_full[k], (n >> 1) & 1, 3, t);
        tcgen05_fence_after();
        const uint64_t dv = umma_desc_sw64(smem0 + W2_OFF_STAGE + k * WA_STAGE_BYTES + WA_Q_BYTES + WA_K_BYTES);
        const uint32_t tm_ot = tm_o + tile * 32;
        if (tile == 0) {
#pragma unroll
          for (int kk = 0; kk < WA_KEYS / 16; ++kk)  // P of keys [80 c, 80 c + 80) sits in columns [80 c, 80 c + 40)
            umma_bf16_ts(tm_ot, tm_s + (kk < 5 ? 8 * kk : 80 + 8 * (kk - 5)), dv + ((kk * 1024) >> 4), idesc_o, kk);
        } else {
#pragma unroll
          for (int kk = 0; kk < WA_KEYS / 16; ++kk) umma_bf16_ss(tm_ot, dp1 + ((kk * 256) >> 4), dv + ((kk * 1024) >> 4), idesc_o, kk);
        }
        umma_commit(&br.o_full[s * 2 + tile]);
        if (tile) umma_commit(&br.v_empty[k]);
        if (t + 1 < 2 * n_s) {
          // the S columns hold P of this tile until P v has read them (tile 0), and every warp has consumed S (tile 1):
          // both are implied by the completion of this tile's P v
          wait2<PROF>(smem, prof, &br.o_full[s * 2 + tile], n & 1, 4, t);
          tcgen05_fence_after();
          issue_s(t + 1);
        }
      }
    }
  } else if (warp < 16) {
    WaItemCtx cx;
    cx.bias_dense = bias_dense; cx.out = out; cx.g = g; cx.n_items = n_items; cx.nwin = nwin; cx.T = T; cx.C = C;
    cx.u_lo = u_lo; cx.n_my = n_my; cx.scale_log2e = scale_log2e; cx.shifted = shifted;
    softmax_stream<PROF>(smem, br, tmem_base, warp >> 3, warp & 3, (warp >> 2) & 1, cx, prof);
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == W2_WARP_TMEM) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
  if (PROF && blockIdx.x == 0 && tid == 0) {
    prof[24 * 8] = n_my;
    prof[24 * 8 + 1] = clock64() - prof[24 * 8 + 2];
  }
}

}  // namespace lrce
