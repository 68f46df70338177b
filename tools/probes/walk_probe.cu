// walk_probe.cu — microbenchmarks that decided the design of the cluster-sharded encoder walk (csrc/encoder_walk.cu).
// Build: tools/probes/build.sh ; run on a B200: tools/probes/walk_probe [m64|xchg|stream ...]
//  (1) m64    : TMEM row -> lane layout of tcgen05.mma cta_group::1 with M = 64, N = 8 (swap-AB tiles of the walk)
//  (2) xchg   : 16-CTA clusters — how many are co-resident with a ~220 KB CTA, and the latency of one all-to-all exchange done
//               as shared::cta -> shared::cluster bulk copies that complete_tx on the receiver's mbarrier
//  (3) stream : per-SM TMA ingest when NC clusters x 16 CTAs each pull "their" 1/16 of the same 14 MB per layer from L2 / HBM
//               through a ring of 24 KB stages (what bounds the walk once the dependency chain is hidden)
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "../../vqa-lrce-kbs-2023_b200/csrc/host_common.h"
#include "../../vqa-lrce-kbs-2023_b200/csrc/lrce_common.cuh"

using namespace lrce;

#define CK(x)                                                                                  \
  do {                                                                                         \
    cudaError_t e_ = (x);                                                                      \
    if (e_ != cudaSuccess) {                                                                   \
      printf("CUDA error %s at %s:%d: %s\n", cudaGetErrorName(e_), __FILE__, __LINE__, cudaGetErrorString(e_)); \
      exit(1);                                                                                 \
    }                                                                                          \
  } while (0)

// ------------------------------------------------------------------------------------------------ (1) M = 64 layout
__device__ __forceinline__ void tmem_st_32x8(uint32_t taddr, const uint32_t* v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(v[0]), "r"(v[1]),
               "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_ld_32x8p(uint32_t taddr, uint32_t* v) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(taddr)
               : "memory");
}

__global__ void __launch_bounds__(128) probe_m64(float* out, int M) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 20480);
  uint32_t* slot = reinterpret_cast<uint32_t*>(smem + 20480 + 16);
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 20480 / 16; i += 128) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  __syncthreads();
  // A[r][k = 0] = r + 1 (128B swizzle: 16-byte chunk c of row r sits at chunk position c ^ (r & 7))
  *reinterpret_cast<bf16*>(smem + (tid / 8) * 1024 + (tid % 8) * 128 + ((tid % 8) * 16)) = __float2bfloat16(float(tid + 1));
  if (tid < 8) *reinterpret_cast<bf16*>(smem + 16384 + tid * 128 + tid * 16) = __float2bfloat16(float(tid + 1));
  fence_proxy_async_smem();
  if (warp == 0) tmem_alloc(slot, 32);
  if (tid == 0) {
    mbar_init(bar, 1);
    fence_barrier_init();
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = *slot;
  uint32_t neg[8];
  for (int i = 0; i < 8; ++i) neg[i] = __float_as_uint(-1.0f);
  tmem_st_32x8(tmem + (static_cast<uint32_t>(warp * 32) << 16), neg);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  if (tid == 0) {
    const uint32_t idesc = umma_idesc_bf16(M, 8);
    for (int k = 0; k < 4; ++k)
      umma_bf16_ss(tmem, umma_desc_k_sw128(smem_u32(smem) + k * 32), umma_desc_k_sw128(smem_u32(smem + 16384) + k * 32), idesc, k);
    umma_commit(bar);
  }
  mbar_wait_parked(bar, 0);
  tcgen05_fence_after();
  uint32_t v[8];
  tmem_ld_32x8p(tmem + (static_cast<uint32_t>(warp * 32) << 16), v);
  tmem_ld_wait();
  for (int i = 0; i < 8; ++i) out[tid * 8 + i] = __uint_as_float(v[i]);
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 32);
}

static void run_m64() {
  float* d;
  CK(cudaMalloc(&d, 128 * 8 * 4));
  CK(cudaFuncSetAttribute(probe_m64, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768));
  for (int M : {128, 64}) {
    CK(cudaMemset(d, 0, 128 * 8 * 4));
    probe_m64<<<1, 128, 32768>>>(d, M);
    CK(cudaDeviceSynchronize());
    std::vector<float> h(128 * 8);
    CK(cudaMemcpy(h.data(), d, 128 * 8 * 4, cudaMemcpyDeviceToHost));
    printf("[m64] M=%d: lane -> row (from column 0; '.' = untouched), consistent = all 8 columns equal (row+1)(n+1)\n", M);
    for (int l = 0; l < 128; ++l) {
      const float r1 = h[l * 8];
      bool ok = r1 > 0;
      for (int n = 0; n < 8 && ok; ++n) ok = h[l * 8 + n] == r1 * (n + 1);
      if (r1 == -1.0f) printf(" .");
      else printf(" %d%s", int(r1) - 1, ok ? "" : "?");
      if (l % 32 == 31) printf("\n");
    }
  }
  cudaFree(d);
}

// ------------------------------------------------------------------------------------------------ (2) cluster exchange
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void bulk_s2c(uint32_t dst_cluster, uint32_t src_cta, uint32_t bytes, uint32_t bar_cluster) {
  asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_cluster),
               "r"(src_cta), "r"(bytes), "r"(bar_cluster)
               : "memory");
}

constexpr int XB = 768;  // bytes per (source, exchange): 4 rows x 48 features fp32
template <int CL>
__global__ void __launch_bounds__(384, 1) probe_xchg(int iters, unsigned long long* times, int* errors) {
  extern __shared__ __align__(1024) uint8_t smem[];
  float* G = reinterpret_cast<float*>(smem);                         // [2][CL][192]
  float* stg = reinterpret_cast<float*>(smem + 2 * CL * XB);        // [2][192]
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 2 * CL * XB + 2 * XB);  // [2]
  const uint32_t rank = cluster_ctarank();
  const int tid = threadIdx.x;
  if (tid == 0) {
    mbar_init(&bar[0], 1);
    mbar_init(&bar[1], 1);
    fence_barrier_init();
  }
  cluster_sync_all();
  unsigned long long t0 = 0;
  if (tid == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  int bad = 0;
  for (int it = 0; it < iters; ++it) {
    const int b = it & 1;
    if (tid < 192) stg[b * 192 + tid] = float(it * 16 + int(rank));
    fence_proxy_async_smem();
    __syncthreads();
    if (tid == 0) {
      mbar_expect_tx(&bar[b], CL * XB);
      for (uint32_t d = 0; d < CL; ++d)
        bulk_s2c(mapa(smem_u32(G + (b * CL + rank) * 192), d), smem_u32(stg + b * 192), XB, mapa(smem_u32(&bar[b]), d));
    }
    mbar_wait_parked(&bar[b], (it >> 1) & 1);
    if (tid < 192)
      for (int s = 0; s < CL; ++s) bad += G[(b * CL + s) * 192 + tid] != float(it * 16 + s);
  }
  if (tid == 0) {
    unsigned long long t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    times[blockIdx.x] = t1 - t0;
  }
  if (bad) atomicAdd(errors, bad);
  cluster_sync_all();
}

template <int CL>
static void run_xchg(int smem_bytes) {
  auto kern = probe_xchg<CL>;
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
  if (CL > 8) CK(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(CL);
  cfg.blockDim = dim3(384);
  cfg.dynamicSmemBytes = smem_bytes;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CL;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  int ncl = 0;
  cudaError_t e = cudaOccupancyMaxActiveClusters(&ncl, kern, &cfg);
  printf("[xchg] cluster size %d, smem %d B: cudaOccupancyMaxActiveClusters = %d (%s)\n", CL, smem_bytes, ncl, cudaGetErrorString(e));
  if (e != cudaSuccess || ncl < 1) {
    cudaGetLastError();
    return;
  }
  unsigned long long* times;
  int* errors;
  CK(cudaMalloc(&times, 256 * 8));
  CK(cudaMalloc(&errors, 4));
  for (int nc : {1, ncl}) {
    CK(cudaMemset(errors, 0, 4));
    cfg.gridDim = dim3(CL * nc);
    const int iters = 2000;
    e = cudaLaunchKernelEx(&cfg, kern, iters, times, errors);
    if (e != cudaSuccess) {
      printf("[xchg] launch failed: %s\n", cudaGetErrorString(e));
      cudaGetLastError();
      break;
    }
    CK(cudaDeviceSynchronize());
    std::vector<unsigned long long> h(CL * nc);
    int herr = 0;
    CK(cudaMemcpy(h.data(), times, CL * nc * 8, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(&herr, errors, 4, cudaMemcpyDeviceToHost));
    unsigned long long mx = 0;
    for (auto t : h) mx = t > mx ? t : mx;
    printf("[xchg] %d clusters x %d CTAs: %.1f ns per all-to-all exchange of %d B per pair (errors %d)\n", nc, CL, double(mx) / iters,
           XB, herr);
  }
  cudaFree(times);
  cudaFree(errors);
}

// ------------------------------------------------------------------------------------------------ (3) weight stream
constexpr int ST_ROWS = 192, ST_BYTES = ST_ROWS * 128, ROWS_PER_LAYER_RANK = 6336, STAGES_PER_LAYER = 33;
__device__ __forceinline__ void tma_prefetch_2d(const CUtensorMap* tm, int c0, int c1) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global [%0, {%1, %2}];" ::"l"(reinterpret_cast<uint64_t>(tm)), "r"(c0), "r"(c1)
               : "memory");
}
template <int CL>
__global__ void __launch_bounds__(128, 1)
probe_stream(const __grid_constant__ CUtensorMap tm, int n_layers, int passes, int nstage, int ahead, int touch,
             unsigned long long* times, float* sink) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + nstage * ST_BYTES);
  uint64_t* empty = full + 8;
  const uint32_t rank = cluster_ctarank();
  const int tid = threadIdx.x;
  if (tid == 0) {
    for (int s = 0; s < nstage; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    fence_barrier_init();
  }
  cluster_sync_all();
  unsigned long long t0 = 0;
  if (tid == 32) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  const int total = passes * n_layers * STAGES_PER_LAYER;
  auto row_of = [&](int i) {
    const int n = (i / STAGES_PER_LAYER) % n_layers, st = i % STAGES_PER_LAYER;
    return (n * CL + int(rank)) * ROWS_PER_LAYER_RANK + st * ST_ROWS;
  };
  if (tid == 0) {
    for (int i = 0; i < total; ++i) {
      const int s = i % nstage;
      mbar_wait_parked(&empty[s], ((i / nstage) & 1) ^ 1);
      mbar_expect_tx(&full[s], ST_BYTES);
      tma_load_2d(smem + s * ST_BYTES, &tm, &full[s], 0, row_of(i));
      if (ahead > 0 && i + ahead < total) tma_prefetch_2d(&tm, 0, row_of(i + ahead));
    }
  } else if (tid == 32) {
    float acc = 0.f;
    for (int i = 0; i < total; ++i) {
      const int s = i % nstage;
      mbar_wait_parked(&full[s], (i / nstage) & 1);
      if (touch) acc += *reinterpret_cast<const float*>(smem + s * ST_BYTES + (i & 63) * 64);
      mbar_arrive(&empty[s]);
    }
    unsigned long long t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    times[blockIdx.x] = t1 - t0;
    if (acc == 123.456f) *sink = acc;
  }
  cluster_sync_all();
}

template <int CL>
static void run_stream() {
  const int n_layers = 12;
  const size_t rows = size_t(n_layers) * CL * ROWS_PER_LAYER_RANK;
  bf16* buf;
  CK(cudaMalloc(&buf, rows * 128));
  CK(cudaMemset(buf, 0, rows * 128));
  CUtensorMap tm;
  if (make_tmap_2d_bf16(&tm, buf, 64, rows, 64, 64, ST_ROWS) != LRCE_OK) {
    printf("tensor map: %s\n", lrce_last_error());
    exit(1);
  }
  auto kern = probe_stream<CL>;
  unsigned long long* times;
  float* sink;
  CK(cudaMalloc(&times, 256 * 8));
  CK(cudaMalloc(&sink, 4));
  if (CL > 8) CK(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * ST_BYTES + 256));
  for (int nstage : {3, 4, 6}) {
    for (int ahead : {0, 8}) {
      for (int nc : {1, 2, 4, 6, 8, 9}) {
        if (nc * CL > 148) continue;
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(CL * nc);
        cfg.blockDim = dim3(128);
        // pad the dynamic smem so that only ONE CTA fits per SM, as in the real kernel
        cfg.dynamicSmemBytes = 8 * ST_BYTES + 256;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = CL;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        int ncl = 0;
        if (cudaOccupancyMaxActiveClusters(&ncl, kern, &cfg) != cudaSuccess || ncl < nc) {
          cudaGetLastError();
          continue;
        }
        const int passes = 3;
        for (int rep = 0; rep < 2; ++rep) {
          cudaError_t e = cudaLaunchKernelEx(&cfg, kern, tm, n_layers, passes, nstage, ahead, 1, times, sink);
          if (e != cudaSuccess) {
            printf("[stream] launch failed: %s\n", cudaGetErrorString(e));
            cudaGetLastError();
            break;
          }
          CK(cudaDeviceSynchronize());
        }
        std::vector<unsigned long long> h(CL * nc);
        CK(cudaMemcpy(h.data(), times, CL * nc * 8, cudaMemcpyDeviceToHost));
        unsigned long long mx = 0;
        for (auto t : h) mx = t > mx ? t : mx;
        const double bytes_cta = double(passes) * n_layers * STAGES_PER_LAYER * ST_BYTES;
        printf("[stream] CL=%d clusters=%d stages=%d prefetch-ahead=%d: %.1f us for %d layer-steps = %.2f us per layer-step, %.1f GB/s per SM, "
               "%.2f TB/s aggregate\n",
               CL, nc, nstage, ahead, mx / 1e3, passes * n_layers, mx / 1e3 / (passes * n_layers), bytes_cta / mx,
               bytes_cta * CL * nc / mx / 1e3);
      }
    }
  }
  cudaFree(buf);
  cudaFree(times);
  cudaFree(sink);
}

// ------------------------------------------------------------------------------------------------ (4) small-N MMA rate
// one elected lane issues `n` tcgen05.mma (M x N x 16, operands in shared memory) into `nacc` rotating accumulators
__global__ void __launch_bounds__(128) probe_mma_rate(long long* out, int M, int N, int n, int nacc, int a_stride) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 98304 + 16384);
  uint32_t* slot = reinterpret_cast<uint32_t*>(smem + 98304 + 16384 + 16);
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < (98304 + 16384) / 16; i += 128) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  fence_proxy_async_smem();
  if (warp == 0) tmem_alloc(slot, 512);
  if (tid == 0) {
    mbar_init(bar, 1);
    fence_barrier_init();
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = *slot;
  if (warp == 1) {
    const uint32_t idesc = umma_idesc_bf16(M, N);
    const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem + 98304);
    uint32_t pred;
    asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(pred));
    const uint32_t a_lo = (a0 >> 4) | 0x10000u, b_lo = (b0 >> 4) | 0x10000u;
    const uint32_t hi = (1024u >> 4) | (1u << 14) | (2u << 29);
    long long t0 = clock64();
    if (pred) {
#pragma unroll 1
      for (int i = 0; i < n; i += 16) {
#pragma unroll
        for (int m = 0; m < 16; ++m)
          asm volatile(
              "{\n.reg .pred p;\n.reg .b64 da, db;\nmov.b64 da, {%1, %4};\nmov.b64 db, {%2, %4};\nsetp.ne.b32 p, 1, 0;\n"
              "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, p;\n}\n" ::"r"(tmem + (m % nacc) * N),
              "r"(a_lo + (((m * a_stride) & 65535) >> 4)), "r"(b_lo + (m & 3) * 2), "r"(idesc), "r"(hi)
              : "memory");
      }
      umma_commit(bar);
    }
    __syncwarp();
    long long t1 = clock64();
    mbar_wait_parked(bar, 0);
    long long t2 = clock64();
    if (pred) {
      out[0] = t1 - t0;
      out[1] = t2 - t0;
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}
static void run_mma_rate() {
  long long* d;
  CK(cudaMalloc(&d, 16));
  CK(cudaFuncSetAttribute(probe_mma_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 98304 + 16384 + 256));
  const int n = 1024;
  for (int M : {64, 128})
    for (int N : {8, 16, 32, 64})
      for (int nacc : {1, 4})
        for (int a_stride : {0, 32, 8192}) {
          probe_mma_rate<<<1, 128, 98304 + 16384 + 256>>>(d, M, N, n, nacc, a_stride);
          CK(cudaDeviceSynchronize());
          long long h[2];
          CK(cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost));
          printf("[mma] M=%d N=%d accumulators=%d A stride %5d B: issue %.1f cyc/MMA, complete %.1f cyc/MMA\n", M, N, nacc, a_stride,
                 double(h[0]) / n, double(h[1]) / n);
        }
  cudaFree(d);
}

// ------------------------------------------------------------------------------------------------ (5) MMA under a TMA stream
// the fc1 loop of the walk in isolation: ring of `nstage` 24 KB slots, per slot 12 MMAs (M = 64, N = 8) on three 64-row
// tiles, commit releases the slot; mode 0: no TMA at all (operands stale), 1: TMA stream feeds the ring
__global__ void __launch_bounds__(128, 1)
probe_mma_stream(const __grid_constant__ CUtensorMap tm, int n_slots, int nstage, int mode, int mma_per_slot, long long* out,
                 int commit_every, int n_mma_warps) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* xb = smem + nstage * ST_BYTES;
  uint64_t* full = reinterpret_cast<uint64_t*>(xb + 12288);
  uint64_t* empty = full + 8;
  uint32_t* slot = reinterpret_cast<uint32_t*>(empty + 8);
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 12288 / 16; i += 128) reinterpret_cast<uint4*>(xb)[i] = make_uint4(0, 0, 0, 0);
  fence_proxy_async_smem();
  if (tid == 0) {
    for (int s = 0; s < nstage; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(slot, 64);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = *slot;
  if (warp == 0 && mode == 1) {
    if (tid == 0) {
      for (int i = 0; i < n_slots; ++i) {
        const int s = i % nstage;
        mbar_wait_parked(&empty[s], ((i / nstage) & 1) ^ 1);
        mbar_expect_tx(&full[s], ST_BYTES);
        tma_load_2d(smem + s * ST_BYTES, &tm, &full[s], 0, (blockIdx.x * 4096 + i) * ST_ROWS % 1000000);
      }
    }
  } else if (warp == 1 || (warp == 3 && n_mma_warps == 2)) {
    const int wsel = warp == 1 ? 0 : 1;
    const uint32_t idesc = umma_idesc_bf16(64, 8);
    const uint32_t ring_lo = (smem_u32(smem) >> 4) | 0x10000u, xb_lo = (smem_u32(xb) >> 4) | 0x10000u;
    const uint32_t hi = (1024u >> 4) | (1u << 14) | (2u << 29);
    const long long t0 = clock64();
    long long stall = 0;
#pragma unroll 1
    for (int i = 0; i < n_slots; ++i) {
      const int s = i % nstage;
      if (n_mma_warps == 2 && (i & 1) != wsel) continue;
      if (mode == 1) {
        const long long w0 = clock64();
        mbar_wait_parked(&full[s], (i / nstage) & 1);
        stall += clock64() - w0;
        tcgen05_fence_after();
      }
      const uint32_t a_lo = ring_lo + s * (ST_BYTES >> 4), b_lo = xb_lo + (i % 12) * 64;
      uint32_t pred;
      asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(pred));
      if (pred) {
#pragma unroll
        for (int m = 0; m < 12; ++m)
          if (m < mma_per_slot)
            asm volatile(
                "{\n.reg .pred p;\n.reg .b64 da, db;\nmov.b64 da, {%1, %4};\nmov.b64 db, {%2, %4};\nsetp.ne.b32 p, 1, 0;\n"
                "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, p;\n}\n" ::"r"(tmem + (m % 3) * 8 + wsel * 32),
                "r"(a_lo + (((m % 3) * 8192 + (m / 3) * 32) >> 4)), "r"(b_lo + (m / 3) * 2), "r"(idesc), "r"(hi)
                : "memory");
        if (mode == 1 || (i % commit_every) == commit_every - 1 || i >= n_slots - 2) umma_commit(&empty[s]);
      }
      __syncwarp();
    }
    // drain: the last commit (mode 0 with sparse commits: parity bookkeeping is skipped, the last two slots always commit)
    if (mode == 1) mbar_wait_parked(&empty[(n_slots - 1) % nstage], ((n_slots - 1) / nstage) & 1);
    else asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    if (tid == 32) {
      out[blockIdx.x * 2] = clock64() - t0;
      out[blockIdx.x * 2 + 1] = stall;
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem, 64);
}
// clean issue loop: compile-time shapes, no divisions; COMMIT_EVERY = 0: never commit inside the loop
template <int MPS, int COMMIT_EVERY, int NW>
__global__ void __launch_bounds__(128, 1) probe_issue(int n_slots, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* xb = smem + 6 * ST_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(xb + 12288);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bars + 16);
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < (6 * ST_BYTES + 12288) / 16; i += 128) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  fence_proxy_async_smem();
  if (tid == 0) {
    for (int s = 0; s < 16; ++s) mbar_init(&bars[s], 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(slot, 64);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = *slot;
  if (warp == 1 || (NW == 2 && warp == 3)) {
    const int wsel = warp == 1 ? 0 : 1;
    const uint32_t idesc = umma_idesc_bf16(64, 8);
    const uint32_t ring_lo = (smem_u32(smem) >> 4) | 0x10000u, xb_lo = (smem_u32(xb) >> 4) | 0x10000u;
    const uint32_t hi = (1024u >> 4) | (1u << 14) | (2u << 29);
    const long long t0 = clock64();
    int s = wsel, since = 0;
#pragma unroll 1
    for (int i = wsel; i < n_slots; i += NW) {
      const uint32_t a_lo = ring_lo + s * (ST_BYTES >> 4);
      uint32_t pred;
      asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(pred));
      if (pred) {
#pragma unroll
        for (int m = 0; m < MPS; ++m)
          asm volatile(
              "{\n.reg .pred p;\n.reg .b64 da, db;\nmov.b64 da, {%1, %4};\nmov.b64 db, {%2, %4};\nsetp.ne.b32 p, 1, 0;\n"
              "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, p;\n}\n" ::"r"(tmem + (m % 3) * 8 + wsel * 32),
              "r"(a_lo + (((m % 3) * 8192 + (m / 3) * 32) >> 4)), "r"(xb_lo + (m / 3) * 2), "r"(idesc), "r"(hi)
              : "memory");
        if (COMMIT_EVERY > 0 && ++since == COMMIT_EVERY) {
          since = 0;
          umma_commit(&bars[s + 8 * wsel]);
        }
      }
      __syncwarp();
      s += NW;
      if (s >= 6) s -= 6;
    }
    const long long t1 = clock64();
    if ((tid & 31) == 0) {
      umma_commit(&bars[6 + 8 * wsel]);
      mbar_wait_parked(&bars[6 + 8 * wsel], 0);
      out[wsel * 2] = t1 - t0;
      out[wsel * 2 + 1] = clock64() - t0;
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem, 64);
}
template <int MPS, int CE, int NW>
static void run_issue_one(long long* d) {
  const int smem_bytes = 6 * ST_BYTES + 12288 + 256, n = 1200;
  CK(cudaFuncSetAttribute(probe_issue<MPS, CE, NW>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
  probe_issue<MPS, CE, NW><<<1, 128, smem_bytes>>>(n, d);
  CK(cudaDeviceSynchronize());
  long long h[4];
  CK(cudaMemcpy(h, d, 32, cudaMemcpyDeviceToHost));
  printf("[issue] %2d MMAs per slot, commit every %d slot(s), %d issuing warp(s): issue loop %.0f cycles per slot, all complete %.0f\n", MPS, CE,
         NW, double(h[0]) / n, double(h[1]) / n);
}
static void run_issue() {
  long long* d;
  CK(cudaMalloc(&d, 64));
  run_issue_one<0, 1, 1>(d);
  run_issue_one<0, 0, 1>(d);
  run_issue_one<6, 1, 1>(d);
  run_issue_one<12, 1, 1>(d);
  run_issue_one<12, 3, 1>(d);
  run_issue_one<12, 0, 1>(d);
  run_issue_one<16, 1, 1>(d);
  run_issue_one<16, 3, 1>(d);
  run_issue_one<12, 1, 2>(d);
  run_issue_one<12, 3, 2>(d);
  run_issue_one<12, 0, 2>(d);
  cudaFree(d);
}

static void run_mma_stream() {
  const size_t rows = 1200000;
  bf16* buf;
  CK(cudaMalloc(&buf, rows * 128));
  CK(cudaMemset(buf, 0, rows * 128));
  CUtensorMap tm;
  if (make_tmap_2d_bf16(&tm, buf, 64, rows, 64, 64, ST_ROWS) != LRCE_OK) exit(1);
  long long* d;
  CK(cudaMalloc(&d, 148 * 16));
  const int smem_bytes = 6 * ST_BYTES + 12288 + 256;
  CK(cudaFuncSetAttribute(probe_mma_stream, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
  const int n_slots = 1200;
  for (int ce : {1, 2, 3, 6})
    for (int nw : {1, 2}) {
      probe_mma_stream<<<1, 128, smem_bytes>>>(tm, n_slots, 6, 0, 12, d, ce, nw);
      CK(cudaDeviceSynchronize());
      long long h[2];
      CK(cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost));
      printf("[mma+tma] no TMA, 12 MMAs per slot, commit every %d slot(s), %d issuing warp(s): %.0f cycles per slot (issue side)\n", ce, nw,
             double(h[0]) / n_slots);
    }
  for (int grid : {1, 112})
    for (int mode : {0, 1})
      for (int mps : {12, 6, 0}) {
        probe_mma_stream<<<grid, 128, smem_bytes>>>(tm, n_slots, 6, mode, mps, d, 1, 1);
        CK(cudaDeviceSynchronize());
        std::vector<long long> h(grid * 2);
        CK(cudaMemcpy(h.data(), d, grid * 16, cudaMemcpyDeviceToHost));
        long long mx = 0, st = 0;
        for (int i = 0; i < grid; ++i)
          if (h[2 * i] > mx) { mx = h[2 * i]; st = h[2 * i + 1]; }
        printf("[mma+tma] grid %3d, %s, %2d MMAs per 24 KB slot: %.0f cycles per slot (%.0f of them waiting for the slot)\n", grid,
               mode ? "TMA stream on " : "TMA stream off", mps, double(mx) / n_slots, double(st) / n_slots);
      }
  cudaFree(buf);
  cudaFree(d);
}

int main(int argc, char** argv) {
  const char* what = argc > 1 ? argv[1] : "all";
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  printf("device: %s, %d SMs, smem/block optin %zu\n", prop.name, prop.multiProcessorCount, prop.sharedMemPerBlockOptin);
  if (!strcmp(what, "m64") || !strcmp(what, "all")) run_m64();
  if (!strcmp(what, "xchg") || !strcmp(what, "all")) {
    run_xchg<16>(220 * 1024);
    run_xchg<16>(100 * 1024);
    run_xchg<8>(220 * 1024);
  }
  if (!strcmp(what, "mma") || !strcmp(what, "all")) run_mma_rate();
  if (!strcmp(what, "mmastream") || !strcmp(what, "all")) run_mma_stream();
  if (!strcmp(what, "issue") || !strcmp(what, "all")) run_issue();
  if (!strcmp(what, "stream") || !strcmp(what, "all")) {
    run_stream<16>();
    run_stream<8>();
  }
  return 0;
}
