// rowops.cu — the HBM-bound row kernels of the Video Swin path: LayerNorm (plain and fused with the PatchMerging
// 2x2 gather), the PatchEmbed3D patch gather (normalise + zero 6th frame + im2col), and the standalone
// shift/partition/reverse remap used by the bit-exact tests. Each row is handled by a group of lanes that keeps the
// whole row in registers (one read, one write: algorithmic bytes = 2 x rows x C x 2 B), 16-byte accesses throughout.
#include "host_common.h"
#include "lrce_common.cuh"
#include "remap.cuh"

namespace lrce {

// ------------------------------------------------------------------------------------------------------------
// LayerNorm over the last dim of a bf16 [rows, C] matrix. MERGE: the input row is the concatenation of 4 source rows
// of width C/4 gathered by PatchMerging's 2x2 rule. LANES lanes cooperate on a row (C = LANES * 8 * VEC).
// ------------------------------------------------------------------------------------------------------------
template <int C, bool MERGE, typename OutT>
__global__ void __launch_bounds__(256) layernorm_rows_kernel(const bf16* __restrict__ x, OutT* __restrict__ y,
                                                             const float* __restrict__ gamma,
                                                             const float* __restrict__ beta, float eps, long long rows,
                                                             int D, int H, int W) {
  griddep_wait();  // programmatic dependent launch (host_common.h): no global access before this point
  griddep_launch();
  constexpr int LANES = (C / 8 < 32) ? C / 8 : 32;
  constexpr int VEC = C / (8 * LANES);
  constexpr int ROWS_PER_WARP = 32 / LANES;
  const int lane = threadIdx.x & 31;
  const int sub = lane / LANES, l = lane % LANES;
  const long long warp_global = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const long long row = warp_global * ROWS_PER_WARP + sub;
  const bool active = row < rows;
  float v[VEC][8];
  float s = 0.f;
  if (active) {
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      const int col = (i * LANES + l) * 8;
      const bf16* src;
      if (MERGE) {
        constexpr int CS = C / 4;  // source channel count
        const int part = col / CS;
        const long long per_seg_out = static_cast<long long>(D) * (H / 2) * (W / 2);
        const long long seg = row / per_seg_out;
        const int out_tok = static_cast<int>(row - seg * per_seg_out);
        const long long tok = seg * (static_cast<long long>(D) * H * W) + merge_source_token(D, H, W, out_tok, part);
        src = x + tok * CS + (col - part * CS);
      } else {
        src = x + row * C + col;
      }
      const uint4 u = *reinterpret_cast<const uint4*>(src);
      float2 f;
      f = unpack_bf16x2(u.x); v[i][0] = f.x; v[i][1] = f.y;
      f = unpack_bf16x2(u.y); v[i][2] = f.x; v[i][3] = f.y;
      f = unpack_bf16x2(u.z); v[i][4] = f.x; v[i][5] = f.y;
      f = unpack_bf16x2(u.w); v[i][6] = f.x; v[i][7] = f.y;
#pragma unroll
      for (int j = 0; j < 8; ++j) s += v[i][j];
    }
  }
#pragma unroll
  for (int o = LANES / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mean = s * (1.0f / C);
  float ss = 0.f;
  if (active) {
#pragma unroll
    for (int i = 0; i < VEC; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float d = v[i][j] - mean;
        ss += d * d;
      }
  }
#pragma unroll
  for (int o = LANES / 2; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  const float rstd = rsqrtf(ss * (1.0f / C) + eps);
  if (!active) return;
#pragma unroll
  for (int i = 0; i < VEC; ++i) {
    const int col = (i * LANES + l) * 8;
    const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + col));
    const float4 g1 = __ldg(reinterpret_cast<const float4*>(gamma + col + 4));
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta + col));
    const float4 b1 = __ldg(reinterpret_cast<const float4*>(beta + col + 4));
    float o[8];
    o[0] = (v[i][0] - mean) * rstd * g0.x + b0.x;
    o[1] = (v[i][1] - mean) * rstd * g0.y + b0.y;
    o[2] = (v[i][2] - mean) * rstd * g0.z + b0.z;
    o[3] = (v[i][3] - mean) * rstd * g0.w + b0.w;
    o[4] = (v[i][4] - mean) * rstd * g1.x + b1.x;
    o[5] = (v[i][5] - mean) * rstd * g1.y + b1.y;
    o[6] = (v[i][6] - mean) * rstd * g1.z + b1.z;
    o[7] = (v[i][7] - mean) * rstd * g1.w + b1.w;
    if (sizeof(OutT) == 2) {
      uint4 u;
      u.x = pack_bf16x2(o[0], o[1]); u.y = pack_bf16x2(o[2], o[3]);
      u.z = pack_bf16x2(o[4], o[5]); u.w = pack_bf16x2(o[6], o[7]);
      *reinterpret_cast<uint4*>(reinterpret_cast<bf16*>(y) + row * C + col) = u;
    } else {
      float4* dst = reinterpret_cast<float4*>(reinterpret_cast<float*>(y) + row * C + col);
      dst[0] = make_float4(o[0], o[1], o[2], o[3]);
      dst[1] = make_float4(o[4], o[5], o[6], o[7]);
    }
  }
}

template <int C, bool MERGE, typename OutT>
static int launch_ln(const void* x, void* y, const float* g, const float* b, float eps, long long rows, int D, int H,
                     int W, cudaStream_t s) {
  constexpr int LANES = (C / 8 < 32) ? C / 8 : 32;
  constexpr int ROWS_PER_WARP = 32 / LANES;
  const long long warps = (rows + ROWS_PER_WARP - 1) / ROWS_PER_WARP;
  const long long blocks = (warps + 7) / 8;
  cudaError_t e = launch_pdl(layernorm_rows_kernel<C, MERGE, OutT>, dim3(static_cast<unsigned>(blocks)), dim3(256), 0, s,
                             reinterpret_cast<const bf16*>(x), reinterpret_cast<OutT*>(y), g, b, eps, rows, D, H, W);
  if (e != cudaSuccess) {
    set_error("cudaLaunchKernelEx(layernorm_rows_kernel): %s", cudaGetErrorString(e));
    return LRCE_ECUDA;
  }
  return check_launch("layernorm_rows_kernel");
}

template <bool MERGE, typename OutT>
static int dispatch_ln(int C, const void* x, void* y, const float* g, const float* b, float eps, long long rows, int D,
                       int H, int W, cudaStream_t s) {
  switch (C) {
    case 128: return launch_ln<128, MERGE, OutT>(x, y, g, b, eps, rows, D, H, W, s);
    case 256: return launch_ln<256, MERGE, OutT>(x, y, g, b, eps, rows, D, H, W, s);
    case 512: return launch_ln<512, MERGE, OutT>(x, y, g, b, eps, rows, D, H, W, s);
    case 768: return launch_ln<768, MERGE, OutT>(x, y, g, b, eps, rows, D, H, W, s);
    case 1024: return launch_ln<1024, MERGE, OutT>(x, y, g, b, eps, rows, D, H, W, s);
    case 2048: return launch_ln<2048, MERGE, OutT>(x, y, g, b, eps, rows, D, H, W, s);
    default: break;
  }
  set_error("LayerNorm width %d is not one of 128/256/512/768/1024/2048", C);
  return LRCE_EINVAL;
}

// ------------------------------------------------------------------------------------------------------------
// PatchEmbed3D gather: clips fp32 (n_seg, T, 3, Hin, Win) in [0,1] -> A bf16 [n_seg * D * Hp * Wp, 96] with
// K = (c, kd, kh, kw) ordered like the Conv3d weight; ImageNet mean/std folded into the load; frames >= T are the
// zero padding the reference appends AFTER normalisation (video_swin_ori.py:472-473). One thread = 16 K-values.
// ------------------------------------------------------------------------------------------------------------
// Pixel type: float in [0,1] (what torchvision's ToTensor hands the reference, e2e_dataset.py) or the uint8 frame itself
// (x / 255 done here in fp32: same value, a quarter of the host -> device bytes).
__device__ __forceinline__ float4 load_px4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float4 load_px4(const uint8_t* p) {
  const uchar4 u = __ldg(reinterpret_cast<const uchar4*>(p));
  return make_float4(u.x / 255.0f, u.y / 255.0f, u.z / 255.0f, u.w / 255.0f);  // torch: byte -> float, then .div(255)
}
template <typename PixT>
__global__ void __launch_bounds__(256) patch_gather_kernel(const PixT* __restrict__ clips, bf16* __restrict__ A,
                                                           int n_seg, int T, int Hin, int Win) {
  griddep_wait();  // the output buffer may still be read by the previous kernel of the stream
  griddep_launch();
  const int D = (T + 1) / 2, Hp = Hin / 4, Wp = Win / 4;
  const long long total = static_cast<long long>(n_seg) * D * Hp * Wp * 6;
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int ckd = static_cast<int>(idx % 6);
  const long long row = idx / 6;
  const int c = ckd >> 1, kd = ckd & 1;
  const int wp = static_cast<int>(row % Wp);
  const int hp = static_cast<int>((row / Wp) % Hp);
  const int d = static_cast<int>((row / (static_cast<long long>(Wp) * Hp)) % D);
  const long long n = row / (static_cast<long long>(Wp) * Hp * D);
  const int t = 2 * d + kd;
  uint4 o[2] = {make_uint4(0, 0, 0, 0), make_uint4(0, 0, 0, 0)};
  if (t < T) {
    const float mean = (c == 0) ? 0.485f : (c == 1) ? 0.456f : 0.406f;
    const float sdev = (c == 0) ? 0.229f : (c == 1) ? 0.224f : 0.225f;
    const PixT* src = clips + (((n * T + t) * 3 + c) * Hin + 4 * hp) * static_cast<long long>(Win) + 4 * wp;
    uint32_t p[8];
#pragma unroll
    for (int kh = 0; kh < 4; ++kh) {
      const float4 f = load_px4(src + static_cast<long long>(kh) * Win);
      // (x - mean) / std in fp32, the order torchvision Normalize uses (video.py:35), then one rounding to bf16
      p[2 * kh + 0] = pack_bf16x2((f.x - mean) / sdev, (f.y - mean) / sdev);
      p[2 * kh + 1] = pack_bf16x2((f.z - mean) / sdev, (f.w - mean) / sdev);
    }
    o[0] = make_uint4(p[0], p[1], p[2], p[3]);
    o[1] = make_uint4(p[4], p[5], p[6], p[7]);
  }
  uint4* dst = reinterpret_cast<uint4*>(A + row * 96 + ckd * 16);
  dst[0] = o[0];
  dst[1] = o[1];
}

// ------------------------------------------------------------------------------------------------------------
// Standalone cyclic-shift + window partition (gather) and its inverse (scatter) on a bf16 [n_seg, D*H*W, C] tensor.
// The production path fuses this index map into the attention kernel's loads/stores; this kernel exists so the map
// can be checked bit-exactly against torch.roll + window_partition / window_reverse, and measured against HBM peak.
// ------------------------------------------------------------------------------------------------------------
template <bool SCATTER>
__global__ void __launch_bounds__(256) window_remap_kernel(const bf16* __restrict__ in, bf16* __restrict__ out,
                                                           StageGeom g, int n_seg, int C) {
  const int chunks = C / 8;
  const int N = window_tokens(g), nwin = windows_per_segment(g);
  const long long total = static_cast<long long>(n_seg) * nwin * N * chunks;
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int ch = static_cast<int>(idx % chunks);
  const long long wrow = idx / chunks;  // row in window order
  const int tok = static_cast<int>(wrow % N);
  const int win = static_cast<int>((wrow / N) % nwin);
  const long long seg = wrow / (static_cast<long long>(N) * nwin);
  const int nw_ = g.W / g.ww, nh_ = g.H / g.wh;
  const int src_tok = is_window_377(g) ? window_source_token_377(g, (win / nw_) % nh_, win % nw_, tok) : window_source_token(g, win, tok);
  const long long nrow = seg * (static_cast<long long>(g.D) * g.H * g.W) + src_tok;
  const uint4* src = reinterpret_cast<const uint4*>(in + (SCATTER ? wrow : nrow) * C) + ch;
  uint4* dst = reinterpret_cast<uint4*>(out + (SCATTER ? nrow : wrow) * C) + ch;
  *dst = __ldg(src);
}

__global__ void remap_index_kernel(int* __restrict__ gather, int* __restrict__ region, int* __restrict__ relpos,
                                   StageGeom g) {
  const int N = window_tokens(g), nwin = windows_per_segment(g);
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= nwin * N) return;
  const int win = idx / N, tok = idx % N;
  if (gather) {
    const int nw_ = g.W / g.ww, nh_ = g.H / g.wh;
    gather[idx] = is_window_377(g) ? window_source_token_377(g, (win / nw_) % nh_, win % nw_, tok) : window_source_token(g, win, tok);
  }
  if (region) region[idx] = shift_region_id(g, win, tok);
  if (relpos && win == 0) relpos[tok] = rel_pos_offset(g, tok);
}

}  // namespace lrce

using namespace lrce;

static int make_geom(StageGeom* g, int D, int H, int W, int wd, int wh, int ww, int sd, int sh, int sw) {
  LRCE_REQUIRE(D > 0 && H > 0 && W > 0 && wd > 0 && wh > 0 && ww > 0, "bad stage geometry");
  LRCE_REQUIRE(D % wd == 0 && H % wh == 0 && W % ww == 0, "token grid (%d,%d,%d) must tile by window (%d,%d,%d)", D, H,
               W, wd, wh, ww);
  LRCE_REQUIRE(sd >= 0 && sd < wd && sh >= 0 && sh < wh && sw >= 0 && sw < ww, "shift must be in [0, window)");
  g->D = D; g->H = H; g->W = W; g->wd = wd; g->wh = wh; g->ww = ww; g->sd = sd; g->sh = sh; g->sw = sw;
  return LRCE_OK;
}

extern "C" int lrce_layernorm_bf16(const void* x, void* y, const float* gamma, const float* beta, float eps,
                                   long long rows, int C, int out_fp32, void* stream) {
  int rc = require_sm100();
  if (rc != LRCE_OK) return rc;
  LRCE_REQUIRE(x && y && gamma && beta && rows > 0, "lrce_layernorm_bf16: bad arguments");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  return out_fp32 ? dispatch_ln<false, float>(C, x, y, gamma, beta, eps, rows, 0, 0, 0, s)
                  : dispatch_ln<false, bf16>(C, x, y, gamma, beta, eps, rows, 0, 0, 0, s);
}

extern "C" int lrce_patch_merge_ln_bf16(const void* x, void* y, const float* gamma, const float* beta, float eps,
                                        int n_seg, int D, int H, int W, int C, void* stream) {
  int rc = require_sm100();
  if (rc != LRCE_OK) return rc;
  LRCE_REQUIRE(x && y && gamma && beta && n_seg > 0, "lrce_patch_merge_ln_bf16: bad arguments");
  LRCE_REQUIRE(H % 2 == 0 && W % 2 == 0, "lrce_patch_merge_ln_bf16: H and W must be even (no padding path), got %dx%d", H, W);
  const long long rows = static_cast<long long>(n_seg) * D * (H / 2) * (W / 2);
  return dispatch_ln<true, bf16>(4 * C, x, y, gamma, beta, eps, rows, D, H, W, reinterpret_cast<cudaStream_t>(stream));
}

template <typename PixT>
static int patch_gather(const PixT* clips, void* A, int n_seg, int T, int Hin, int Win, void* stream, const char* what) {
  int rc = require_sm100();
  if (rc != LRCE_OK) return rc;
  LRCE_REQUIRE(clips && A && n_seg > 0 && T > 0, "%s: bad arguments", what);
  LRCE_REQUIRE(Hin % 4 == 0 && Win % 4 == 0, "%s: frame size %dx%d must be a multiple of the 4x4 patch", what, Hin, Win);
  LRCE_REQUIRE((reinterpret_cast<uintptr_t>(clips) & (4 * sizeof(PixT) - 1)) == 0, "%s: clips must be aligned to 4 pixels", what);
  const long long total = static_cast<long long>(n_seg) * ((T + 1) / 2) * (Hin / 4) * (Win / 4) * 6;
  cudaError_t e = launch_pdl(patch_gather_kernel<PixT>, dim3(static_cast<unsigned>((total + 255) / 256)), dim3(256), 0,
                             reinterpret_cast<cudaStream_t>(stream), clips, reinterpret_cast<bf16*>(A), n_seg, T, Hin, Win);
  if (e != cudaSuccess) {
    set_error("cudaLaunchKernelEx(patch_gather_kernel): %s", cudaGetErrorString(e));
    return LRCE_ECUDA;
  }
  return check_launch("patch_gather_kernel");
}

extern "C" int lrce_patch_gather_f32(const float* clips, void* A, int n_seg, int T, int Hin, int Win, void* stream) {
  return patch_gather<float>(clips, A, n_seg, T, Hin, Win, stream, "lrce_patch_gather_f32");
}

extern "C" int lrce_patch_gather_u8(const unsigned char* clips, void* A, int n_seg, int T, int Hin, int Win, void* stream) {
  return patch_gather<uint8_t>(clips, A, n_seg, T, Hin, Win, stream, "lrce_patch_gather_u8");
}

extern "C" int lrce_window_remap_bf16(const void* in, void* out, int n_seg, int D, int H, int W, int C, int wd, int wh,
                                      int ww, int sd, int sh, int sw, int inverse, void* stream) {
  int rc = require_sm100();
  if (rc != LRCE_OK) return rc;
  StageGeom g;
  rc = make_geom(&g, D, H, W, wd, wh, ww, sd, sh, sw);
  if (rc != LRCE_OK) return rc;
  LRCE_REQUIRE(in && out && in != out && n_seg > 0 && C % 8 == 0, "lrce_window_remap_bf16: bad arguments");
  const long long total = static_cast<long long>(n_seg) * D * H * W * (C / 8);
  const unsigned blocks = static_cast<unsigned>((total + 255) / 256);
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (inverse)
    window_remap_kernel<true><<<blocks, 256, 0, s>>>(reinterpret_cast<const bf16*>(in), reinterpret_cast<bf16*>(out), g, n_seg, C);
  else
    window_remap_kernel<false><<<blocks, 256, 0, s>>>(reinterpret_cast<const bf16*>(in), reinterpret_cast<bf16*>(out), g, n_seg, C);
  return check_launch("window_remap_kernel");
}

extern "C" int lrce_remap_index(int* gather, int* region, int* relpos, int D, int H, int W, int wd, int wh, int ww,
                                int sd, int sh, int sw, void* stream) {
  int rc = require_sm100();
  if (rc != LRCE_OK) return rc;
  StageGeom g;
  rc = make_geom(&g, D, H, W, wd, wh, ww, sd, sh, sw);
  if (rc != LRCE_OK) return rc;
  const int total = D * H * W;
  remap_index_kernel<<<(total + 255) / 256, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(gather, region, relpos, g);
  return check_launch("remap_index_kernel");
}
