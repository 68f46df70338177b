// encoder.cu — the recurrent cross-modal encoder's small kernels (lrce/models/fusionv3.py, embedding.py).
//
// The heavy part of the encoder — the K/V in-projection of every memory token for all 12 layers — is one tcgen05 GEMM
// (gemm_tc.cu); the summarisation token's walk through the 12 layers x S segments is the persistent kernel of
// encoder_walk.cu. Kernels here:
//   video_posembed_ln / text_posembed_ln : embedding.py:47-63 / :17-23 fused (CLS row, 3 adds, LayerNorm eps 1e-12)
#include "host_common.h"
#include "encoder_common.cuh"

namespace lrce {

// out[b,s,t,p] = LN(emb_pos[p] + emb_len[t] + emb_clip[s] + (p == 0 ? emb_cls : proj[b,s,t,p-1]))
__global__ void __launch_bounds__(256) video_posembed_ln_kernel(const bf16* __restrict__ proj, const float* __restrict__ emb_cls,
                                                                const float* __restrict__ emb_pos, const float* __restrict__ emb_len,
                                                                const float* __restrict__ emb_clip, const float* __restrict__ gamma,
                                                                const float* __restrict__ beta, float eps, bf16* __restrict__ out,
                                                                int B, int S, int T, int P) {
  griddep_wait();  // programmatic dependent launch (host_common.h): no global access before this point
  griddep_launch();
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
  griddep_wait();
  griddep_launch();
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
  cudaError_t e = launch_pdl(video_posembed_ln_kernel, dim3(static_cast<unsigned>((rows + 7) / 8)), dim3(256), 0,
                             reinterpret_cast<cudaStream_t>(stream), reinterpret_cast<const bf16*>(proj), emb_cls, emb_pos, emb_len,
                             emb_clip, gamma, beta, eps, reinterpret_cast<bf16*>(out), B, S, T, P);
  if (e != cudaSuccess) {
    set_error("cudaLaunchKernelEx(video_posembed_ln_kernel): %s", cudaGetErrorString(e));
    return LRCE_ECUDA;
  }
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
  cudaError_t e;
  if (text_fp32)
    e = launch_pdl(text_posembed_ln_kernel<float>, dim3(blocks), dim3(256), 0, s, reinterpret_cast<const float*>(text), emb_cls, emb_pos,
                   gamma, beta, eps, reinterpret_cast<bf16*>(out), Bt, L);
  else
    e = launch_pdl(text_posembed_ln_kernel<bf16>, dim3(blocks), dim3(256), 0, s, reinterpret_cast<const bf16*>(text), emb_cls, emb_pos,
                   gamma, beta, eps, reinterpret_cast<bf16*>(out), Bt, L);
  if (e != cudaSuccess) {
    set_error("cudaLaunchKernelEx(text_posembed_ln_kernel): %s", cudaGetErrorString(e));
    return LRCE_ECUDA;
  }
  return check_launch("text_posembed_ln_kernel");
}

