// encoder_common.cuh — row helpers shared by the recurrent cross-modal encoder's kernels (encoder.cu, encoder_walk.cu).
#pragma once
#include "lrce_common.cuh"

namespace lrce {

constexpr int ENC_D = 768;

// ---------------------------------------------------------------------------------------------------------------
// row helpers: one warp owns a 768-wide row, lane holds 24 values as 3 chunks of 8 at columns (i*32 + lane)*8
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void row768_ln_store(float (&v)[24], const float* __restrict__ gamma,
                                                const float* __restrict__ beta, float eps, int lane,
                                                bf16* out_bf16, float* out_f32) {
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 24; ++i) s += v[i];
  const float mean = warp_sum(s) * (1.0f / ENC_D);
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < 24; ++i) { const float d = v[i] - mean; ss += d * d; }
  const float rstd = rsqrtf(warp_sum(ss) * (1.0f / ENC_D) + eps);
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const int col = (c * 32 + lane) * 8;
    float o[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = (v[c * 8 + j] - mean) * rstd * __ldg(gamma + col + j) + __ldg(beta + col + j);
    if (out_bf16) {
      uint4 u;
      u.x = pack_bf16x2(o[0], o[1]); u.y = pack_bf16x2(o[2], o[3]);
      u.z = pack_bf16x2(o[4], o[5]); u.w = pack_bf16x2(o[6], o[7]);
      *reinterpret_cast<uint4*>(out_bf16 + col) = u;
    }
    if (out_f32) {
      *reinterpret_cast<float4*>(out_f32 + col) = make_float4(o[0], o[1], o[2], o[3]);
      *reinterpret_cast<float4*>(out_f32 + col + 4) = make_float4(o[4], o[5], o[6], o[7]);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) v[c * 8 + j] = o[j];
  }
}

__device__ __forceinline__ void row768_add_f32(float (&v)[24], const float* __restrict__ src, int lane) {
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const int col = (c * 32 + lane) * 8;
    const float4 a = __ldg(reinterpret_cast<const float4*>(src + col));
    const float4 b = __ldg(reinterpret_cast<const float4*>(src + col + 4));
    v[c * 8 + 0] += a.x; v[c * 8 + 1] += a.y; v[c * 8 + 2] += a.z; v[c * 8 + 3] += a.w;
    v[c * 8 + 4] += b.x; v[c * 8 + 5] += b.y; v[c * 8 + 6] += b.z; v[c * 8 + 7] += b.w;
  }
}
__device__ __forceinline__ void row768_add_bf16(float (&v)[24], const bf16* __restrict__ src, int lane) {
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const int col = (c * 32 + lane) * 8;
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(src + col));
    float2 f;
    f = unpack_bf16x2(u.x); v[c * 8 + 0] += f.x; v[c * 8 + 1] += f.y;
    f = unpack_bf16x2(u.y); v[c * 8 + 2] += f.x; v[c * 8 + 3] += f.y;
    f = unpack_bf16x2(u.z); v[c * 8 + 4] += f.x; v[c * 8 + 5] += f.y;
    f = unpack_bf16x2(u.w); v[c * 8 + 6] += f.x; v[c * 8 + 7] += f.y;
  }
}

__device__ __forceinline__ void mma16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                         uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
      "{%0, %1, %2, %3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

}  // namespace lrce
