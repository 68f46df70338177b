// mufu_probe.cu — MUFU throughput per SM: ex2.approx.ftz.f32 vs ex2.approx.f16x2 / bf16x2, tanh.approx.f32 vs f16x2
// (does the packed form give two results per MUFU slot?). Build: tools/probes/build.sh mufu_probe
#include <cuda_fp16.h>
#include <stdio.h>
#include <stdlib.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e_)); exit(1); } } while (0)

template <int MODE>
__global__ void __launch_bounds__(1024) k(float* out, long long* cyc, int iters) {
  float a[8];
  unsigned h[8];
  for (int i = 0; i < 8; ++i) { a[i] = -0.001f * (threadIdx.x + i); h[i] = 0xB800B400u + threadIdx.x + i; }
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
      if (MODE == 1) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(h[i]));
      if (MODE == 2) asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(h[i]));
      if (MODE == 3) asm volatile("tanh.approx.f32 %0, %0;" : "+f"(a[i]));
      if (MODE == 4) asm volatile("tanh.approx.f16x2 %0, %0;" : "+r"(h[i]));
    }
  }
  const long long t1 = clock64();
  __syncthreads();
  float s = 0;
  for (int i = 0; i < 8; ++i) s += a[i] + __uint_as_float(h[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int MODE>
static void run(const char* name, int vals_per_op) {
  float* out; long long* cyc;
  CK(cudaMalloc(&out, 148 * 1024 * 4)); CK(cudaMalloc(&cyc, 148 * 8));
  const int iters = 2000;
  for (int threads : {128, 512, 1024}) {
    k<MODE><<<148, threads>>>(out, cyc, iters);
    CK(cudaDeviceSynchronize());
    long long h[148]; CK(cudaMemcpy(h, cyc, 148 * 8, cudaMemcpyDeviceToHost));
    long long mx = 0; for (int i = 0; i < 148; ++i) mx = h[i] > mx ? h[i] : mx;
    const double ops = double(threads) * 8 * iters;
    printf("%-26s %4d threads/SM: %.2f MUFU ops/clk/SM = %.2f results/clk/SM\n", name, threads, ops / mx, ops * vals_per_op / mx);
  }
  cudaFree(out); cudaFree(cyc);
}
int main() {
  run<0>("ex2.approx.ftz.f32", 1);
  run<1>("ex2.approx.f16x2", 2);
  run<2>("ex2.approx.ftz.bf16x2", 2);
  run<3>("tanh.approx.f32", 1);
  run<4>("tanh.approx.f16x2", 2);
  return 0;
}
