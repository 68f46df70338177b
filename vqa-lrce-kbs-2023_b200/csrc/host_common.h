// host_common.h — host-side helpers private to liblrce_b200.so: error convention, arch gate, TMA descriptor encode.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/lrce_b200.h"

namespace lrce {

// thread-local last-error message (retrieved through lrce_last_error())
void set_error(const char* fmt, ...);
// returns LRCE_OK only on a compute-capability 10.x device; there is no fallback path.
int require_sm100();
// cudaGetLastError() -> LRCE_ECUDA + message
int check_launch(const char* what);

// 2-D bf16 tensor map: global tensor [outer][inner] with row pitch ld_elems, box = box_outer x box_inner, smem side
// swizzled with a 128-byte (UMMA SW128 K-major operand tiles; box_inner * 2 B == 128) or 64-byte (box_inner * 2 B == 64)
// pattern.
int make_tmap_2d_bf16(CUtensorMap* tm, const void* base, uint64_t inner, uint64_t outer, uint64_t ld_elems,
                      uint32_t box_inner, uint32_t box_outer, int swizzle_bytes = 128);

// N-D (rank <= 5) bf16 tensor map: dims[0] is the contiguous dimension; strides_bytes[i] is the byte stride of dims[i+1];
// swizzle_bytes in {0, 64, 128}; l2_promo_bytes in {0, 64, 128, 256}.
int make_tmap_nd_bf16(CUtensorMap* tm, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                      const uint32_t* box, int swizzle_bytes, int l2_promo_bytes);

int sm_count();
// id of the calling thread's current CUDA device (-1 on error)
int current_device();
// Per-(call site, device) one-time setup (cudaFuncSetAttribute and friends are per device): `mask` is the call site's
// static thread_local bitmap. needs_device_setup() is true until mark_device_setup() was called for the current device.
bool needs_device_setup(const uint64_t* mask);
void mark_device_setup(uint64_t* mask);

// Programmatic dependent launch: a kernel launched with this attribute may start (block scheduling, barrier / TMEM set-up, tensor
// map prefetch) while the tail of the previous kernel of the stream drains, once every block of that kernel has executed
// griddepcontrol.launch_dependents; the kernel itself must execute griddepcontrol.wait (griddep_wait() in lrce_common.cuh) before
// its first access to global memory. LRCE_B200_PDL=0 launches without the attribute (A/B runs).
bool pdl_enabled();
inline int pdl_attr(cudaLaunchAttribute* a) {
  if (!pdl_enabled()) return 0;
  a->id = cudaLaunchAttributeProgrammaticStreamSerialization;
  a->val.programmaticStreamSerializationAllowed = 1;
  return 1;
}
template <typename... KArgs, typename... Args>
cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute at[1];
  cfg.numAttrs = pdl_attr(at);
  cfg.attrs = at;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

}  // namespace lrce

#define LRCE_REQUIRE(cond, ...)          \
  do {                                   \
    if (!(cond)) {                       \
      ::lrce::set_error(__VA_ARGS__);    \
      return LRCE_EINVAL;                \
    }                                    \
  } while (0)
