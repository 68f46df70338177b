// host_common.cu — error convention, arch gate and TMA descriptor encoding for liblrce_b200.so.
#include "host_common.h"

#include <stdlib.h>

#include <stdarg.h>
#include <stdio.h>
#include <string.h>

namespace lrce {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

const char* last_error() { return g_err; }

int require_sm100() {
  static thread_local int cached_dev = -1;
  static thread_local int cached_rc = LRCE_EARCH;
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) {
    set_error("cudaGetDevice failed: %s", cudaGetErrorString(e));
    return LRCE_ECUDA;
  }
  if (dev == cached_dev) return cached_rc;
  int major = 0, minor = 0;
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
  cached_dev = dev;
  if (major != 10) {
    set_error("liblrce_b200 requires an sm_100a (B200) device, found compute capability %d.%d; no fallback exists",
              major, minor);
    cached_rc = LRCE_EARCH;
  } else {
    cached_rc = LRCE_OK;
  }
  return cached_rc;
}

int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: CUDA error %s (%s)", what, cudaGetErrorName(e), cudaGetErrorString(e));
    return LRCE_ECUDA;
  }
  return LRCE_OK;
}

bool pdl_enabled() {
  static const bool on = [] {
    const char* e = getenv("LRCE_B200_PDL");
    return !(e && e[0] == '0');
  }();
  return on;
}

int sm_count() {
  static thread_local int cached_dev = -1;
  static thread_local int cached = 0;
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev != cached_dev) {
    cudaDeviceGetAttribute(&cached, cudaDevAttrMultiProcessorCount, dev);
    cached_dev = dev;
  }
  return cached;
}

int current_device() {
  int dev = -1;
  if (cudaGetDevice(&dev) != cudaSuccess) return -1;
  return dev;
}

bool needs_device_setup(const uint64_t* mask) {
  const int dev = current_device();
  return dev < 0 || dev >= 64 || !((*mask >> dev) & 1ull);
}
void mark_device_setup(uint64_t* mask) {
  const int dev = current_device();
  if (dev >= 0 && dev < 64) *mask |= 1ull << dev;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    if (e == cudaSuccess && q == cudaDriverEntryPointSuccess) fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

int make_tmap_2d_bf16(CUtensorMap* tm, const void* base, uint64_t inner, uint64_t outer, uint64_t ld_elems,
                      uint32_t box_inner, uint32_t box_outer, int swizzle_bytes) {
  EncodeTiledFn enc = get_encode();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled entry point not available from the driver");
    return LRCE_EDRIVER;
  }
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0 || ((ld_elems * 2) & 15) != 0) {
    set_error("TMA operand must be 16-byte aligned with a 16-byte multiple row pitch (base=%p, ld=%llu elems)", base,
              (unsigned long long)ld_elems);
    return LRCE_EINVAL;
  }
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {ld_elems * 2};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d (inner=%llu outer=%llu ld=%llu box=%ux%u)", (int)r,
              (unsigned long long)inner, (unsigned long long)outer, (unsigned long long)ld_elems, box_inner, box_outer);
    return LRCE_EDRIVER;
  }
  return LRCE_OK;
}

int make_tmap_nd_bf16(CUtensorMap* tm, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                      const uint32_t* box, int swizzle_bytes, int l2_promo_bytes) {
  EncodeTiledFn enc = get_encode();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled entry point not available from the driver");
    return LRCE_EDRIVER;
  }
  if (rank < 1 || rank > 5 || (reinterpret_cast<uintptr_t>(base) & 15) != 0) {
    set_error("TMA operand must be 16-byte aligned, rank 1..5 (base=%p rank=%d)", base, rank);
    return LRCE_EINVAL;
  }
  cuuint64_t d[5], st[4];
  cuuint32_t b[5], es[5];
  for (int i = 0; i < rank; ++i) {
    d[i] = dims[i];
    b[i] = box[i];
    es[i] = 1;
    if (i + 1 < rank) {
      st[i] = strides_bytes[i];
      if (st[i] & 15) {
        set_error("TMA strides must be multiples of 16 bytes (stride %d = %llu)", i, (unsigned long long)st[i]);
        return LRCE_EINVAL;
      }
    }
  }
  const CUtensorMapSwizzle sw = swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                                : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_NONE;
  const CUtensorMapL2promotion pr = l2_promo_bytes == 256 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B
                                    : l2_promo_bytes == 128 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B
                                    : l2_promo_bytes == 64 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B : CU_TENSOR_MAP_L2_PROMOTION_NONE;
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, static_cast<cuuint32_t>(rank), const_cast<void*>(base), d, st, b, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, sw, pr, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled (rank %d) failed with CUresult %d", rank, (int)r);
    return LRCE_EDRIVER;
  }
  return LRCE_OK;
}

}  // namespace lrce

extern "C" const char* lrce_last_error(void) { return lrce::last_error(); }
extern "C" int lrce_abi_version(void) { return LRCE_ABI_VERSION; }
