// Host-side TMA tensor-map construction without linking libcuda: cuTensorMapEncodeTiled is resolved through
// cudaGetDriverEntryPoint at first use (the .so must dlopen on machines that have no driver, e.g. the CPU CI box).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace tma {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// bf16 tensor, 3 dims (d0 fastest, contiguous), byte strides of d1 and d2, box sizes, SWIZZLE_128B (box0 * 2 bytes must be 128)
inline int make_map_3d(CUtensorMap* m, const void* base, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t stride1_bytes, uint64_t stride2_bytes,
                       uint32_t box0, uint32_t box1, uint32_t box2) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return -2000;
  cuuint64_t dims[3] = {d0, d1, d2};
  cuuint64_t strides[2] = {stride1_bytes, stride2_bytes};
  cuuint32_t box[3] = {box0, box1, box2};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -2001 - (int)r;
}

// cuTensorMapEncodeTiled is a DRIVER entry point: it fails with CUDA_ERROR_INVALID_CONTEXT on a host thread that has not touched the
// runtime yet (PyTorch's autograd worker calling the backward first).  Bind the primary context of the device that owns `p`.
inline void bind_context(const void* p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) == cudaSuccess && a.type == cudaMemoryTypeDevice) {
    int cur = -1;
    if (cudaGetDevice(&cur) != cudaSuccess || cur != a.device) cudaSetDevice(a.device);
  } else {
    (void)cudaGetLastError();
  }
  cudaFree(nullptr);
}

}  // namespace tma
