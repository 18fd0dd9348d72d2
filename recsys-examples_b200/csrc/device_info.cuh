// Host-side device facts shared by every launcher in this library (per device, looked up once; never hard-code 148 SMs).
#pragma once
#include <cuda_runtime.h>

#include <atomic>

namespace devinfo {
constexpr int kMaxDevices = 64;
inline int current_device() { int d = 0; cudaGetDevice(&d); return (d < 0 || d >= kMaxDevices) ? 0 : d; }
// number of SMs of the current device (B200: 148)
inline int sm_count() {
  static std::atomic<int> cache[kMaxDevices];
  const int d = current_device();
  int v = cache[d].load(std::memory_order_relaxed);
  if (v <= 0) {
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, d) != cudaSuccess || v <= 0) v = 148;
    cache[d].store(v, std::memory_order_relaxed);
  }
  return v;
}
// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is per device: run `f` once per device (idempotent, so a race only repeats it)
template <typename F>
inline cudaError_t once_per_device(std::atomic<int>* flags /*[kMaxDevices], zero-initialised*/, F&& f) {
  const int d = current_device();
  if (flags[d].load(std::memory_order_acquire)) return cudaSuccess;
  cudaError_t e = f();
  if (e == cudaSuccess) flags[d].store(1, std::memory_order_release);
  return e;
}
}  // namespace devinfo
