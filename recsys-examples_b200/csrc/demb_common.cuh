// Shared device helpers for the DynamicEmb sm_100a kernels.
//
// Table image contract (must stay byte-identical to the reference so that a table written by
// either implementation can be read by the other): reference
// corelib/dynamicemb/src/table_operation/types.cuh:242-284 and
// corelib/dynamicemb/dynamicemb/scored_hashtable.py:378-425.
//   bucket b (C slots) lives at storage + b * C * (8 + 1 + 8*ns):
//     keys   [C] u64   at +0
//     digests[C] u8    at +8C
//     scores [C][ns] u64 at +9C   (AoS per key; eviction ranks by word ns-1)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace demb {

constexpr uint64_t kEmptyKey = 0xFFFFFFFFFFFFFFFFull;    // types.cuh:117
constexpr uint64_t kLockedKey = 0xFFFFFFFFFFFFFFFDull;   // types.cuh:118
constexpr uint64_t kReclaimKey = 0xFFFFFFFFFFFFFFFEull;  // types.cuh:119
constexpr uint64_t kReserveMask = 0xFFFFFFFFFFFFFFFCull; // types.cuh:121

// ABI enums — numeric values are part of the boundary (score.cuh:30-42, types.cuh:52-61).
enum Policy : int { kConst = 0, kAssign = 1, kAccumulate = 2, kGlobalTimer = 3, kLruLfu = 4 };
enum InsertResult : uint8_t { kInsert = 0, kReclaim = 1, kAssignHit = 2, kEvict = 3, kDuplicated = 4, kBusy = 5, kIllegal = 6, kInit = 7 };

__host__ __device__ __forceinline__ uint64_t fmix64(uint64_t k) {
  k ^= k >> 33; k *= 0xff51afd7ed558ccdull;
  k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ull;
  k ^= k >> 33;
  return k;
}
// types.cuh:123-131
__host__ __device__ __forceinline__ int64_t hash63(uint64_t key) { return (int64_t)(fmix64(key) & 0x7FFFFFFFFFFFFFFFull); }
__host__ __device__ __forceinline__ uint8_t digest_of(int64_t h) { return (uint8_t)(h >> 32); }
__host__ __device__ __forceinline__ bool key_is_valid(uint64_t key) { return (key & kReserveMask) != kReserveMask; }
__host__ __device__ __forceinline__ uint8_t empty_digest() { return digest_of(hash63(kEmptyKey)); }

struct Table {
  uint8_t* storage;
  const int64_t* bkt_off;  // [T+1] device, global bucket offsets per logical table
  int64_t C;               // bucket capacity (multiple of 16)
  int ns;                  // score words per key
  __device__ __forceinline__ int64_t bucket_bytes() const { return C * (9 + 8 * (int64_t)ns); }
  __device__ __forceinline__ uint8_t* bucket(int64_t b) const { return storage + b * bucket_bytes(); }
  __device__ __forceinline__ uint64_t* keys(uint8_t* bk) const { return reinterpret_cast<uint64_t*>(bk); }
  __device__ __forceinline__ uint8_t* digests(uint8_t* bk) const { return bk + 8 * C; }
  __device__ __forceinline__ uint64_t* scores(uint8_t* bk, int64_t it) const { return reinterpret_cast<uint64_t*>(bk + 9 * C) + it * ns; }
};

// Where a key lives: global bucket, table-local slot base, start slot. cap==0 => illegal/absent table.
struct Locus { int64_t h; int64_t bucket; int64_t bkt_begin; int64_t cap; };
__device__ __forceinline__ Locus locate(const Table& t, uint64_t key, int64_t tid) {
  Locus L{0, 0, 0, 0};
  if (key_is_valid(key)) {
    L.h = hash63(key);
    L.bkt_begin = t.bkt_off[tid];
    L.cap = (t.bkt_off[tid + 1] - L.bkt_begin) * t.C;
    if (L.cap > 0) L.bucket = L.bkt_begin + (L.h % L.cap) / t.C;   // kernels.cuh:108-116
  }
  return L;
}

__device__ __forceinline__ uint64_t globaltimer() { uint64_t t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }

// Streaming 16-byte accesses (rows are touched once per step; keep them out of L1).
__device__ __forceinline__ float4 ld_nc_f4(const float* p) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ float4 ld_f4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st_cs_f4(float* p, float4 v) {
  asm volatile("st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void st_f4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }

// Thread-per-key probe of one bucket. Order restated from types.cuh:325-396: 16-aligned start,
// 16 digests per step; inside each 4-byte group all digest matches (ascending byte) are key-checked
// before any empty-digest byte; a confirmed EmptyKey ends the probe.  Returns slot or -1.
// `empty_out` (optional) receives the first empty slot in probe order, or -1.
__device__ __forceinline__ int64_t probe_thread(const Table& t, uint8_t* bk, uint64_t key, int64_t h, int64_t* empty_out) {
  const uint64_t* keys = t.keys(bk);
  const uint8_t* dg = t.digests(bk);
  const uint32_t want = (uint32_t)digest_of(h) * 0x01010101u;
  const uint32_t emp = (uint32_t)empty_digest() * 0x01010101u;
  int64_t it = (h % t.C) & ~(int64_t)15;
  if (empty_out) *empty_out = -1;
  for (int64_t step = 0; step < t.C; step += 16) {
    uint4 buf = *reinterpret_cast<const uint4*>(dg + it);
    uint32_t w[4] = {buf.x, buf.y, buf.z, buf.w};
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      uint32_t m = __vcmpeq4(w[g], want) & 0x01010101u;
      while (m) {
        int o = (__ffs(m) - 1) >> 3; m &= m - 1;
        int64_t p = it + g * 4 + o;
        if (__ldcv(keys + p) == key) return p;
      }
      m = __vcmpeq4(w[g], emp) & 0x01010101u;
      while (m) {
        int o = (__ffs(m) - 1) >> 3; m &= m - 1;
        int64_t p = it + g * 4 + o;
        if (__ldcv(keys + p) == kEmptyKey) { if (empty_out) *empty_out = p; return -1; }
      }
    }
    it += 16; if (it >= t.C) it = 0;
  }
  return -1;
}

}  // namespace demb

#include "device_info.cuh"
namespace demb { using devinfo::kMaxDevices; using devinfo::current_device; using devinfo::sm_count; using devinfo::once_per_device; }

#define DEMB_CHECK_LAST() do { cudaError_t e__ = cudaGetLastError(); if (e__ != cudaSuccess) return -(int)e__; } while (0)
