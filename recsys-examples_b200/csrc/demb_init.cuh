// Row initializers shared by demb_rows.cu (demb_init_rows) and demb_train.cu (fused prefetch).
#pragma once
#include "../../include/dynamicemb_b200.h"
#include "demb_common.cuh"

namespace demb {

// Philox4x32-10 keyed by (seed, key): the value of a row depends only on (seed, key, column), not on
// which thread or batch position initialises it (the reference draws from a pool of per-thread
// curand states, initializer.cu:26-112, so its values depend on launch geometry).
__device__ __forceinline__ uint4 philox4x32(uint4 ctr, uint2 key) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
    uint32_t hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += W0; key.y += W1;
  }
  return ctr;
}
__device__ __forceinline__ float u01(uint32_t x) { return ((float)(x >> 8) + 0.5f) * (1.0f / 16777216.0f); }   // (0,1)
__device__ __forceinline__ void boxmuller(uint32_t a, uint32_t b, float& z0, float& z1) {
  float r = sqrtf(-2.0f * logf(u01(a))), th = 6.28318530717958647692f * u01(b);
  z0 = r * cosf(th); z1 = r * sinf(th);
}

struct InitArgs { int mode; float p0, p1, p2, p3; uint64_t seed; };   // uniform(lower,upper) normal(mean,std) trunc(mean,std,lower,upper) const(value)
static_assert(sizeof(InitArgs) == sizeof(demb_init_args_t) && sizeof(InitArgs) == 32, "InitArgs is the device image of demb_init_args_t");

__device__ __forceinline__ float4 init4(const InitArgs& a, uint64_t key, int c /*float4 chunk*/) {
  if (a.mode == DEMB_INIT_CONSTANT) return make_float4(a.p0, a.p0, a.p0, a.p0);
  if (a.mode == DEMB_INIT_DEBUG) { float v = (float)(key % 100000ull); return make_float4(v, v, v, v); }   // initializer.cuh:142-156
  uint4 rnd = philox4x32(make_uint4((uint32_t)c, 0u, (uint32_t)key, (uint32_t)(key >> 32)), make_uint2((uint32_t)a.seed, (uint32_t)(a.seed >> 32)));
  if (a.mode == DEMB_INIT_UNIFORM) {
    float lo = a.p0, w = a.p1 - a.p0;
    return make_float4(lo + w * u01(rnd.x), lo + w * u01(rnd.y), lo + w * u01(rnd.z), lo + w * u01(rnd.w));
  }
  float z[4];
  boxmuller(rnd.x, rnd.y, z[0], z[1]); boxmuller(rnd.z, rnd.w, z[2], z[3]);
  if (a.mode == DEMB_INIT_TRUNCATED_NORMAL) {
    // resample from further Philox counters until inside [lower, upper] (initializer.cuh truncated normal rejects likewise)
    uint32_t extra = 1;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float v = a.p0 + a.p1 * z[k];
      while (v < a.p2 || v > a.p3) {
        uint4 r2 = philox4x32(make_uint4((uint32_t)c, extra++ * 4u + (uint32_t)k, (uint32_t)key, (uint32_t)(key >> 32)), make_uint2((uint32_t)a.seed, (uint32_t)(a.seed >> 32)));
        float y0, y1; boxmuller(r2.x, r2.y, y0, y1);
        v = a.p0 + a.p1 * y0;
        if (v < a.p2 || v > a.p3) v = a.p0 + a.p1 * y1;
        if (extra > 64) { v = fminf(fmaxf(v, a.p2), a.p3); }
      }
      z[k] = (v - a.p0) / (a.p1 == 0.f ? 1.f : a.p1);
    }
  }
  return make_float4(a.p0 + a.p1 * z[0], a.p0 + a.p1 * z[1], a.p0 + a.p1 * z[2], a.p0 + a.p1 * z[3]);
}


}  // namespace demb
