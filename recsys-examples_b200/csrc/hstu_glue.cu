// HSTU layer glue: the row-wise elementwise stages between the GEMMs and the attention kernel of one HSTU layer,
//     y = LN(x)                      -> uvqk GEMM -> SiLU -> split u,v,q,k -> attention -> y = dropout(LN(attn) * u) -> proj GEMM + residual
// as hand-written sm_100a kernels.  Replaces the reference's Triton kernels
//     examples/hstu/ops/triton_ops/triton_layer_norm.py      (_weighted_layer_norm_fwd :80, _weighted_layer_norm_bwd_dx :171, _layer_norm_bwd_dwdb :284)
//     examples/hstu/ops/triton_ops/triton_norm_mul_dropout.py (_ln_mul_dropout_fwd :37, _ln_mul_dropout_bwd_dx_du :131, _ln_mul_dropout_bwd_dwdb :332)
//     examples/hstu/ops/triton_ops/triton_silu.py             (_silu_forward :48, _silu_backward :69)
// as they are called from FusedHSTULayerFunction (examples/hstu/ops/fused_hstu_op.py:196-251, :421-483 and its backward).
//
// All of it is HBM-bound: a warp owns a row (hidden <= 1024: the row lives in registers, 16-byte loads, one pass), fp32 math, statistics by
// warp shuffles.  What is fused that the reference runs as separate passes:
//   * LN backward computes dx AND the weight / bias gradient partials in the same pass over dy and x (the reference re-reads both in
//     `_layer_norm_bwd_dwdb`); a residual gradient can be added into dx on the way out (the layer's `+ x`);
//   * LN*u*dropout backward produces dx, du, the dw / db partials and (optionally) the recomputed forward output y in one pass;
//   * SiLU backward reads the four gradient pieces (du, dv, dq, dk) where the producers left them (own pointer / row stride each) and
//     writes the dense [T, 4*H*D] gradient of the uvqk GEMM output: no torch.cat.
// Dropout: counter-based Philox4x32-10 keyed by (seed, row, column group) — forward and backward regenerate the same mask, nothing is
// stored.  The reference draws from Triton's tl.rand(seed, row * BLOCK_D + col); the two streams differ, the distribution does not
// (keep probability quantised to 1/65536 here).
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/hstu_b200.h"
#include "device_info.cuh"

namespace glue {

constexpr int kWarps = 4;
constexpr int kThreads = kWarps * 32;
constexpr int kMaxD = 1024;            // warp-per-row kernels: NV <= 4 vectors of 8 elements per lane; wider rows take the CTA-per-row kernels

#define GLUE_CHECK_LAST() do { cudaError_t e__ = cudaGetLastError(); if (e__ != cudaSuccess) return -(int)e__; } while (0)

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ---- 8 consecutive elements: packed in registers (Raw8) <-> 8 floats ----------------------------------------------------------------------
template <typename T> struct Raw8;
template <> struct Raw8<float> { float4 a, b; };
template <> struct Raw8<__nv_bfloat16> { uint4 r; };
template <> struct Raw8<__half> { uint4 r; };

__device__ __forceinline__ Raw8<float> load_raw(const float* p) {
  Raw8<float> r;
  r.a = *reinterpret_cast<const float4*>(p);
  r.b = *reinterpret_cast<const float4*>(p + 4);
  return r;
}
__device__ __forceinline__ Raw8<__nv_bfloat16> load_raw(const __nv_bfloat16* p) { Raw8<__nv_bfloat16> r; r.r = *reinterpret_cast<const uint4*>(p); return r; }
__device__ __forceinline__ Raw8<__half> load_raw(const __half* p) { Raw8<__half> r; r.r = *reinterpret_cast<const uint4*>(p); return r; }

__device__ __forceinline__ void unpack(const Raw8<float>& r, float* v) {
  v[0] = r.a.x; v[1] = r.a.y; v[2] = r.a.z; v[3] = r.a.w; v[4] = r.b.x; v[5] = r.b.y; v[6] = r.b.z; v[7] = r.b.w;
}
__device__ __forceinline__ void unpack(const Raw8<__nv_bfloat16>& r, float* v) {
  const uint32_t w[4] = {r.r.x, r.r.y, r.r.z, r.r.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {                // bf16 -> fp32 is a 16-bit shift
    v[2 * i] = __uint_as_float(w[i] << 16);
    v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
  }
}
__device__ __forceinline__ void unpack(const Raw8<__half>& r, float* v) {
  const uint32_t w[4] = {r.r.x, r.r.y, r.r.z, r.r.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[i]));
    v[2 * i] = f.x; v[2 * i + 1] = f.y;
  }
}
__device__ __forceinline__ void store8(float* p, const float* v) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
}
__device__ __forceinline__ void store8(__nv_bfloat16* p, const float* v) {
  uint32_t w[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
    w[i] = *reinterpret_cast<const uint32_t*>(&h);
  }
  *reinterpret_cast<uint4*>(p) = make_uint4(w[0], w[1], w[2], w[3]);
}
__device__ __forceinline__ void store8(__half* p, const float* v) {
  uint32_t w[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const __half2 h = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
    w[i] = *reinterpret_cast<const uint32_t*>(&h);
  }
  *reinterpret_cast<uint4*>(p) = make_uint4(w[0], w[1], w[2], w[3]);
}
template <typename T> struct V8 {
  static __device__ __forceinline__ void load(const T* p, float* v) { unpack(load_raw(p), v); }
  static __device__ __forceinline__ void store(T* p, const float* v) { store8(p, v); }
};
// a row's packed vectors: lane `lane` owns columns [(k * 32 + lane) * 8, +8) for k < NV
template <typename T, int NV>
__device__ __forceinline__ void load_row(const T* __restrict__ base, int lane, int D, Raw8<T>* r) {
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int c = (k * 32 + lane) * 8;
    if (c < D) r[k] = load_raw(base + c);
  }
}

// ---- dropout mask ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 philox4x32(uint4 ctr, uint2 key) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
    const uint32_t hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += W0; key.y += W1;
  }
  return ctr;
}
struct Drop {
  uint64_t seed;
  uint32_t thr;       // element kept iff its 16 random bits >= thr;  0 = keep everything (eval / ratio 0)
  float scale;        // 1 / (1 - thr / 65536)
};
// keep-scale of the 8 elements of column group `cg` of `row`: scale or 0
__device__ __forceinline__ void drop_scales(const Drop& d, int64_t row, int cg, float* m) {
  if (d.thr == 0) {
#pragma unroll
    for (int j = 0; j < 8; ++j) m[j] = 1.0f;
    return;
  }
  const uint4 r = philox4x32(make_uint4((uint32_t)cg, (uint32_t)row, (uint32_t)((uint64_t)row >> 32), 0x48535455u),
                             make_uint2((uint32_t)d.seed, (uint32_t)(d.seed >> 32)));
  const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
  for (int j = 0; j < 8; ++j) m[j] = ((w[j >> 1] >> (16 * (j & 1))) & 0xffffu) >= d.thr ? d.scale : 0.0f;
}
static Drop make_drop(float ratio, uint64_t seed, int training) {
  Drop d{seed, 0u, 1.0f};
  if (training && ratio > 0.0f) {
    long t = lroundf(ratio * 65536.0f);
    if (t < 1) t = 1;
    if (t > 65536) t = 65536;
    d.thr = (uint32_t)t;
    d.scale = t >= 65536 ? 0.0f : 1.0f / (1.0f - (float)t / 65536.0f);
  }
  return d;
}

// ---- forward: y = LN(x) [* u, dropout] ---------------------------------------------------------------------------------------------------
// All global loads of a row (x and u) are issued before anything depends on them; weight / bias stay packed in registers.
template <typename T, int NV, bool MUL>
__global__ void __launch_bounds__(kThreads) ln_fwd_kernel(const T* __restrict__ x, int64_t sx, const T* __restrict__ w, const T* __restrict__ b,
                                                          const T* __restrict__ u, int64_t su, T* __restrict__ y, int64_t sy,
                                                          float* __restrict__ mean, float* __restrict__ rstd, int64_t rows, int D, float eps, Drop drop) {
  const int lane = threadIdx.x & 31;
  const float invD = 1.0f / (float)D;
  Raw8<T> wr[NV], br[NV];
  if (w) load_row<T, NV>(w, lane, D, wr);
  if (b) load_row<T, NV>(b, lane, D, br);
  const int64_t wstride = (int64_t)gridDim.x * kWarps;
  constexpr bool PF = false;                // row-ahead prefetch (as in the backward) costs 10-31 registers here (ptxas: 128 -> 138, 162 -> 193) = one
                                            // resident CTA fewer per SM; with 12-16 warps per SM the other warps already cover the load latency: not enabled
  int64_t row = (int64_t)blockIdx.x * kWarps + (threadIdx.x >> 5);
  Raw8<T> rx[NV], ru[NV], nx[NV], nu[NV];
  if (PF && row < rows) {
    load_row<T, NV>(x + row * sx, lane, D, rx);
    if (MUL) load_row<T, NV>(u + row * su, lane, D, ru);
  }
  for (; row < rows; row += wstride) {
    if (PF) {
      const int64_t nrow = row + wstride;
      if (nrow < rows) {
        load_row<T, NV>(x + nrow * sx, lane, D, nx);
        if (MUL) load_row<T, NV>(u + nrow * su, lane, D, nu);
      }
    } else {
      load_row<T, NV>(x + row * sx, lane, D, rx);
      if (MUL) load_row<T, NV>(u + row * su, lane, D, ru);
    }
    float v[NV][8];
    float s = 0.0f;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int c = (k * 32 + lane) * 8;
      if (c < D) {
        unpack(rx[k], v[k]);
#pragma unroll
        for (int j = 0; j < 8; ++j) s += v[k][j];
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) v[k][j] = 0.0f;
      }
    }
    const float m = warp_sum(s) * invD;
    float q = 0.0f;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int c = (k * 32 + lane) * 8;
      if (c < D) {
#pragma unroll
        for (int j = 0; j < 8; ++j) { v[k][j] -= m; q += v[k][j] * v[k][j]; }
      }
    }
    const float r = 1.0f / sqrtf(warp_sum(q) * invD + eps);
    if (lane == 0) { mean[row] = m; rstd[row] = r; }
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int c = (k * 32 + lane) * 8;
      if (c < D) {
        float o[8], wv[8], bv[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) { wv[j] = 1.0f; bv[j] = 0.0f; }
        if (w) unpack(wr[k], wv);
        if (b) unpack(br[k], bv);
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = v[k][j] * r * wv[j] + bv[j];
        if (MUL) {
          float uv[8], ms[8];
          unpack(ru[k], uv);
          drop_scales(drop, row, c >> 3, ms);
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] = o[j] * uv[j] * ms[j];
        }
        store8(y + row * sy + c, o);
      }
    }
    if (PF) {
#pragma unroll
      for (int k = 0; k < NV; ++k) { rx[k] = nx[k]; if (MUL) ru[k] = nu[k]; }
    }
  }
}

// ---- backward: dx (+ residual gradient), [du, y], per-CTA partial dw / db -------------------------------------------------------------------
// part[blockIdx.x][0][D] = sum over this CTA's rows of dln * xhat, part[blockIdx.x][1][D] = sum of dln   (dln = gradient at the LN output)
// Per row: every load is issued up front (x, dy, u, residual gradient), and with PF the NEXT row's x / dy are already in flight while this
// row's two warp reductions and stores run — at ~250 registers per thread only 8 warps fit an SM, so the latency has to be hidden
// inside the warp (PF is off where it would spill: fp32 rows, and the LN*u backward at 1024 columns — moving its dw / db accumulators to
// shared memory was tried and does NOT remove those spills: ptxas stays at 255 registers + 240 B, the pressure is the unrolled per-row
// temporaries, not the accumulators).  The row stays packed (bf16 / fp16) in registers and is unpacked where it is used; dln * w is the one fp32 array kept
// across the reductions.
template <typename T, int NV, bool MUL, bool PF>
__global__ void __launch_bounds__(kThreads) ln_bwd_kernel(const T* __restrict__ dy, int64_t sdy, const T* __restrict__ x, int64_t sx,
                                                          const T* __restrict__ w, const T* __restrict__ b, const float* __restrict__ mean,
                                                          const float* __restrict__ rstd, const T* __restrict__ u, int64_t su,
                                                          const T* __restrict__ dx_add, int64_t sadd, T* __restrict__ dx, int64_t sdx,
                                                          T* __restrict__ du, int64_t sdu, T* __restrict__ y_out, int64_t sy,
                                                          float* __restrict__ part, int64_t rows, int D, Drop drop) {
  extern __shared__ float red[];            // [(kWarps - 1)][D]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float invD = 1.0f / (float)D;
  float dwa[NV][8], dba[NV][8];
#pragma unroll
  for (int k = 0; k < NV; ++k)
#pragma unroll
    for (int j = 0; j < 8; ++j) { dwa[k][j] = 0.0f; dba[k][j] = 0.0f; }
  Raw8<T> wr[NV], br[NV];
  if (w) load_row<T, NV>(w, lane, D, wr);
  if (MUL && b) load_row<T, NV>(b, lane, D, br);
  const int64_t wstride = (int64_t)gridDim.x * kWarps;
  int64_t row = (int64_t)blockIdx.x * kWarps + warp;
  Raw8<T> rx[NV], rdy[NV], nx[NV], ndy[NV];
  float m = 0.0f, r = 0.0f, nm = 0.0f, nr = 0.0f;
  if (row < rows) {
    load_row<T, NV>(x + row * sx, lane, D, rx);
    load_row<T, NV>(dy + row * sdy, lane, D, rdy);
    m = mean[row]; r = rstd[row];
  }
  for (; row < rows; row += wstride) {
    Raw8<T> ru[NV], radd[NV];
    if (MUL) load_row<T, NV>(u + row * su, lane, D, ru);
    if (dx_add) load_row<T, NV>(dx_add + row * sadd, lane, D, radd);
    const int64_t nrow = row + wstride;
    if (PF && nrow < rows) {
      load_row<T, NV>(x + nrow * sx, lane, D, nx);
      load_row<T, NV>(dy + nrow * sdy, lane, D, ndy);
      nm = mean[nrow]; nr = rstd[nrow];
    }
    float gw[NV][8];                        // dln * w
    float c1 = 0.0f, c2 = 0.0f;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int c = (k * 32 + lane) * 8;
      if (c < D) {
        float xh[8], g[8], wv[8];
        unpack(rx[k], xh);
        unpack(rdy[k], g);
#pragma unroll
        for (int j = 0; j < 8; ++j) { xh[j] = (xh[j] - m) * r; wv[j] = 1.0f; }
        if (w) unpack(wr[k], wv);
        if (MUL) {
          float uv[8], ms[8], bv[8], o[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) bv[j] = 0.0f;
          if (b) unpack(br[k], bv);
          unpack(ru[k], uv);
          drop_scales(drop, row, c >> 3, ms);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float ln = xh[j] * wv[j] + bv[j];
            const float gj = g[j] * ms[j];             // gradient at ln * u
            o[j] = gj * ln;                            // du
            g[j] = gj * uv[j];                         // dln
            bv[j] = ln * uv[j] * ms[j];                // y (recomputed forward output)
          }
          store8(du + row * sdu + c, o);
          if (y_out) store8(y_out + row * sy + c, bv);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          dwa[k][j] += g[j] * xh[j];
          dba[k][j] += g[j];
          gw[k][j] = g[j] * wv[j];
          c1 += gw[k][j] * xh[j];
          c2 += gw[k][j];
        }
      }
    }
    c1 = warp_sum(c1) * invD;
    c2 = warp_sum(c2) * invD;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int c = (k * 32 + lane) * 8;
      if (c < D) {
        float xh[8], o[8];
        unpack(rx[k], xh);
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = (gw[k][j] - ((xh[j] - m) * r * c1 + c2)) * r;
        if (dx_add) {
          float a[8];
          unpack(radd[k], a);
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] += a[j];
        }
        store8(dx + row * sdx + c, o);
      }
    }
    if (PF) {
#pragma unroll
      for (int k = 0; k < NV; ++k) { rx[k] = nx[k]; rdy[k] = ndy[k]; }
      m = nm; r = nr;
    } else if (nrow < rows) {
      load_row<T, NV>(x + nrow * sx, lane, D, rx);
      load_row<T, NV>(dy + nrow * sdy, lane, D, rdy);
      m = mean[nrow]; r = rstd[nrow];
    }
  }
  // CTA reduction of the weight / bias gradient partials: warps 1.. hand theirs to warp 0 through shared memory, dw then db
#pragma unroll
  for (int which = 0; which < 2; ++which) {
    if (warp > 0) {
#pragma unroll
      for (int k = 0; k < NV; ++k) {
        const int c = (k * 32 + lane) * 8;
        if (c < D) {
#pragma unroll
          for (int j = 0; j < 8; ++j) red[(warp - 1) * D + c + j] = which == 0 ? dwa[k][j] : dba[k][j];
        }
      }
    }
    __syncthreads();
    if (warp == 0) {
#pragma unroll
      for (int k = 0; k < NV; ++k) {
        const int c = (k * 32 + lane) * 8;
        if (c < D) {
          float o[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            o[j] = which == 0 ? dwa[k][j] : dba[k][j];
#pragma unroll
            for (int ww = 0; ww < kWarps - 1; ++ww) o[j] += red[ww * D + c + j];
          }
          store8(part + ((int64_t)blockIdx.x * 2 + which) * D + c, o);
        }
      }
    }
    __syncthreads();
  }
}

// ---- wide rows (1024 < D <= 8192): one CTA of 256 threads per row ------------------------------------------------------------------------
// Same math and the same column ownership idea as the warp-per-row kernels — thread t owns columns [(k * 256 + t) * 8, +8), k < NVW <= 4, so the
// weight / bias gradient accumulators stay in registers over all rows of the CTA and go straight to its partial row (no cross-warp hand-over) —
// with the row statistics combined through shared memory (two block reductions per row forward, one pair backward).
constexpr int kWideThreads = 256;
constexpr int kMaxWideD = 8192;

// sum over the CTA; `sm` holds 2 x 8 floats and is used alternately (parity) so one __syncthreads per reduction is enough
__device__ __forceinline__ float block_sum(float v, float* sm, int parity) {
  v = warp_sum(v);
  float* s = sm + parity * 8;
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = 0.0f;
#pragma unroll
  for (int i = 0; i < kWideThreads / 32; ++i) t += s[i];
  return t;
}

template <typename T, int NVW>
__device__ __forceinline__ void load_row_wide(const T* __restrict__ base, int D, Raw8<T>* r) {
#pragma unroll
  for (int k = 0; k < NVW; ++k) {
    const int c = (k * kWideThreads + (int)threadIdx.x) * 8;
    if (c < D) r[k] = load_raw(base + c);
  }
}

template <typename T, int NVW, bool MUL>
__global__ void __launch_bounds__(kWideThreads) ln_fwd_wide_kernel(const T* __restrict__ x, int64_t sx, const T* __restrict__ w, const T* __restrict__ b,
                                                                   const T* __restrict__ u, int64_t su, T* __restrict__ y, int64_t sy,
                                                                   float* __restrict__ mean, float* __restrict__ rstd, int64_t rows, int D, float eps,
                                                                   Drop drop) {
  __shared__ float sm[16];
  const float invD = 1.0f / (float)D;
  int parity = 0;
  for (int64_t row = blockIdx.x; row < rows; row += gridDim.x) {
    Raw8<T> rx[NVW], ru[NVW];
    load_row_wide<T, NVW>(x + row * sx, D, rx);
    if (MUL) load_row_wide<T, NVW>(u + row * su, D, ru);
    float v[NVW][8];
    float s = 0.0f;
#pragma unroll
    for (int k = 0; k < NVW; ++k) {
      const int c = (k * kWideThreads + (int)threadIdx.x) * 8;
      if (c < D) {
        unpack(rx[k], v[k]);
#pragma unroll
        for (int j = 0; j < 8; ++j) s += v[k][j];
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) v[k][j] = 0.0f;
      }
    }
    const float m = block_sum(s, sm, parity) * invD;
    parity ^= 1;
    float q = 0.0f;
#pragma unroll
    for (int k = 0; k < NVW; ++k) {
      const int c = (k * kWideThreads + (int)threadIdx.x) * 8;
      if (c < D) {
#pragma unroll
        for (int j = 0; j < 8; ++j) { v[k][j] -= m; q += v[k][j] * v[k][j]; }
      }
    }
    const float r = 1.0f / sqrtf(block_sum(q, sm, parity) * invD + eps);
    parity ^= 1;
    if (threadIdx.x == 0) { mean[row] = m; rstd[row] = r; }
#pragma unroll
    for (int k = 0; k < NVW; ++k) {
      const int c = (k * kWideThreads + (int)threadIdx.x) * 8;
      if (c < D) {
        float o[8], wv[8], bv[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) { wv[j] = 1.0f; bv[j] = 0.0f; }
        if (w) V8<T>::load(w + c, wv);
        if (b) V8<T>::load(b + c, bv);
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = v[k][j] * r * wv[j] + bv[j];
        if (MUL) {
          float uv[8], ms[8];
          unpack(ru[k], uv);
          drop_scales(drop, row, c >> 3, ms);
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] = o[j] * uv[j] * ms[j];
        }
        store8(y + row * sy + c, o);
      }
    }
  }
}

template <typename T, int NVW, bool MUL>
__global__ void __launch_bounds__(kWideThreads) ln_bwd_wide_kernel(const T* __restrict__ dy, int64_t sdy, const T* __restrict__ x, int64_t sx,
                                                                   const T* __restrict__ w, const T* __restrict__ b, const float* __restrict__ mean,
                                                                   const float* __restrict__ rstd, const T* __restrict__ u, int64_t su,
                                                                   const T* __restrict__ dx_add, int64_t sadd, T* __restrict__ dx, int64_t sdx,
                                                                   T* __restrict__ du, int64_t sdu, T* __restrict__ y_out, int64_t sy,
                                                                   float* __restrict__ part, int64_t rows, int D, Drop drop) {
  __shared__ float sm[32];                  // two reductions per row (c1, c2), each with its own pair of parity buffers
  const float invD = 1.0f / (float)D;
  float dwa[NVW][8], dba[NVW][8];
#pragma unroll
  for (int k = 0; k < NVW; ++k)
#pragma unroll
    for (int j = 0; j < 8; ++j) { dwa[k][j] = 0.0f; dba[k][j] = 0.0f; }
  int parity = 0;
  for (int64_t row = blockIdx.x; row < rows; row += gridDim.x) {
    Raw8<T> rx[NVW], rdy[NVW], ru[NVW], radd[NVW];
    load_row_wide<T, NVW>(x + row * sx, D, rx);
    load_row_wide<T, NVW>(dy + row * sdy, D, rdy);
    if (MUL) load_row_wide<T, NVW>(u + row * su, D, ru);
    if (dx_add) load_row_wide<T, NVW>(dx_add + row * sadd, D, radd);
    const float m = mean[row], r = rstd[row];
    float gw[NVW][8];
    float c1 = 0.0f, c2 = 0.0f;
#pragma unroll
    for (int k = 0; k < NVW; ++k) {
      const int c = (k * kWideThreads + (int)threadIdx.x) * 8;
      if (c < D) {
        float xh[8], g[8], wv[8];
        unpack(rx[k], xh);
        unpack(rdy[k], g);
#pragma unroll
        for (int j = 0; j < 8; ++j) { xh[j] = (xh[j] - m) * r; wv[j] = 1.0f; }
        if (w) V8<T>::load(w + c, wv);
        if (MUL) {
          float uv[8], ms[8], bv[8], o[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) bv[j] = 0.0f;
          if (b) V8<T>::load(b + c, bv);
          unpack(ru[k], uv);
          drop_scales(drop, row, c >> 3, ms);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float ln = xh[j] * wv[j] + bv[j];
            const float gj = g[j] * ms[j];
            o[j] = gj * ln;
            g[j] = gj * uv[j];
            bv[j] = ln * uv[j] * ms[j];
          }
          store8(du + row * sdu + c, o);
          if (y_out) store8(y_out + row * sy + c, bv);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          dwa[k][j] += g[j] * xh[j];
          dba[k][j] += g[j];
          gw[k][j] = g[j] * wv[j];
          c1 += gw[k][j] * xh[j];
          c2 += gw[k][j];
        }
      }
    }
    c1 = block_sum(c1, sm, parity) * invD;
    c2 = block_sum(c2, sm + 16, parity) * invD;
    parity ^= 1;
#pragma unroll
    for (int k = 0; k < NVW; ++k) {
      const int c = (k * kWideThreads + (int)threadIdx.x) * 8;
      if (c < D) {
        float xh[8], o[8];
        unpack(rx[k], xh);
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = (gw[k][j] - ((xh[j] - m) * r * c1 + c2)) * r;
        if (dx_add) {
          float a[8];
          unpack(radd[k], a);
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] += a[j];
        }
        store8(dx + row * sdx + c, o);
      }
    }
  }
#pragma unroll
  for (int k = 0; k < NVW; ++k) {
    const int c = (k * kWideThreads + (int)threadIdx.x) * 8;
    if (c < D) {
      store8(part + ((int64_t)blockIdx.x * 2 + 0) * D + c, dwa[k]);
      store8(part + ((int64_t)blockIdx.x * 2 + 1) * D + c, dba[k]);
    }
  }
}

// dw[c] = sum_blocks part[blk][0][c], db[c] = sum_blocks part[blk][1][c].  One CTA = 32 columns x 32 chunk lanes: lane (cx, cy) adds the partials of
// blocks cy, cy + 32, ... for its column (128-byte coalesced across cx), the 32 chunk sums are combined through shared memory in a fixed order
// => bit-reproducible for a given grid.  (First version: one thread per column walking all ~600 partials, 8 CTAs — 49 us, a fifth of the LN
// backward; this shape: 2 * D / 32 CTAs of 1024 threads.)
__global__ void __launch_bounds__(1024) colsum_kernel(const float* __restrict__ part, int nblocks, int D, float* __restrict__ dw, float* __restrict__ db) {
  __shared__ float sm[32][33];
  const int cx = threadIdx.x & 31, cy = threadIdx.x >> 5;
  const int i = blockIdx.x * 32 + cx;                 // index into [dw (D) | db (D)]; a CTA may straddle the two, so `which` is per lane
  float s = 0.0f;
  if (i < 2 * D) {
    const int which = i / D, c = i - which * D;
    for (int blk = cy; blk < nblocks; blk += 32) s += part[((int64_t)blk * 2 + which) * D + c];
  }
  sm[cy][cx] = s;
  __syncthreads();
  if (cy == 0 && i < 2 * D) {
    float t = 0.0f;
#pragma unroll
    for (int k = 0; k < 32; ++k) t += sm[k][cx];
    const int which = i / D, c = i - which * D;
    float* out = which == 0 ? dw : db;
    if (out) out[c] = t;
  }
}

// ---- SiLU -------------------------------------------------------------------------------------------------------------------------------
// sigmoid: fp32 I/O uses expf and an IEEE division (~25 instructions per element: at 4096 columns the kernels were ISSUE-bound, ncu
// sm__throughput 80 %, not memory-bound); 16-bit I/O uses ex2.approx + rcp.approx (relative error ~1e-6, far below one bf16 / fp16 ulp)
template <typename T> __device__ __forceinline__ float sigmoid_t(float x) { return 1.0f / (1.0f + expf(-x)); }
__device__ __forceinline__ float sigmoid_fast(float x) {          // 4 instructions: FMUL, MUFU.EX2, FADD, MUFU.RCP
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * -1.4426950408889634f));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));
  return r;
}
template <> __device__ __forceinline__ float sigmoid_t<__nv_bfloat16>(float x) { return sigmoid_fast(x); }
template <> __device__ __forceinline__ float sigmoid_t<__half>(float x) { return sigmoid_fast(x); }

constexpr int kSiluUnroll = 4;            // independent 16-byte vectors in flight per thread
template <typename T>
__global__ void __launch_bounds__(256) silu_fwd_kernel(const T* __restrict__ x, T* __restrict__ y, int64_t n8) {
  const int64_t step = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < n8; i0 += step * kSiluUnroll) {
    Raw8<T> r[kSiluUnroll];
#pragma unroll
    for (int t = 0; t < kSiluUnroll; ++t) {
      const int64_t i = i0 + t * step;
      if (i < n8) r[t] = load_raw(x + i * 8);
    }
#pragma unroll
    for (int t = 0; t < kSiluUnroll; ++t) {
      const int64_t i = i0 + t * step;
      if (i < n8) {
        float v[8];
        unpack(r[t], v);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = v[j] * sigmoid_t<T>(v[j]);
        store8(y + i * 8, v);
      }
    }
  }
}

struct Segs {                     // the gradient of silu's output arrives in up to 4 column segments, each with its own base / row stride
  const void* p[4];
  int64_t stride[4];
  int begin[5];                   // column range of segment s: [begin[s], begin[s + 1])
  int n;
};
// COLSUM: the launcher makes gridDim.x * 256 a multiple of W / 8, so every vector of a thread lies in the SAME 8 columns; the thread sums its
// dx values per column and leaves them in colpart[(blockIdx.x * 256 + threadIdx.x)][8] — read as a [gridDim.x * 256 / (W / 8)][W] matrix of
// partial column sums, reduced by colsum_rows_kernel.  That is the uvqk bias gradient (dz.sum(0) in the reference, a separate pass over 1 GB).
template <typename T, bool COLSUM>
__global__ void __launch_bounds__(256) silu_bwd_kernel(Segs seg, const T* __restrict__ x, T* __restrict__ dx, int64_t rows, int W,
                                                       float* __restrict__ colpart) {
  const int W8 = W >> 3;
  const int64_t n8 = rows * W8;
  const int64_t step = (int64_t)gridDim.x * blockDim.x;
  constexpr int U = 2;                      // two (x, dy) vector pairs in flight per thread
  float bacc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) bacc[j] = 0.0f;
  for (int64_t i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < n8; i0 += step * U) {
    Raw8<T> rx[U], rg[U];
#pragma unroll
    for (int t = 0; t < U; ++t) {
      const int64_t i = i0 + t * step;
      if (i < n8) {
        const int64_t row = i / W8;
        const int c = (int)(i - row * W8) * 8;
        const void* gp = seg.p[0];            // segment select by compares: a dynamic index would copy the parameter struct to local memory
        int64_t gs = seg.stride[0];
        int gb = 0;
#pragma unroll
        for (int q = 1; q < 4; ++q)
          if (q < seg.n && c >= seg.begin[q]) { gp = seg.p[q]; gs = seg.stride[q]; gb = seg.begin[q]; }
        rx[t] = load_raw(x + i * 8);
        rg[t] = load_raw(reinterpret_cast<const T*>(gp) + row * gs + (c - gb));
      }
    }
#pragma unroll
    for (int t = 0; t < U; ++t) {
      const int64_t i = i0 + t * step;
      if (i < n8) {
        float v[8], g[8];
        unpack(rx[t], v);
        unpack(rg[t], g);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float sg = sigmoid_t<T>(v[j]);
          v[j] = g[j] * sg * (1.0f + v[j] * (1.0f - sg));
          if (COLSUM) bacc[j] += v[j];
        }
        store8(dx + i * 8, v);
      }
    }
  }
  if (COLSUM) store8(colpart + ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 8, bacc);
}

// out[c] = sum_r part[r][c] for a row-major [nrows][ncols] matrix of partial sums; 32 columns x 32 row lanes per CTA, fixed combine order
__global__ void __launch_bounds__(1024) colsum_rows_kernel(const float* __restrict__ part, int nrows, int ncols, float* __restrict__ out) {
  __shared__ float sm[32][33];
  const int cx = threadIdx.x & 31, cy = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cx;
  float s = 0.0f;
  if (c < ncols)
    for (int r = cy; r < nrows; r += 32) s += part[(int64_t)r * ncols + c];
  sm[cy][cx] = s;
  __syncthreads();
  if (cy == 0 && c < ncols) {
    float t = 0.0f;
#pragma unroll
    for (int k = 0; k < 32; ++k) t += sm[k][cx];
    out[c] = t;
  }
}

__global__ void dropout_mask_kernel(Drop drop, int64_t rows, int D, uint8_t* __restrict__ keep) {
  const int D8 = D >> 3;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < rows * D8; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = i / D8;
    const int cg = (int)(i - row * D8);
    float m[8];
    drop_scales(drop, row, cg, m);
#pragma unroll
    for (int j = 0; j < 8; ++j) keep[row * D + cg * 8 + j] = m[j] != 0.0f;
  }
}

// ---- launch helpers ---------------------------------------------------------------------------------------------------------------------
inline int row_grid(int64_t rows, int per_sm) {
  const int64_t want = (rows + kWarps - 1) / kWarps;
  const int64_t cap = (int64_t)devinfo::sm_count() * per_sm;
  return (int)(want < cap ? (want < 1 ? 1 : want) : cap);
}
inline int bwd_grid_max() { return devinfo::sm_count() * 4; }
inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

#define GLUE_DISPATCH_NV(D, ...)                         \
  do {                                                   \
    if ((D) <= 256) { constexpr int NV = 1; __VA_ARGS__; } \
    else if ((D) <= 512) { constexpr int NV = 2; __VA_ARGS__; } \
    else { constexpr int NV = 4; __VA_ARGS__; }          \
  } while (0)
#define GLUE_DISPATCH_T(dtype, ...)                                        \
  do {                                                                     \
    if ((dtype) == 0) { using T = float; __VA_ARGS__; }                    \
    else if ((dtype) == 1) { using T = __half; __VA_ARGS__; }              \
    else { using T = __nv_bfloat16; __VA_ARGS__; }                         \
  } while (0)

#define GLUE_DISPATCH_NVW(D, ...)                          \
  do {                                                     \
    if ((D) <= 2048) { constexpr int NVW = 1; __VA_ARGS__; } \
    else if ((D) <= 4096) { constexpr int NVW = 2; __VA_ARGS__; } \
    else { constexpr int NVW = 4; __VA_ARGS__; }           \
  } while (0)
inline int wide_grid(int64_t rows, int per_sm) {
  const int64_t cap = (int64_t)devinfo::sm_count() * per_sm;
  return (int)(rows < cap ? (rows < 1 ? 1 : rows) : cap);
}

static int check_rows(int64_t rows, int D, int dtype) {
  if (rows < 0 || D <= 0 || (D & 7) || dtype < 0 || dtype > 2) return HSTU_ERR_ARG;
  if (D > kMaxWideD) return HSTU_ERR_UNSUPPORTED;
  return 0;
}

}  // namespace glue

using namespace glue;

extern "C" int64_t hstu_glue_workspace_bytes(int D) { return (int64_t)bwd_grid_max() * 2 * (int64_t)(D > 0 ? D : 0) * 4 + 256; }

extern "C" int hstu_layer_norm_fwd(const void* x, int64_t x_stride, const void* weight, const void* bias, void* y, int64_t y_stride, float* mean,
                                   float* rstd, int64_t rows, int D, float eps, int dtype, void* stream) {
  if (int rc = check_rows(rows, D, dtype)) return rc;
  if (rows == 0) return 0;
  if (!x || !y || !mean || !rstd || (x_stride & 7) || (y_stride & 7) || !aligned16(x) || !aligned16(y)) return HSTU_ERR_ARG;
  const Drop nodrop{0, 0u, 1.0f};
  if (D > kMaxD) {
    GLUE_DISPATCH_T(dtype, GLUE_DISPATCH_NVW(D, (ln_fwd_wide_kernel<T, NVW, false><<<wide_grid(rows, 8), kWideThreads, 0, (cudaStream_t)stream>>>(
        (const T*)x, x_stride, (const T*)weight, (const T*)bias, nullptr, 0, (T*)y, y_stride, mean, rstd, rows, D, eps, nodrop))));
    GLUE_CHECK_LAST();
    return 0;
  }
  GLUE_DISPATCH_T(dtype, GLUE_DISPATCH_NV(D, (ln_fwd_kernel<T, NV, false><<<row_grid(rows, 8), kThreads, 0, (cudaStream_t)stream>>>(
      (const T*)x, x_stride, (const T*)weight, (const T*)bias, nullptr, 0, (T*)y, y_stride, mean, rstd, rows, D, eps, nodrop))));
  GLUE_CHECK_LAST();
  return 0;
}

extern "C" int hstu_ln_mul_dropout_fwd(const void* x, int64_t x_stride, const void* u, int64_t u_stride, const void* weight, const void* bias, void* y,
                                       int64_t y_stride, float* mean, float* rstd, int64_t rows, int D, float eps, float dropout_ratio,
                                       uint64_t seed, int training, int dtype, void* stream) {
  if (int rc = check_rows(rows, D, dtype)) return rc;
  if (rows == 0) return 0;
  if (!x || !u || !y || !mean || !rstd || (x_stride & 7) || (u_stride & 7) || (y_stride & 7) || !aligned16(x) || !aligned16(u) || !aligned16(y) ||
      dropout_ratio < 0.0f || dropout_ratio > 1.0f)
    return HSTU_ERR_ARG;
  if (rows == 0) return 0;
  const Drop drop = make_drop(dropout_ratio, seed, training);
  if (D > kMaxD) {
    GLUE_DISPATCH_T(dtype, GLUE_DISPATCH_NVW(D, (ln_fwd_wide_kernel<T, NVW, true><<<wide_grid(rows, 8), kWideThreads, 0, (cudaStream_t)stream>>>(
        (const T*)x, x_stride, (const T*)weight, (const T*)bias, (const T*)u, u_stride, (T*)y, y_stride, mean, rstd, rows, D, eps, drop))));
    GLUE_CHECK_LAST();
    return 0;
  }
  GLUE_DISPATCH_T(dtype, GLUE_DISPATCH_NV(D, (ln_fwd_kernel<T, NV, true><<<row_grid(rows, 8), kThreads, 0, (cudaStream_t)stream>>>(
      (const T*)x, x_stride, (const T*)weight, (const T*)bias, (const T*)u, u_stride, (T*)y, y_stride, mean, rstd, rows, D, eps, drop))));
  GLUE_CHECK_LAST();
  return 0;
}

static int ln_bwd_launch(bool mul, const void* dy, int64_t sdy, const void* x, int64_t sx, const void* w, const void* b, const float* mean,
                         const float* rstd, const void* u, int64_t su, const void* dx_add, int64_t sadd, void* dx, int64_t sdx, void* du, int64_t sdu,
                         void* y_out, int64_t sy, float* dw, float* db, void* workspace, int64_t ws_bytes, int64_t rows, int D, Drop drop, int dtype,
                         cudaStream_t st) {
  const bool wide = D > kMaxD;
  const int grid = wide ? wide_grid(rows, 4) : row_grid(rows, 4);
  if (!workspace || ws_bytes < (int64_t)grid * 2 * D * 4) return HSTU_ERR_WORKSPACE;
  float* part = reinterpret_cast<float*>(workspace);
  const size_t smem = (size_t)(kWarps - 1) * D * sizeof(float);
  if (wide) {
    if (mul) {
      GLUE_DISPATCH_T(dtype, GLUE_DISPATCH_NVW(D, (ln_bwd_wide_kernel<T, NVW, true><<<grid, kWideThreads, 0, st>>>(
          (const T*)dy, sdy, (const T*)x, sx, (const T*)w, (const T*)b, mean, rstd, (const T*)u, su, (const T*)dx_add, sadd, (T*)dx, sdx, (T*)du, sdu,
          (T*)y_out, sy, part, rows, D, drop))));
    } else {
      GLUE_DISPATCH_T(dtype, GLUE_DISPATCH_NVW(D, (ln_bwd_wide_kernel<T, NVW, false><<<grid, kWideThreads, 0, st>>>(
          (const T*)dy, sdy, (const T*)x, sx, (const T*)w, (const T*)b, mean, rstd, nullptr, 0, (const T*)dx_add, sadd, (T*)dx, sdx, nullptr, 0,
          nullptr, 0, part, rows, D, drop))));
    }
  } else if (mul) {
    GLUE_DISPATCH_T(dtype, GLUE_DISPATCH_NV(D, (ln_bwd_kernel<T, NV, true, (sizeof(T) == 2 && NV <= 2)><<<grid, kThreads, smem, st>>>(
        (const T*)dy, sdy, (const T*)x, sx, (const T*)w, (const T*)b, mean, rstd, (const T*)u, su, (const T*)dx_add, sadd, (T*)dx, sdx, (T*)du, sdu,
        (T*)y_out, sy, part, rows, D, drop))));
  } else {
    GLUE_DISPATCH_T(dtype, GLUE_DISPATCH_NV(D, (ln_bwd_kernel<T, NV, false, (sizeof(T) == 2)><<<grid, kThreads, smem, st>>>(
        (const T*)dy, sdy, (const T*)x, sx, (const T*)w, (const T*)b, mean, rstd, nullptr, 0, (const T*)dx_add, sadd, (T*)dx, sdx, nullptr, 0,
        nullptr, 0, part, rows, D, drop))));
  }
  GLUE_CHECK_LAST();
  if (dw || db) {
    colsum_kernel<<<(2 * D + 31) / 32, 1024, 0, st>>>(part, grid, D, dw, db);
    GLUE_CHECK_LAST();
  }
  return 0;
}

extern "C" int hstu_layer_norm_bwd(const void* dy, int64_t dy_stride, const void* x, int64_t x_stride, const void* weight, const float* mean,
                                   const float* rstd, const void* dx_add, int64_t dx_add_stride, void* dx, int64_t dx_stride, float* dweight,
                                   float* dbias, void* workspace, int64_t workspace_bytes, int64_t rows, int D, int dtype, void* stream) {
  if (int rc = check_rows(rows, D, dtype)) return rc;
  if (rows == 0) {
    if (dweight) cudaMemsetAsync(dweight, 0, (size_t)D * 4, (cudaStream_t)stream);
    if (dbias) cudaMemsetAsync(dbias, 0, (size_t)D * 4, (cudaStream_t)stream);
    return 0;
  }
  if (!dy || !x || !dx || !mean || !rstd || (dy_stride & 7) || (x_stride & 7) || (dx_stride & 7) || (dx_add && (dx_add_stride & 7)) || !aligned16(dy) ||
      !aligned16(x) || !aligned16(dx) || !aligned16(dx_add))
    return HSTU_ERR_ARG;
  return ln_bwd_launch(false, dy, dy_stride, x, x_stride, weight, nullptr, mean, rstd, nullptr, 0, dx_add, dx_add_stride, dx, dx_stride, nullptr, 0,
                       nullptr, 0, dweight, dbias, workspace, workspace_bytes, rows, D, Drop{0, 0u, 1.0f}, dtype, (cudaStream_t)stream);
}

extern "C" int hstu_ln_mul_dropout_bwd(const void* dy, int64_t dy_stride, const void* x, int64_t x_stride, const void* u, int64_t u_stride,
                                       const void* weight, const void* bias, const float* mean, const float* rstd, void* dx, int64_t dx_stride,
                                       void* du, int64_t du_stride, void* y_out, int64_t y_stride, float* dweight, float* dbias, void* workspace,
                                       int64_t workspace_bytes, int64_t rows, int D, float dropout_ratio, uint64_t seed, int training, int dtype,
                                       void* stream) {
  if (int rc = check_rows(rows, D, dtype)) return rc;
  if (rows == 0) {
    if (dweight) cudaMemsetAsync(dweight, 0, (size_t)D * 4, (cudaStream_t)stream);
    if (dbias) cudaMemsetAsync(dbias, 0, (size_t)D * 4, (cudaStream_t)stream);
    return 0;
  }
  if (!dy || !x || !u || !dx || !du || !mean || !rstd || (dy_stride & 7) || (x_stride & 7) || (u_stride & 7) || (dx_stride & 7) || (du_stride & 7) ||
      (y_out && (y_stride & 7)) || !aligned16(dy) || !aligned16(x) || !aligned16(u) || !aligned16(dx) || !aligned16(du) || !aligned16(y_out) ||
      dropout_ratio < 0.0f || dropout_ratio > 1.0f)
    return HSTU_ERR_ARG;
  return ln_bwd_launch(true, dy, dy_stride, x, x_stride, weight, bias, mean, rstd, u, u_stride, nullptr, 0, dx, dx_stride, du, du_stride, y_out,
                       y_stride, dweight, dbias, workspace, workspace_bytes, rows, D, make_drop(dropout_ratio, seed, training), dtype,
                       (cudaStream_t)stream);
}

extern "C" int hstu_silu_fwd(const void* x, void* y, int64_t n, int dtype, void* stream) {
  if (n < 0 || (n & 7) || dtype < 0 || dtype > 2 || (n && (!x || !y)) || !aligned16(x) || !aligned16(y)) return HSTU_ERR_ARG;
  if (n == 0) return 0;
  const int64_t n8 = n >> 3;
  const int64_t want = (n8 + 255) / 256, cap = (int64_t)devinfo::sm_count() * 16;
  GLUE_DISPATCH_T(dtype, (silu_fwd_kernel<T><<<(int)(want < cap ? want : cap), 256, 0, (cudaStream_t)stream>>>((const T*)x, (T*)y, n8)));
  GLUE_CHECK_LAST();
  return 0;
}

// grid for the column-sum variant: a multiple of q = (W/8) / gcd(W/8, 256) blocks, so that gridDim.x * 256 is a multiple of W/8; 0 = not possible
static int silu_colsum_grid(int W, int64_t n8) {
  const int W8 = W >> 3;
  int a = W8, b = 256;
  while (b) { const int t = a % b; a = b; b = t; }
  const int64_t q = W8 / a, cap = (int64_t)devinfo::sm_count() * 16;
  if (q > cap) return 0;
  int64_t want = ((n8 + 255) / 256 + q - 1) / q * q;
  const int64_t most = cap / q * q;
  if (want > most) want = most;
  return (int)(want < q ? q : want);
}
extern "C" int64_t hstu_silu_bwd_bias_workspace_bytes(int W) {
  if (W <= 0 || (W & 7)) return 0;
  const int g = silu_colsum_grid(W, (int64_t)1 << 40);
  return g == 0 ? 0 : (int64_t)g * 256 * 8 * 4 + 256;
}

static int silu_bwd_impl(int num_segments, const void* const* seg_ptr, const int64_t* seg_stride, const int32_t* seg_width, const void* x, void* dx,
                         float* dbias, void* workspace, int64_t workspace_bytes, int64_t rows, int dtype, void* stream) {
  if (num_segments < 1 || num_segments > 4 || !seg_ptr || !seg_stride || !seg_width || rows < 0 || dtype < 0 || dtype > 2) return HSTU_ERR_ARG;
  Segs s{};
  s.n = num_segments;
  int W = 0;
  for (int i = 0; i < 4; ++i) {
    s.begin[i] = W;
    if (i < num_segments) {
      if (!seg_ptr[i] || seg_width[i] <= 0 || (seg_width[i] & 7) || (seg_stride[i] & 7) || !aligned16(seg_ptr[i])) return HSTU_ERR_ARG;
      s.p[i] = seg_ptr[i];
      s.stride[i] = seg_stride[i];
      W += seg_width[i];
    }
  }
  s.begin[4] = W;
  for (int i = num_segments; i < 5; ++i) s.begin[i] = W;
  cudaStream_t st = (cudaStream_t)stream;
  if (rows == 0) {
    if (dbias) cudaMemsetAsync(dbias, 0, (size_t)W * 4, st);
    return 0;
  }
  if (!x || !dx || !aligned16(x) || !aligned16(dx)) return HSTU_ERR_ARG;
  const int64_t n8 = rows * (W >> 3);
  if (!dbias) {
    const int64_t want = (n8 + 255) / 256, cap = (int64_t)devinfo::sm_count() * 16;
    GLUE_DISPATCH_T(dtype, (silu_bwd_kernel<T, false><<<(int)(want < cap ? want : cap), 256, 0, st>>>(s, (const T*)x, (T*)dx, rows, W, nullptr)));
    GLUE_CHECK_LAST();
    return 0;
  }
  const int grid = silu_colsum_grid(W, n8);
  if (grid == 0) return HSTU_ERR_UNSUPPORTED;
  if (!workspace || !aligned16(workspace) || workspace_bytes < (int64_t)grid * 256 * 8 * 4) return HSTU_ERR_WORKSPACE;
  float* colpart = reinterpret_cast<float*>(workspace);
  GLUE_DISPATCH_T(dtype, (silu_bwd_kernel<T, true><<<grid, 256, 0, st>>>(s, (const T*)x, (T*)dx, rows, W, colpart)));
  GLUE_CHECK_LAST();
  colsum_rows_kernel<<<(W + 31) / 32, 1024, 0, st>>>(colpart, (int)((int64_t)grid * 256 / (W >> 3)), W, dbias);
  GLUE_CHECK_LAST();
  return 0;
}

extern "C" int hstu_silu_bwd(int num_segments, const void* const* seg_ptr, const int64_t* seg_stride, const int32_t* seg_width, const void* x, void* dx,
                             int64_t rows, int dtype, void* stream) {
  return silu_bwd_impl(num_segments, seg_ptr, seg_stride, seg_width, x, dx, nullptr, nullptr, 0, rows, dtype, stream);
}

extern "C" int hstu_silu_bwd_bias(int num_segments, const void* const* seg_ptr, const int64_t* seg_stride, const int32_t* seg_width, const void* x,
                                  void* dx, float* dbias, void* workspace, int64_t workspace_bytes, int64_t rows, int dtype, void* stream) {
  if (!dbias) return HSTU_ERR_ARG;
  return silu_bwd_impl(num_segments, seg_ptr, seg_stride, seg_width, x, dx, dbias, workspace, workspace_bytes, rows, dtype, stream);
}

extern "C" int hstu_dropout_mask(int64_t rows, int D, float dropout_ratio, uint64_t seed, uint8_t* keep, void* stream) {
  if (rows < 0 || D <= 0 || (D & 7) || !keep || dropout_ratio < 0.0f || dropout_ratio > 1.0f) return HSTU_ERR_ARG;
  if (rows == 0) return 0;
  const int64_t n = rows * (D >> 3);
  const int64_t want = (n + 255) / 256, cap = (int64_t)devinfo::sm_count() * 16;
  dropout_mask_kernel<<<(int)(want < cap ? want : cap), 256, 0, (cudaStream_t)stream>>>(make_drop(dropout_ratio, seed, 1), rows, D, keep);
  GLUE_CHECK_LAST();
  return 0;
}
