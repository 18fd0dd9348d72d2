// Micro-benchmark (development aid): issue rate of the special-function and packed-fp32 instructions the HSTU SiLU uses.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mufu_bench mufu_bench.cu ; run on the GPU box.
#include <cstdio>
#include <cuda_runtime.h>
enum Op { TANH, EX2, RCP, FMA2, FMA1, CVT, TANH_H2, SILU_TANH, SILU_MIX };
__device__ __forceinline__ float tanh_a(float x) { float y; asm volatile("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float ex2_a(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcp_a(float x) { float y; asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ unsigned cvt2(float a, float b) { unsigned r; asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a)); return r; }
__device__ __forceinline__ unsigned tanh_h2(unsigned x) { unsigned y; asm volatile("tanh.approx.bf16x2 %0, %1;" : "=r"(y) : "r"(x)); return y; }

template <int OP>
__global__ void k(float* out, long long* cyc, int iters) {
  float v[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = threadIdx.x * 0.001f + i * 0.01f;
  unsigned long long w[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) w[i] = ((unsigned long long)__float_as_uint(v[2 * i]) << 32) | __float_as_uint(v[2 * i + 1]);
  unsigned hacc = 0;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      if (OP == TANH) v[i] = tanh_a(v[i]);
      if (OP == EX2) v[i] = ex2_a(v[i]);
      if (OP == RCP) v[i] = rcp_a(v[i]);
      if (OP == FMA1) v[i] = fmaf(v[i], 0.999f, 0.001f);
      if (OP == TANH_H2) { unsigned u = tanh_h2(__float_as_uint(v[i])); v[i] = __uint_as_float(u); }
    }
    if (OP == FMA2) {
#pragma unroll
      for (int i = 0; i < 8; ++i) { w[i] = fma2(w[i], w[(i + 1) & 7], w[i]); }
#pragma unroll
      for (int i = 0; i < 8; ++i) { w[i] = fma2(w[i], w[(i + 3) & 7], w[i]); }
    }
    if (OP == CVT) {
#pragma unroll
      for (int i = 0; i < 16; i += 2) { unsigned u = cvt2(v[i], v[i + 1]); hacc ^= u; v[i] = __uint_as_float(u & 0x3fffffff); }
#pragma unroll
      for (int i = 0; i < 16; i += 2) { unsigned u = cvt2(v[i + 1], v[i]); hacc ^= u; v[i + 1] = __uint_as_float(u & 0x3fffffff); }
    }
    if (OP == SILU_TANH) {   // h + h*tanh(h) for 16 values, packed mul/fma + pack to bf16 (what the fwd kernel does per 16 scores)
#pragma unroll
      for (int i = 0; i < 16; i += 2) {
        float h0 = v[i] * 0.044f, h1 = v[i + 1] * 0.044f;
        float p0 = fmaf(h0, tanh_a(h0), h0), p1 = fmaf(h1, tanh_a(h1), h1);
        unsigned u = cvt2(p0, p1); hacc ^= u; v[i] = p0; v[i + 1] = p1;
      }
    }
    if (OP == SILU_MIX) {    // half the values through tanh, half through ex2 + rcp
#pragma unroll
      for (int i = 0; i < 16; i += 4) {
        float h0 = v[i] * 0.044f, h1 = v[i + 1] * 0.044f;
        float p0 = fmaf(h0, tanh_a(h0), h0), p1 = fmaf(h1, tanh_a(h1), h1);
        float x2 = v[i + 2] * 0.088f, x3 = v[i + 3] * 0.088f;
        float p2 = x2 * rcp_a(1.f + ex2_a(-1.4427f * x2)), p3 = x3 * rcp_a(1.f + ex2_a(-1.4427f * x3));
        unsigned u = cvt2(p0, p1) ^ cvt2(p2, p3); hacc ^= u; v[i] = p0; v[i + 1] = p1; v[i + 2] = p2; v[i + 3] = p3;
      }
    }
  }
  long long t1 = clock64();
  float s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += v[i];
#pragma unroll
  for (int i = 0; i < 8; ++i) s += (float)(w[i] & 0xff);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s + hacc;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int OP>
void run(const char* name, int per_iter_ops) {
  float* out; long long* cyc; cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
  const int iters = 2000;
  for (int threads : {32, 128, 256, 512}) {
    k<OP><<<148, threads>>>(out, cyc, iters); cudaDeviceSynchronize();
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    const double warps_per_smsp = threads / 128.0 < 1 ? 1 : threads / 128.0;
    printf("%-10s threads/SM %4d : %7.2f cycles per warp-instruction per SMSP (%.1f cyc/iter/warp)\n", name, threads,
           (double)c / iters / per_iter_ops / warps_per_smsp, (double)c / iters);
  }
  cudaFree(out); cudaFree(cyc);
}
int main() {
  run<TANH>("tanh", 16); run<EX2>("ex2", 16); run<RCP>("rcp", 16); run<FMA1>("ffma", 16); run<FMA2>("ffma2", 16); run<CVT>("cvt.bf16x2", 16);
  run<TANH_H2>("tanh.bf16x2", 16); run<SILU_TANH>("silu16_tanh", 16); run<SILU_MIX>("silu16_mix", 16);
  return 0;
}
