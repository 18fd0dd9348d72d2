// Micro-benchmark (development aid): issue-to-completion rate of tcgen05.mma kind::f16 by shape and operand source.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../recsys-examples_b200/csrc -o umma_bench umma_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "sm100_ptx.cuh"
using namespace sm100;

struct Cfg { int N; int ts; int b_mn; int n_acc; };

__global__ void __launch_bounds__(128) k(Cfg c, int reps, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base;
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (threadIdx.x < 32) tmem_alloc<512>(&tmem_base);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base;
  if (threadIdx.x < 32 && elect_one()) {
    const uint32_t idesc = umma_idesc_bf16(128, c.N, 0, c.b_mn);
    const uint32_t aA = smem_u32(smem), aB = smem_u32(smem) + 32768;
    long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
      const uint32_t d = tmem + (c.n_acc > 1 ? (r & 1) * 256 : 0);
#pragma unroll
      for (int kk = 0; kk < 8; ++kk) {
        const uint64_t db = c.b_mn ? umma_desc_sw128(aB + kk * 2048, 16384, 1024) : umma_desc_sw128(aB + (kk >> 2) * (c.N * 128) + (kk & 3) * 32, 16, 1024);
        if (c.ts) umma_ts(d, tmem + 448 + kk * 8, db, idesc, 1);
        else umma_ss(d, umma_desc_sw128(aA + (kk >> 2) * 16384 + (kk & 3) * 32, 16, 1024), db, idesc, 1);
      }
    }
    umma_commit(&bar);
    mbar_wait(&bar, 0);
    long long t1 = clock64();
    if (blockIdx.x == 0) *out = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc<512>(tmem);
}

int main() {
  long long* d; cudaMalloc(&d, 8);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const Cfg cfgs[] = {{128, 0, 0, 1}, {128, 0, 1, 1}, {128, 1, 1, 1}, {128, 1, 0, 1}, {256, 0, 0, 1}, {256, 1, 0, 1}, {64, 0, 1, 1}, {64, 1, 1, 1}, {128, 0, 0, 2}, {128, 1, 1, 2}, {192, 0, 0, 1}};
  for (const Cfg& c : cfgs) for (int grid : {1, 148}) {
    const int reps = 500;
    k<<<grid, 128, 200 * 1024>>>(c, reps, d); cudaError_t e = cudaDeviceSynchronize();
    long long cyc; cudaMemcpy(&cyc, d, 8, cudaMemcpyDeviceToHost);
    const double per = (double)cyc / (reps * 8);
    printf("M128 N%-3d %s B %-8s acc %d grid %3d : %6.1f cycles per MMA (K=16) -> %6.0f flop/clk/SM  %s\n", c.N, c.ts ? "A=tmem" : "A=smem", c.b_mn ? "MN-major" : "K-major",
           c.n_acc, grid, per, 2.0 * 128 * c.N * 16 / per, e == cudaSuccess ? "" : cudaGetErrorString(e));
  }
  return 0;
}
