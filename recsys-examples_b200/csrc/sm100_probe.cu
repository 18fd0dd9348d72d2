// Descriptor probe: one-CTA 128x128x128 bf16 GEMM through TMA + tcgen05 with every operand layout the HSTU
// kernels rely on.  Development/test aid (tests/test_sm100_probe_gpu.py): it pins, on real hardware, the
// shared-memory descriptor conventions (K-major / MN-major SWIZZLE_128B, LBO/SBO, per-k-step advance),
// the instruction descriptor bits and the TMEM accumulator layout that sm100_ptx.cuh encodes by hand.
//   variant 0: C = A  * B^T   A[m][k] K-major,  B[n][k] K-major        (Q K^T)
//   variant 1: C = A  * Bt    A[m][k] K-major,  Bt[k][n] MN-major B    (P V)
//   variant 2: C = At^T * B^T At[k][m] MN-major A, B[n][k] K-major     (dS^T as A of dQ)
//   variant 3: C = At^T * Bt  both MN-major                            (dQ = dS K)
//   variant 4: C = A  * B^T   A from TENSOR MEMORY (.ts MMA, bf16x2 packed by tcgen05.st), B K-major
//   variant 5: C = A  * Bt    A from tensor memory, Bt[k][n] MN-major  (P V with P kept in TMEM)
#include "../../include/hstu_b200.h"
#include "sm100_ptx.cuh"
#include "tma_host.cuh"

using namespace sm100;

namespace {

struct ProbeParams {
  uint32_t a_lbo, a_sbo, a_kstep, a_khalf;   // descriptor bytes; kstep = start advance per UMMA_K; khalf = advance after 4 k-steps (K-major)
  uint32_t b_lbo, b_sbo, b_kstep, b_khalf;
  int a_mn, b_mn;
  int a_tmem;                                // variants 4/5: A operand staged in tensor memory
  const uint32_t* a_global;                  // A as packed bf16 pairs [128][64]
};

__global__ void __launch_bounds__(128) probe_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                                                    float* __restrict__ C, ProbeParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;            // 32 KB: two [128 rows][128 B] swizzled halves
  uint8_t* sB = smem + 32768;
  __shared__ uint64_t bar_load, bar_mma;
  __shared__ uint32_t tmem_base;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { mbar_init(&bar_load, 1); mbar_init(&bar_mma, 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc<256>(&tmem_base);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base;
  if (p.a_tmem) {                            // thread = row m: 64 packed columns at TMEM columns 128..191
    const uint32_t lane_off = (uint32_t)(warp * 32) << 16;
    for (int c = 0; c < 64; c += 8) {
      uint32_t r[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) r[i] = p.a_global[threadIdx.x * 64 + c + i];
      tmem_st8(tmem + lane_off + 128 + c, r);
    }
    tmem_st_wait();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
  }
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(&bar_load, 65536);
    tma_load_3d(sA, &map_a, &bar_load, 0, 0, 0);
    tma_load_3d(sA + 16384, &map_a, &bar_load, 64, 0, 0);
    tma_load_3d(sB, &map_b, &bar_load, 0, 0, 0);
    tma_load_3d(sB + 16384, &map_b, &bar_load, 64, 0, 0);
    mbar_wait(&bar_load, 0);
    tc_fence_after();
    const uint32_t idesc = umma_idesc_bf16(128, 128, p.a_mn, p.b_mn);
    for (int k = 0; k < 8; ++k) {
      uint32_t a_off = (k & 3) * p.a_kstep + (k >> 2) * p.a_khalf;
      uint32_t b_off = (k & 3) * p.b_kstep + (k >> 2) * p.b_khalf;
      uint64_t da = umma_desc_sw128(smem_u32(sA) + a_off, p.a_lbo, p.a_sbo);
      uint64_t db = umma_desc_sw128(smem_u32(sB) + b_off, p.b_lbo, p.b_sbo);
      if (p.a_tmem) umma_ts(tmem, tmem + 128 + k * 8, db, idesc, k > 0);   // 16 k values = 8 packed columns per step
      else umma_ss(tmem, da, db, idesc, k > 0);
    }
    umma_commit(&bar_mma);
  }
  __syncwarp();
  mbar_wait(&bar_mma, 0);
  tc_fence_after();
  const int row = threadIdx.x;   // warp w owns TMEM lanes 32w..32w+31
  for (int c = 0; c < 128; c += 32) {
    uint32_t r[32];
    tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + c, r);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 32; ++i) C[row * 128 + c + i] = __uint_as_float(r[i]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<256>(tmem);
}

}  // namespace

extern "C" int sm100_probe_gemm(const void* A, const void* B, float* C, int variant, const uint32_t* overrides /*8 host u32, nullable*/, void* stream) {
  tma::bind_context(A);
  CUtensorMap ma, mb;
  // both operands are [128][128] bf16 row-major tensors; a 3D map (inner 128, 1, 128 rows), box (64, 1, 128), SWIZZLE_128B
  int rc = tma::make_map_3d(&ma, A, 128, 1, 128, 128 * 2, 128 * 2, 64, 1, 128);
  if (rc) return rc;
  rc = tma::make_map_3d(&mb, B, 128, 1, 128, 128 * 2, 128 * 2, 64, 1, 128);
  if (rc) return rc;
  ProbeParams p;
  // K-major: rows of 128 B, 8-row groups 1024 B apart (SBO), LBO unused (1); +32 B per k-step, next 64-col half at +16384
  // MN-major: LBO = next 64-element MN chunk (+16384), SBO = next 8-row k group (+1024); +2048 B per k-step (16 k rows)
  const uint32_t kmaj[4] = {16, 1024, 32, 16384}, mnmaj[4] = {16384, 1024, 2048, 8192};
  p.a_mn = (variant == 2 || variant == 3);
  p.b_mn = (variant == 1 || variant == 3 || variant == 5);
  p.a_tmem = (variant == 4 || variant == 5);
  p.a_global = static_cast<const uint32_t*>(A);
  const uint32_t* a = p.a_mn ? mnmaj : kmaj;
  const uint32_t* b = p.b_mn ? mnmaj : kmaj;
  p.a_lbo = a[0]; p.a_sbo = a[1]; p.a_kstep = a[2]; p.a_khalf = a[3];
  p.b_lbo = b[0]; p.b_sbo = b[1]; p.b_kstep = b[2]; p.b_khalf = b[3];
  if (overrides) {
    p.a_lbo = overrides[0]; p.a_sbo = overrides[1]; p.a_kstep = overrides[2]; p.a_khalf = overrides[3];
    p.b_lbo = overrides[4]; p.b_sbo = overrides[5]; p.b_kstep = overrides[6]; p.b_khalf = overrides[7];
  }
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 66 * 1024);
  probe_kernel<<<1, 128, 66 * 1024, (cudaStream_t)stream>>>(ma, mb, C, p);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : -(int)e;
}
