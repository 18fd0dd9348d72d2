// HSTU jagged attention forward for sm_100a:  O_i = (1/N) * sum_{j in mask(i)} silu(alpha * q_i.k_j) * v_j
// per (sequence b, head h), N = scaling_seqlen.  Replaces the reference's CuTe-DSL kernel
// HSTUAttentionForwardSm100 (third_party/FBGEMM/fbgemm_gpu/experimental/hstu/src/hstu_blackwell/hstu_fwd.py:33-2022,
// host wrapper hstu_ops_gpu.py:85-252); mask rule = hstu_blackwell/mask.py:61-127 (causal / local window,
// target groups, contexts).  Hand-written tcgen05 / TMA / TMEM:
//
//   CTA = one 128-row Q tile of one (b, h); 12 warps:
//     warp 0    TMA producer  : Q once, then K_j / V_j tiles (128 x D bf16, SWIZZLE_128B boxes of 64 columns) into 2-stage rings
//     warp 1    MMA issuer    : S_j = Q K_j^T  (SS, K-major x K-major, fp32 in TMEM, double-buffered S0/S1 so QK^T(j+1) overlaps
//                               the SiLU of tile j);  O += P_j V_j  (SS, P K-major from smem, V MN-major straight from its TMA tile)
//     warp 2    TMEM alloc/dealloc (512 columns: S0 @0, S1 @128, O @256)
//     warps 4-11 two SiLU warpgroups (64 score columns each): thread = accumulator row; tcgen05.ld 32 columns at a time -> h + h*tanh.approx(h), h = alpha/2*s
//                               (1 FMUL + 1 MUFU + 1 FFMA per score), mask only on boundary tiles, bf16 pack, st.shared into the
//                               swizzled K-major P tile; finally O: TMEM -> regs -> *1/N -> bf16 -> 16-byte global stores.
//   The 1/N scale is applied once to O (linear), not to every P element.
#include <cuda_bf16.h>

#include "../../include/hstu_b200.h"
#include "hstu_mask.cuh"
#include "sm100_ptx.cuh"
#include "tma_host.cuh"

using namespace sm100;

namespace hstu {

struct FwdParams {
  const int32_t* cu_seqlens;
  const int32_t* num_targets;    // nullable
  const int32_t* num_contexts;   // nullable
  __nv_bfloat16* out;            // [T, H, D] contiguous
  int H;
  float half_alpha;              // alpha / 2
  float inv_scale;               // 1 / scaling_seqlen
  int target_group;
  int win_left, win_right;       // -1 = unbounded
  volatile int* dbg;             // optional host-mapped progress buffer (hstu_set_debug_buffer), nullptr in production
};

// progress marks readable from the host even if the kernel never finishes (development aid)
// cycle accounting for CTA (0,0,0) when a debug buffer is installed: HSTU_T0 / HSTU_ACC(slot) accumulate clock64() deltas
#define HSTU_T0() long long t__0 = (p.dbg ? clock64() : 0)
#define HSTU_ACC(slot) do { if (p.dbg && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) { long long t__1 = clock64(); p.dbg[slot] += (int)(t__1 - t__0); t__0 = t__1; } } while (0)
#define HSTU_DBG(slot, val) do { if (p.dbg && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) { p.dbg[slot] = (val); __threadfence_system(); } } while (0)

template <int D>
struct FwdSmem {
  static constexpr int kTile = 128 * D * 2;         // bytes of one 128 x D bf16 tile
  static constexpr int kQ = 0;
  static constexpr int kK = kQ + kTile;             // 2 stages
  static constexpr int kV = kK + 2 * kTile;         // 2 stages
  static constexpr int kP = kV + 2 * kTile;         // 128 x 128 bf16
  static constexpr int kTotal = kP + 32768;
};

template <int D>
__global__ void __launch_bounds__(384, 1) hstu_fwd_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k,
                                                          const __grid_constant__ CUtensorMap map_v, FwdParams p) {
  using SM = FwdSmem<D>;
  constexpr int NH = D / 64;                         // 64-column (128-byte) halves per tile row
  const int b = blockIdx.z, h = blockIdx.y;
  const int m_tile = gridDim.x - 1 - blockIdx.x;     // heaviest (longest causal row) tiles first
  const int seq_start = p.cu_seqlens[b];
  const int L = p.cu_seqlens[b + 1] - seq_start;
  const int r0 = m_tile * 128;
  if (r0 >= L) return;
  const int r1 = min(L, r0 + 128) - 1;

  SeqMask mk;
  mk.L = L; mk.G = p.target_group; mk.wl = p.win_left; mk.wr = p.win_right;
  mk.has_t = p.num_targets != nullptr; mk.has_c = p.num_contexts != nullptr;
  mk.seqlen_c = mk.has_c ? p.num_contexts[b] : 0;
  mk.seqlen_h = L - (mk.has_t ? p.num_targets[b] : 0);
  int n_end = (mk.wr >= 0) ? min(L, r1 + mk.wr + 1) : L;
  if (mk.has_c && r0 < mk.seqlen_c) n_end = max(n_end, mk.seqlen_h);
  const int nb0 = (mk.wl >= 0) ? max(0, r0 - mk.wl) / 128 : 0;
  const int nb1 = (n_end + 127) / 128;
  const int n_iter = nb1 - nb0;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t q_full, k_full[2], k_empty[2], v_full[2], v_empty[2], s_full[2], s_empty[2], p_full, p_empty, o_full;
  __shared__ uint32_t tmem_base_s;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(&q_full, 1); mbar_init(&p_full, 8); mbar_init(&p_empty, 1); mbar_init(&o_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&k_full[i], 1); mbar_init(&k_empty[i], 1); mbar_init(&v_full[i], 1); mbar_init(&v_empty[i], 1);
      mbar_init(&s_full[i], 1); mbar_init(&s_empty[i], 8);
    }
    fence_barrier_init();
    tma_prefetch_desc(&map_q); tma_prefetch_desc(&map_k); tma_prefetch_desc(&map_v);
  }
  if (warp == 2) tmem_alloc<512>(&tmem_base_s);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  const uint32_t tS[2] = {tmem, tmem + 128};
  const uint32_t tO = tmem + 256;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      HSTU_DBG(0, 1);
      mbar_arrive_expect_tx(&q_full, SM::kTile);
#pragma unroll
      for (int hf = 0; hf < NH; ++hf) tma_load_3d(smem + SM::kQ + hf * 16384, &map_q, &q_full, hf * 64, h, seq_start + r0);
      for (int j = 0; j < n_iter; ++j) {
        const int st = j & 1, ph = (j >> 1) & 1;
        const int row = seq_start + (nb0 + j) * 128;
        mbar_wait(&k_empty[st], ph ^ 1);
        HSTU_DBG(1, j + 1);
        mbar_arrive_expect_tx(&k_full[st], SM::kTile);
#pragma unroll
        for (int hf = 0; hf < NH; ++hf) tma_load_3d(smem + SM::kK + st * SM::kTile + hf * 16384, &map_k, &k_full[st], hf * 64, h, row);
        mbar_wait(&v_empty[st], ph ^ 1);
        HSTU_DBG(2, j + 1);
        mbar_arrive_expect_tx(&v_full[st], SM::kTile);
#pragma unroll
        for (int hf = 0; hf < NH; ++hf) tma_load_3d(smem + SM::kV + st * SM::kTile + hf * 16384, &map_v, &v_full[st], hf * 64, h, row);
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (one thread)
    if (lane == 0) {
      constexpr uint32_t idesc_qk = umma_idesc_bf16(128, 128, 0, 0);
      constexpr uint32_t idesc_pv = umma_idesc_bf16(128, D, 0, 1);
      const uint32_t aQ = smem_u32(smem + SM::kQ), aP = smem_u32(smem + SM::kP);
      auto issue_qk = [&](int j) {
        const int st = j & 1, ph = (j >> 1) & 1;
        HSTU_T0();
        mbar_wait(&k_full[st], ph);
        HSTU_ACC(40);
        mbar_wait(&s_empty[st], ph ^ 1);
        HSTU_ACC(41);
        tc_fence_after();
        const uint32_t aK = smem_u32(smem + SM::kK + st * SM::kTile);
#pragma unroll
        for (int k = 0; k < D / 16; ++k) {
          const uint32_t off = (k >> 2) * 16384 + (k & 3) * 32;
          umma_ss(tS[st], umma_desc_sw128(aQ + off, 16, 1024), umma_desc_sw128(aK + off, 16, 1024), idesc_qk, k > 0);
        }
        umma_commit(&s_full[st]);
        umma_commit(&k_empty[st]);
      };
      HSTU_DBG(8, n_iter);
      mbar_wait(&q_full, 0);
      HSTU_DBG(9, 1);
      issue_qk(0);
      HSTU_DBG(10, 1);
      for (int j = 0; j < n_iter; ++j) {
        if (j + 1 < n_iter) issue_qk(j + 1);
        const int st = j & 1, ph = (j >> 1) & 1;
        HSTU_T0();
        mbar_wait(&v_full[st], ph);
        HSTU_ACC(42);
        mbar_wait(&p_full, j & 1);
        HSTU_ACC(43);
        tc_fence_after();
        const uint32_t aV = smem_u32(smem + SM::kV + st * SM::kTile);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const uint32_t offp = (k >> 2) * 16384 + (k & 3) * 32;
          umma_ss(tO, umma_desc_sw128(aP + offp, 16, 1024), umma_desc_sw128(aV + k * 2048, 16384, 1024), idesc_pv, (j > 0 || k > 0));
        }
        umma_commit(&v_empty[st]);
        umma_commit(&p_empty);
      }
      umma_commit(&o_full);
      HSTU_DBG(13, 1);
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ SiLU warpgroup + epilogue
    // two SiLU warpgroups split the 128 score columns of every tile (warps 4-7: columns 0-63, warps 8-11: columns 64-127) so
    // two warps per SM sub-partition hide each other's TMEM-load / MUFU latency
    const int wq = warp & 3;                       // TMEM lane quadrant
    const int ch = (warp - 4) >> 2;                // column half
    const int rit = wq * 32 + lane;                // row in tile
    const int row = r0 + rit;
    const uint32_t lane_off = (uint32_t)(wq * 32) << 16;
    uint8_t* sP = smem + SM::kP + ch * 16384 + rit * 128;
    const Intervals iv = cols_of_row(mk, row);
    for (int j = 0; j < n_iter; ++j) {
      const int st = j & 1, ph = (j >> 1) & 1;
      const int c_base = (nb0 + j) * 128 + ch * 64;
      const bool full = mk.tile_full(r0, r1, c_base, c_base + 63);
      HSTU_T0();
      mbar_wait(&s_full[st], ph);
      if (threadIdx.x == 128) HSTU_ACC(48);
      tc_fence_after();
      uint32_t s0[32], s1[32];
      tmem_ld32(tS[st] + lane_off + ch * 64, s0);
      tmem_ld32(tS[st] + lane_off + ch * 64 + 32, s1);
      tmem_ld_wait();
      tc_fence_before();                           // S_j fully read: hand the buffer back to the MMA warp
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_empty[st]);
      if (threadIdx.x == 128) HSTU_ACC(49);
      uint32_t pk[32];
      const f32x2 ha2 = pack2(p.half_alpha, p.half_alpha);
#pragma unroll
      for (int i = 0; i < 64; i += 2) {
        const uint32_t a = i < 32 ? s0[i] : s1[i - 32], b = i < 32 ? s0[i + 1] : s1[i - 31];
        const f32x2 h2 = mul2(pack2(__uint_as_float(a), __uint_as_float(b)), ha2);     // h = alpha/2 * s
        f32x2 p2 = fma2(h2, tanh2(h2), h2);                                          // silu(alpha s) = h + h tanh(h)
        if (!full) {
          const int col = c_base + i;
          float p0, p1; unpack2(p2, p0, p1);
          p2 = pack2(iv.has(col) ? p0 : 0.f, iv.has(col + 1) ? p1 : 0.f);
        }
        pk[i >> 1] = pack_bf16x2_v(p2);
      }
      if (threadIdx.x == 128) HSTU_ACC(50);
      mbar_wait(&p_empty, (j & 1) ^ 1);            // PV(j-1) has consumed the previous P tile
      if (threadIdx.x == 128) HSTU_ACC(51);
#pragma unroll
      for (int q8 = 0; q8 < 8; ++q8)
        *reinterpret_cast<uint4*>(sP + ((q8 ^ (rit & 7)) << 4)) = make_uint4(pk[4 * q8], pk[4 * q8 + 1], pk[4 * q8 + 2], pk[4 * q8 + 3]);
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full);
      if (threadIdx.x == 128) HSTU_ACC(52);
    }
    // epilogue: each warpgroup stores half of the D output columns of its rows
    mbar_wait(&o_full, 0);
    if (threadIdx.x == 128) HSTU_DBG(19, 1);
    tc_fence_after();
    __nv_bfloat16* orow = p.out + ((int64_t)(seq_start + row) * p.H + h) * D;
#pragma unroll
    for (int cc = 0; cc < D / 64; ++cc) {
      const int c = ch * (D / 2) + cc * 32;
      uint32_t o[32];
      tmem_ld32(tO + lane_off + c, o);
      tmem_ld_wait();
      if (row < L) {
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) {
          uint4 v;
          v.x = pack_bf16x2(__uint_as_float(o[8 * q4 + 0]) * p.inv_scale, __uint_as_float(o[8 * q4 + 1]) * p.inv_scale);
          v.y = pack_bf16x2(__uint_as_float(o[8 * q4 + 2]) * p.inv_scale, __uint_as_float(o[8 * q4 + 3]) * p.inv_scale);
          v.z = pack_bf16x2(__uint_as_float(o[8 * q4 + 4]) * p.inv_scale, __uint_as_float(o[8 * q4 + 5]) * p.inv_scale);
          v.w = pack_bf16x2(__uint_as_float(o[8 * q4 + 6]) * p.inv_scale, __uint_as_float(o[8 * q4 + 7]) * p.inv_scale);
          *reinterpret_cast<uint4*>(orow + c + q4 * 8) = v;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc<512>(tmem);
}

template <int D>
int launch_fwd(const CUtensorMap& mq, const CUtensorMap& mkk, const CUtensorMap& mv, const FwdParams& p, int B, int max_seqlen, cudaStream_t stream) {
  constexpr int smem = FwdSmem<D>::kTotal + 1024;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(hstu_fwd_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return -(int)e;
    configured = true;
  }
  dim3 grid((max_seqlen + 127) / 128, p.H, B);
  hstu_fwd_kernel<D><<<grid, 384, smem, stream>>>(mq, mkk, mv, p);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : -(int)e;
}

}  // namespace hstu

static volatile int* g_hstu_dbg = nullptr;
extern "C" int hstu_set_debug_buffer(int* host_mapped) { g_hstu_dbg = host_mapped; return 0; }
extern "C" volatile int* hstu_get_debug_buffer() { return g_hstu_dbg; }

extern "C" int hstu_fwd_sm100(const void* q, const void* k, const void* v, void* out, const int32_t* cu_seqlens, const int32_t* num_contexts,
                              const int32_t* num_targets, int batch, int heads, int head_dim, int total_tokens, int max_seqlen, int scaling_seqlen,
                              int target_group_size, int window_left, int window_right, float alpha, const int64_t* strides /*q_t,q_h,k_t,k_h,v_t,v_h (elements)*/,
                              void* stream) {
  if (batch <= 0 || total_tokens <= 0) return 0;
  if (head_dim != 64 && head_dim != 128) return HSTU_ERR_UNSUPPORTED;
  if (target_group_size < 1 || scaling_seqlen <= 0) return HSTU_ERR_ARG;
  for (int i = 0; i < 6; ++i) if (strides[i] % 8) return HSTU_ERR_ARG;                       // TMA: 16-byte aligned strides
  if ((reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(k) | reinterpret_cast<uintptr_t>(v) | reinterpret_cast<uintptr_t>(out)) & 15) return HSTU_ERR_ARG;
  CUtensorMap mq, mk, mv;
  int rc;
  if ((rc = tma::make_map_3d(&mq, q, head_dim, heads, total_tokens, strides[1] * 2, strides[0] * 2, 64, 1, 128))) return rc;
  if ((rc = tma::make_map_3d(&mk, k, head_dim, heads, total_tokens, strides[3] * 2, strides[2] * 2, 64, 1, 128))) return rc;
  if ((rc = tma::make_map_3d(&mv, v, head_dim, heads, total_tokens, strides[5] * 2, strides[4] * 2, 64, 1, 128))) return rc;
  hstu::FwdParams p;
  p.cu_seqlens = cu_seqlens; p.num_targets = num_targets; p.num_contexts = num_contexts;
  p.out = reinterpret_cast<__nv_bfloat16*>(out);
  p.H = heads; p.half_alpha = 0.5f * alpha; p.inv_scale = 1.0f / (float)scaling_seqlen;
  p.target_group = target_group_size; p.win_left = window_left; p.win_right = window_right;
  p.dbg = g_hstu_dbg;
  if (head_dim == 128) return hstu::launch_fwd<128>(mq, mk, mv, p, batch, max_seqlen, (cudaStream_t)stream);
  return hstu::launch_fwd<64>(mq, mk, mv, p, batch, max_seqlen, (cudaStream_t)stream);
}
