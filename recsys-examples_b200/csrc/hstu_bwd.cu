// HSTU jagged attention backward for sm_100a.  Replaces the reference's HSTUAttentionBackwardSm100
// (third_party/FBGEMM/fbgemm_gpu/experimental/hstu/src/hstu_blackwell/hstu_bwd.py:102-2664, wrapper hstu_ops_gpu.py:257-512).
//
// Math (forward: S = Q K^T, P = mask * silu(alpha S), O = P V / N):
//   dV = P^T dO / N          dP = dO V^T          dS = mask * dP * silu'(alpha S)
//   dQ = (alpha/N) dS K      dK = (alpha/N) dS^T Q
//   with h = alpha S / 2, t = tanh(h):  silu(alpha S) = h + h t,   silu'(alpha S) = 0.5 (1 + t) (1 + h (1 - t)).
//
// B200 design — two tcgen05 kernels from one template, nothing accumulated through global memory:
//   * dKV kernel, KV-stationary: a CTA owns 128 keys of one (b, h) and streams 64-row Q/dO tiles.  Per tile:
//       S^T = K Q^T, dP^T = V dO^T (SS, fp32 in TMEM, double buffered) -> 8 SiLU warps write bf16 P^T and dS^T (thread = key row,
//       so the tiles are K-major A operands as written) -> dV += P^T dO, dK += dS^T Q (B = the streamed dO / Q tiles read MN-major).
//   * dQ kernel, Q-stationary: a CTA owns 128 queries and streams 64-row K/V tiles: S = Q K^T, dP = dO V^T, dS -> dQ += dS K.
//   The reference keeps one KV-stationary kernel and reduce-adds dQ tiles (TMA reduce) into a dense fp32
//   [B, H, max_seqlen, D] workspace that is zero-filled before and converted after every call (hstu_ops_gpu.py:373-382,
//   hstu_bwd.py:2177-2240): 3 extra passes over a padded tensor, non-deterministic.  Recomputing the two score GEMMs costs
//   7 instead of 5 tile-GEMMs per (q,k) tile pair but needs no workspace, no atomics and is bit-reproducible.
#include <cuda_bf16.h>

#include <type_traits>

#include "../../include/hstu_b200.h"
#include "hstu_mask.cuh"
#include "sm100_ptx.cuh"
#include "tma_host.cuh"

using namespace sm100;

namespace hstu_bwd {

struct Params {
  const int32_t* cu_seqlens;
  const int32_t* num_targets;
  const int32_t* num_contexts;
  __nv_bfloat16* out0;   // dKV: dV   dQ: dQ
  __nv_bfloat16* out1;   // dKV: dK
  int H;
  float half_alpha;
  float scale0;          // dKV: 1/N (dV)        dQ: alpha/N
  float scale1;          // dKV: alpha/N (dK)
  int target_group, win_left, win_right;
  volatile int* dbg;     // optional host-mapped progress buffer
};
#define HSTU_DBG(slot, val) do { if (p.dbg && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) { p.dbg[(kIsDQ ? 64 : 32) + (slot)] = (val); __threadfence_system(); } } while (0)

using hstu::SeqMask;
using hstu::Intervals;

template <int D, bool kIsDQ>
struct Smem {
  static constexpr int kX = 128 * D * 2;             // stationary tile bytes (128 rows)
  static constexpr int kY = 64 * D * 2;              // streamed tile bytes (64 rows)
  static constexpr int kPD = 128 * 64 * 2;           // bf16 [128 x 64] operand tile
  static constexpr int oX1 = 0, oX2 = kX;
  // streamed-tile ring depth: a stage is only released when the accumulate GEMMs that read it MN-major have retired, so a
  // 2-deep ring exposes the whole TMA latency every iteration (measured: ~4000 cycles per 128x64 tile); fill the 227 KB.
  static constexpr int kStages = (D == 128) ? (kIsDQ ? 4 : 3) : 4;
  static constexpr int oY = 2 * kX;                  // kStages x (Y1, Y2)
  static constexpr int oDS = oY + 2 * kStages * kY;  // 2 buffers
  static constexpr int oP = oDS + 2 * kPD;           // 2 buffers (dKV only)
  static constexpr int kTotal = oP + (kIsDQ ? 0 : 2 * kPD);
};

// kIsDQ = false: X1 = K, X2 = V (stationary, 128 keys), Y1 = Q, Y2 = dO (streamed, 64 queries)
// kIsDQ = true : X1 = Q, X2 = dO (stationary, 128 queries), Y1 = K, Y2 = V (streamed, 64 keys)
template <int D, bool kIsDQ>
__global__ void __launch_bounds__(384, 1) hstu_bwd_kernel(const __grid_constant__ CUtensorMap map_x1, const __grid_constant__ CUtensorMap map_x2,
                                                          const __grid_constant__ CUtensorMap map_y1, const __grid_constant__ CUtensorMap map_y2,
                                                          Params p) {
  using SM = Smem<D, kIsDQ>;
  constexpr int NH = D / 64;
  const int b = blockIdx.z, h = blockIdx.y;
  const int x_tile = kIsDQ ? (gridDim.x - 1 - blockIdx.x) : blockIdx.x;     // heaviest tiles first
  const int seq_start = p.cu_seqlens[b];
  const int L = p.cu_seqlens[b + 1] - seq_start;
  const int x0 = x_tile * 128;
  if (x0 >= L) return;
  const int x1 = min(L, x0 + 128) - 1;

  SeqMask mk;
  mk.L = L; mk.G = p.target_group; mk.wl = p.win_left; mk.wr = p.win_right;
  mk.has_t = p.num_targets != nullptr; mk.has_c = p.num_contexts != nullptr;
  mk.seqlen_c = mk.has_c ? p.num_contexts[b] : 0;
  mk.seqlen_h = L - (mk.has_t ? p.num_targets[b] : 0);

  // streamed 64-row tiles: regular range [y_lo, y_hi] plus (dKV only) leading context-row tiles that see all history keys
  int y_lo, y_hi, n_ctx = 0;
  if (kIsDQ) {
    int n_end = (mk.wr >= 0) ? min(L, x1 + mk.wr + 1) : L;
    if (mk.has_c && x0 < mk.seqlen_c) n_end = max(n_end, mk.seqlen_h);
    y_lo = (mk.wl >= 0) ? max(0, x0 - mk.wl) / 64 : 0;
    y_hi = (n_end + 63) / 64 - 1;
  } else {
    y_lo = (mk.wr >= 0) ? max(0, x0 - mk.wr) / 64 : 0;
    y_hi = (mk.wl >= 0) ? min(L - 1, x1 + mk.wl) / 64 : (L - 1) / 64;
    if (mk.has_c && x0 < mk.seqlen_h) n_ctx = min(y_lo, (mk.seqlen_c + 63) / 64);
  }
  const int n_iter = n_ctx + (y_hi - y_lo + 1);
  auto y_tile_of = [&](int j) { return j < n_ctx ? j : y_lo + (j - n_ctx); };

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr int NS = SM::kStages;
  __shared__ uint64_t x_full, y_full[NS], y_empty[NS], s_full[2], s_empty[2], pd_full[2], pd_empty[2], acc_full;
  __shared__ uint32_t tmem_base_s;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(&x_full, 1); mbar_init(&acc_full, 1);
    for (int i = 0; i < NS; ++i) { mbar_init(&y_full[i], 1); mbar_init(&y_empty[i], 1); }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_full[i], 1); mbar_init(&s_empty[i], 8);
      mbar_init(&pd_full[i], 8); mbar_init(&pd_empty[i], 1);
    }
    fence_barrier_init();
    tma_prefetch_desc(&map_x1); tma_prefetch_desc(&map_x2); tma_prefetch_desc(&map_y1); tma_prefetch_desc(&map_y2);
  }
  if (warp == 2) tmem_alloc<512>(&tmem_base_s);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  // TMEM columns: S[2] @0,64   dP[2] @128,192   acc0 @256 (dV | dQ)   acc1 @384 (dK)
  const uint32_t tS[2] = {tmem, tmem + 64}, tDP[2] = {tmem + 128, tmem + 192};
  const uint32_t tA0 = tmem + 256, tA1 = tmem + 384;

  if (warp == 0) {
    if (elect_one()) {
      mbar_arrive_expect_tx(&x_full, 2 * SM::kX);
#pragma unroll
      for (int hf = 0; hf < NH; ++hf) {
        tma_load_3d(smem + SM::oX1 + hf * 16384, &map_x1, &x_full, hf * 64, h, seq_start + x0);
        tma_load_3d(smem + SM::oX2 + hf * 16384, &map_x2, &x_full, hf * 64, h, seq_start + x0);
      }
      for (int j = 0; j < n_iter; ++j) {
        const int st = j % NS, ph = (j / NS) & 1;
        const int row = seq_start + y_tile_of(j) * 64;
        mbar_wait(&y_empty[st], ph ^ 1);
        HSTU_DBG(1, j + 1);
        mbar_arrive_expect_tx(&y_full[st], 2 * SM::kY);
        uint8_t* y1 = smem + SM::oY + st * 2 * SM::kY;
#pragma unroll
        for (int hf = 0; hf < NH; ++hf) {
          tma_load_3d(y1 + hf * 8192, &map_y1, &y_full[st], hf * 64, h, row);
          tma_load_3d(y1 + SM::kY + hf * 8192, &map_y2, &y_full[st], hf * 64, h, row);
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      constexpr uint32_t idesc_s = umma_idesc_bf16(128, 64, 0, 0);       // [128 x 64] = X (K-major) * Y^T (K-major)
      constexpr uint32_t idesc_acc = umma_idesc_bf16(128, D, 0, 1);      // [128 x D] += PD (K-major, K = 64) * Y (MN-major)
      const uint32_t aX1 = smem_u32(smem + SM::oX1), aX2 = smem_u32(smem + SM::oX2);
      auto issue_scores = [&](int j) {
        const int st = j & 1, ph = (j >> 1) & 1;           // S / dP TMEM double buffer
        const int ys = j % NS, yph = (j / NS) & 1;         // streamed-tile ring
        mbar_wait(&y_full[ys], yph);
        mbar_wait(&s_empty[st], ph ^ 1);
        tc_fence_after();
        const uint32_t aY1 = smem_u32(smem + SM::oY + ys * 2 * SM::kY), aY2 = aY1 + SM::kY;
#pragma unroll
        for (int k = 0; k < D / 16; ++k) {
          const uint32_t offx = (k >> 2) * 16384 + (k & 3) * 32, offy = (k >> 2) * 8192 + (k & 3) * 32;
          umma_ss(tS[st], umma_desc_sw128(aX1 + offx, 16, 1024), umma_desc_sw128(aY1 + offy, 16, 1024), idesc_s, k > 0);
        }
#pragma unroll
        for (int k = 0; k < D / 16; ++k) {
          const uint32_t offx = (k >> 2) * 16384 + (k & 3) * 32, offy = (k >> 2) * 8192 + (k & 3) * 32;
          umma_ss(tDP[st], umma_desc_sw128(aX2 + offx, 16, 1024), umma_desc_sw128(aY2 + offy, 16, 1024), idesc_s, k > 0);
        }
        umma_commit(&s_full[st]);
      };
      HSTU_DBG(8, n_iter);
      mbar_wait(&x_full, 0);
      HSTU_DBG(9, 1);
      issue_scores(0);
      HSTU_DBG(10, 1);
      for (int j = 0; j < n_iter; ++j) {
        if (j + 1 < n_iter) issue_scores(j + 1);
        const int st = j & 1, ph = (j >> 1) & 1;
        const int ys = j % NS;
        HSTU_DBG(11, j + 1);
        mbar_wait(&pd_full[st], ph);
        HSTU_DBG(12, j + 1);
        tc_fence_after();
        const uint32_t aY1 = smem_u32(smem + SM::oY + ys * 2 * SM::kY), aY2 = aY1 + SM::kY;
        const uint32_t aDS = smem_u32(smem + SM::oDS + st * SM::kPD);
#pragma unroll
        for (int k = 0; k < 4; ++k)      // K = 64 streamed rows: 4 steps of 16 rows (2048 B of the MN-major Y tile)
          umma_ss(kIsDQ ? tA0 : tA1, umma_desc_sw128(aDS + k * 32, 16, 1024), umma_desc_sw128(aY1 + k * 2048, 8192, 1024), idesc_acc, (j > 0 || k > 0));
        if (!kIsDQ) {
          const uint32_t aP = smem_u32(smem + SM::oP + st * SM::kPD);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_ss(tA0, umma_desc_sw128(aP + k * 32, 16, 1024), umma_desc_sw128(aY2 + k * 2048, 8192, 1024), idesc_acc, (j > 0 || k > 0));
        }
        umma_commit(&pd_empty[st]);
        umma_commit(&y_empty[ys]);
      }
      umma_commit(&acc_full);
      HSTU_DBG(13, 1);
    }
  } else if (warp >= 4) {
    const int wq = warp & 3;                        // TMEM lane quadrant
    const int ch = (warp - 4) >> 2;                 // which 32-column half of the 64 streamed columns
    const int rit = wq * 32 + lane;                 // stationary row in tile
    const int xi = x0 + rit;                        // stationary index (dKV: key, dQ: query)
    const uint32_t lane_off = (uint32_t)(wq * 32) << 16;
    const Intervals iv = kIsDQ ? hstu::cols_of_row(mk, xi) : hstu::rows_of_col(mk, xi);
    for (int j = 0; j < n_iter; ++j) {
      const int st = j & 1, ph = (j >> 1) & 1;
      const int y0 = y_tile_of(j) * 64;
      const bool full = kIsDQ ? mk.tile_full(x0, x1, y0, y0 + 63) : mk.tile_full(y0, y0 + 63, x0, x1);
      if (threadIdx.x == 128) HSTU_DBG(16, j + 1);
      mbar_wait(&s_full[st], ph);
      if (threadIdx.x == 128) HSTU_DBG(17, j + 1);
      tc_fence_after();
      mbar_wait(&pd_empty[st], ph ^ 1);            // the accumulate GEMMs of tile j-2 have finished reading these operand buffers
      const uint32_t dds = smem_u32(smem + SM::oDS + st * SM::kPD + rit * 128);
      const uint32_t dpp = smem_u32(smem + SM::oP + st * SM::kPD + rit * 128);
      const f32x2 ha2 = pack2(p.half_alpha, p.half_alpha), one2 = pack2(1.f, 1.f), mone2 = pack2(-1.f, -1.f), mhalf2 = pack2(-0.5f, -0.5f);
      // 16 score columns at a time, mask test hoisted out of the tile (see hstu_fwd.cu: a per-pair `if (!full)` splits the unrolled
      // loop into basic blocks that ptxas cannot schedule the MUFU latency across; 64-register score arrays starve it of registers)
      auto tile = [&](auto masked_tag) {
        constexpr bool kMasked = decltype(masked_tag)::value;
        uint32_t s[16], dp[16];
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          tmem_ld16(tS[st] + lane_off + ch * 32 + c * 16, s);
          tmem_ld16(tDP[st] + lane_off + ch * 32 + c * 16, dp);
          tmem_ld_wait();
          if (c == 1) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&s_empty[st]);
          }
          f32x2 h2[8], t2[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) h2[i] = mul2(pack2(__uint_as_float(s[2 * i]), __uint_as_float(s[2 * i + 1])), ha2);
#pragma unroll
          for (int i = 0; i < 8; ++i) t2[i] = tanh2(h2[i]);
          uint32_t pk_ds[8], pk_p[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            f32x2 pe2 = fma2(h2[i], t2[i], h2[i]);                         // silu = h + h t
            const f32x2 u2 = fma2(t2[i], mone2, one2);                     // u = 1 - t
            // silu' = 0.5 (1 + t) (1 + h (1 - t)) = (1 - u/2)(1 + h u)
            f32x2 de2 = mul2(pack2(__uint_as_float(dp[2 * i]), __uint_as_float(dp[2 * i + 1])), mul2(fma2(u2, mhalf2, one2), fma2(h2[i], u2, one2)));
            if (kMasked) {
              const int yi = y0 + ch * 32 + c * 16 + 2 * i;
              const bool ok0 = iv.has(yi), ok1 = iv.has(yi + 1);
              float a0, a1, b0, b1; unpack2(pe2, a0, a1); unpack2(de2, b0, b1);
              pe2 = pack2(ok0 ? a0 : 0.f, ok1 ? a1 : 0.f);
              de2 = pack2(ok0 ? b0 : 0.f, ok1 ? b1 : 0.f);
            }
            pk_ds[i] = pack_bf16x2_v(de2);
            if (!kIsDQ) pk_p[i] = pack_bf16x2_v(pe2);
          }
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            const uint32_t sw = (uint32_t)((ch * 4 + c * 2 + q) ^ (rit & 7)) << 4;
            sts128(dds + sw, pk_ds[4 * q], pk_ds[4 * q + 1], pk_ds[4 * q + 2], pk_ds[4 * q + 3]);
            if (!kIsDQ) sts128(dpp + sw, pk_p[4 * q], pk_p[4 * q + 1], pk_p[4 * q + 2], pk_p[4 * q + 3]);
          }
        }
      };
      if (full) tile(std::false_type{}); else tile(std::true_type{});
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&pd_full[st]);
      if (threadIdx.x == 128) HSTU_DBG(18, j + 1);
    }
    // epilogue: dKV: warps 4-7 store dV (acc0), warps 8-11 store dK (acc1); dQ: the two warpgroups split the D columns
    mbar_wait(&acc_full, 0);
    if (threadIdx.x == 128) HSTU_DBG(19, 1);
    tc_fence_after();
    {
      // tcgen05.ld is warp-collective (.sync.aligned): every lane must execute it; only the global stores are predicated.
      const uint32_t tacc = kIsDQ ? tA0 : (ch == 0 ? tA0 : tA1);
      __nv_bfloat16* dst = (kIsDQ || ch == 0) ? p.out0 : p.out1;
      const float sc = (kIsDQ || ch == 0) ? p.scale0 : p.scale1;
      __nv_bfloat16* orow = dst + ((int64_t)(seq_start + xi) * p.H + h) * D;
      const int c_begin = kIsDQ ? ch * (D / 2) : 0, c_end = kIsDQ ? (ch + 1) * (D / 2) : D;
      for (int c = c_begin; c < c_end; c += 32) {
        uint32_t o[32];
        tmem_ld32(tacc + lane_off + c, o);
        tmem_ld_wait();
        if (xi < L) {
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4) {
            uint4 v;
            v.x = pack_bf16x2(__uint_as_float(o[8 * q4 + 0]) * sc, __uint_as_float(o[8 * q4 + 1]) * sc);
            v.y = pack_bf16x2(__uint_as_float(o[8 * q4 + 2]) * sc, __uint_as_float(o[8 * q4 + 3]) * sc);
            v.z = pack_bf16x2(__uint_as_float(o[8 * q4 + 4]) * sc, __uint_as_float(o[8 * q4 + 5]) * sc);
            v.w = pack_bf16x2(__uint_as_float(o[8 * q4 + 6]) * sc, __uint_as_float(o[8 * q4 + 7]) * sc);
            *reinterpret_cast<uint4*>(orow + c + q4 * 8) = v;
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc<512>(tmem);
}

template <int D, bool kIsDQ>
int launch(const CUtensorMap& x1, const CUtensorMap& x2, const CUtensorMap& y1, const CUtensorMap& y2, const Params& p, int B, int max_seqlen,
           cudaStream_t stream) {
  constexpr int smem = Smem<D, kIsDQ>::kTotal + 1024;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(hstu_bwd_kernel<D, kIsDQ>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return -(int)e;
    configured = true;
  }
  dim3 grid((max_seqlen + 127) / 128, p.H, B);
  hstu_bwd_kernel<D, kIsDQ><<<grid, 384, smem, stream>>>(x1, x2, y1, y2, p);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : -(int)e;
}

}  // namespace hstu_bwd

extern "C" volatile int* hstu_get_debug_buffer();
extern "C" int hstu_bwd_sm100(const void* dout, const void* q, const void* k, const void* v, void* dq, void* dk, void* dv, const int32_t* cu_seqlens,
                              const int32_t* num_contexts, const int32_t* num_targets, int batch, int heads, int head_dim, int total_tokens,
                              int max_seqlen, int scaling_seqlen, int target_group_size, int window_left, int window_right, float alpha,
                              const int64_t* strides /*q_t,q_h,k_t,k_h,v_t,v_h,do_t,do_h*/, void* stream_) {
  if (batch <= 0 || total_tokens <= 0) return 0;
  if (head_dim != 64 && head_dim != 128) return HSTU_ERR_UNSUPPORTED;
  if (target_group_size < 1 || scaling_seqlen <= 0) return HSTU_ERR_ARG;
  for (int i = 0; i < 8; ++i) if (strides[i] % 8) return HSTU_ERR_ARG;
  cudaStream_t stream = (cudaStream_t)stream_;
  const void* ptr[4] = {q, k, v, dout};
  CUtensorMap big[4], small[4];     // 128-row and 64-row boxes of q, k, v, dO
  for (int i = 0; i < 4; ++i) {
    int rc;
    if ((rc = tma::make_map_3d(&big[i], ptr[i], head_dim, heads, total_tokens, strides[2 * i + 1] * 2, strides[2 * i] * 2, 64, 1, 128))) return rc;
    if ((rc = tma::make_map_3d(&small[i], ptr[i], head_dim, heads, total_tokens, strides[2 * i + 1] * 2, strides[2 * i] * 2, 64, 1, 64))) return rc;
  }
  hstu_bwd::Params p;
  p.cu_seqlens = cu_seqlens; p.num_targets = num_targets; p.num_contexts = num_contexts;
  p.H = heads; p.half_alpha = 0.5f * alpha;
  p.target_group = target_group_size; p.win_left = window_left; p.win_right = window_right;
  const float invN = 1.0f / (float)scaling_seqlen;
  p.dbg = hstu_get_debug_buffer();
  int rc;
  // dK, dV: X = (K, V), Y = (Q, dO)
  p.out0 = reinterpret_cast<__nv_bfloat16*>(dv); p.out1 = reinterpret_cast<__nv_bfloat16*>(dk); p.scale0 = invN; p.scale1 = alpha * invN;
  rc = head_dim == 128 ? hstu_bwd::launch<128, false>(big[1], big[2], small[0], small[3], p, batch, max_seqlen, stream)
                       : hstu_bwd::launch<64, false>(big[1], big[2], small[0], small[3], p, batch, max_seqlen, stream);
  if (rc) return rc;
  // dQ: X = (Q, dO), Y = (K, V)
  p.out0 = reinterpret_cast<__nv_bfloat16*>(dq); p.out1 = nullptr; p.scale0 = alpha * invN; p.scale1 = 0.f;
  rc = head_dim == 128 ? hstu_bwd::launch<128, true>(big[0], big[3], small[1], small[2], p, batch, max_seqlen, stream)
                       : hstu_bwd::launch<64, true>(big[0], big[3], small[1], small[2], p, batch, max_seqlen, stream);
  return rc;
}
