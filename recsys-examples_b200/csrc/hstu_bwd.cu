// HSTU jagged attention backward for sm_100a.  Replaces the reference's HSTUAttentionBackwardSm100
// (third_party/FBGEMM/fbgemm_gpu/experimental/hstu/src/hstu_blackwell/hstu_bwd.py:102-2664, wrapper hstu_ops_gpu.py:257-512).
//
// Math (forward: S = Q K^T, P = mask * silu(alpha S), O = P V / N):
//   dV = P^T dO / N          dP = dO V^T          dS = mask * dP * silu'(alpha S)
//   dQ = (alpha/N) dS K      dK = (alpha/N) dS^T Q
//   with h = alpha S / 2, t = tanh(h):  silu(alpha S) = h + h t,   silu'(alpha S) = 0.5 (1 + t) (1 + h (1 - t)).
//
// B200 design — two tcgen05 kernels from one template, nothing accumulated through global memory.  The stationary operands live in
// TENSOR MEMORY (bf16x2 packed once per CTA) so every score GEMM is a .ts MMA (N = 64: 42 cycles per K=16 step against 75 for the
// shared-memory form, tools/ubench/umma_bench.cu):
//   * dKV kernel, key-stationary: a CTA owns 128 keys of one (b, h), K and V in TMEM, and streams 64-row Q/dO tiles.  Per tile:
//       S^T = K Q^T, dP^T = V dO^T -> 8 SiLU warps pull both into registers (the single S^T / dP^T buffer is handed straight back),
//       write bf16 P^T and dS^T to shared memory (thread = key row, so the tiles are K-major A operands as written) ->
//       dV += P^T dO, dK += dS^T Q (.ss, B = the streamed dO / Q tiles read MN-major).  512 TMEM columns = K, V (128) + dV, dK (256)
//       + S^T, dP^T (128): no room for a second score buffer or for the bf16 operand tiles.
//   * dQ kernel, query-stationary: a CTA owns 128 queries (Q and dO in TMEM) and streams 64-row K/V tiles: S = Q K^T, dP = dO V^T
//       (double-buffered), the SiLU warps write dS back over dP in tensor memory -> dQ += dS K (.ts again; no shared-memory operand
//       tile, no "empty" barriers: the tensor pipe executes one thread's MMAs in issue order).
//   Both kernels are persistent (tile scheduler + tile I/O warpgroup, see the kernel comment).  Measured per 128x64 iteration
//   (tools/hstu_cycles.py, all CTAs): dKV 1940 cycles, dQ 1320.  Splitting dKV into an all-.ts dV kernel and an all-.ts dK kernel was
//   tried: 1070 + 1400 cycles — slower than the fused kernel, so it was dropped.
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
#include "device_info.cuh"

using namespace sm100;

namespace hstu_bwd {

struct Params {
  const int32_t* cu_seqlens;
  const int32_t* num_targets;
  const int32_t* num_contexts;
  const __nv_bfloat16 *x1, *x2;   // stationary operands (dKV: K, V   dQ: Q, dO), element strides below
  int64_t x1_t, x1_h, x2_t, x2_h;
  __nv_bfloat16* out0;   // dKV: dV   dQ: dQ        [T, H, D] contiguous
  __nv_bfloat16* out1;   // dKV: dK
  int H;
  float half_alpha;
  float scale0;          // dKV: 1/N (dV)        dQ: alpha/N
  float scale1;          // dKV: alpha/N (dK)
  int target_group, win_left, win_right;
  int n_x, n_tiles;      // 128-row stationary tiles per sequence (max_seqlen based); total = n_x * H * B
  int* tile_counter;     // zeroed before every launch: the persistent CTAs pull tile indices from it
  volatile int* prof;    // optional cycle-accounting buffer (hstu_set_debug_buffer; DEVICE memory), nullptr in production
};
// cycle accounting (all CTAs add their counters, units of 16 cycles; tools/hstu_cycles.py): clock64() deltas accumulated in registers
#define BWD_T0() long long t__0 = p.prof ? clock64() : 0
#define BWD_ACC(i) do { if (p.prof) { long long t__1 = clock64(); acc__[i] += (int)(t__1 - t__0); t__0 = t__1; } } while (0)
#define BWD_FLUSH(base, n) do { if (p.prof) { for (int i__ = 0; i__ < (n); ++i__) atomicAdd(const_cast<int*>(p.prof) + (kIsDQ ? 96 : 64) + (base) + i__, acc__[i__] >> 4); } } while (0)

using hstu::SeqMask;
using hstu::Intervals;

__device__ __forceinline__ uint4 hstu_bwd_ldg(const void* ptr) {
  uint4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(ptr));
  return v;
}


template <int D, bool kIsDQ>
struct Smem {
  static constexpr int kY = 64 * D * 2;              // streamed tile bytes (64 rows)
  static constexpr int kPD = 128 * 64 * 2;           // bf16 [128 x 64] operand tile (dKV only)
  // streamed-tile ring (Y1, Y2 per stage): a stage is only released when the accumulate GEMMs that read it MN-major have retired
  static constexpr int kStages = kIsDQ ? 6 : 4;
  static constexpr int oY = 0;
  static constexpr int oDS = oY + 2 * kStages * kY;  // 2 buffers (dKV only)
  static constexpr int oP = oDS + 2 * kPD;           // 2 buffers (dKV only)
  static constexpr int kTotal = kIsDQ ? oDS : oP + 2 * kPD;
};

// One unit of work: a 128-row stationary tile of one (sequence, head) and the streamed 64-row tiles it meets.
struct Tile {
  int b, h, seq_start, L, x0, x1, y_lo, n_ctx, n_iter;
  SeqMask mk;
  __device__ __forceinline__ int y_tile_of(int j) const { return j < n_ctx ? j : y_lo + (j - n_ctx); }
};
struct TileMsg { int w, seq_start, L, seqlen_c, num_t; };   // published by the scheduler thread (it alone touches cu_seqlens & co.)

template <bool kIsDQ>
__device__ __forceinline__ bool decode_tile(const Params& p, const TileMsg& m, Tile& t) {
  const int bh = m.w / p.n_x, xr = m.w - bh * p.n_x;
  const int x_tile = kIsDQ ? (p.n_x - 1 - xr) : xr;                        // heaviest tiles first
  t.b = bh / p.H; t.h = bh - t.b * p.H;
  t.seq_start = m.seq_start; t.L = m.L;
  t.x0 = x_tile * 128;
  if (t.x0 >= t.L) return false;
  t.x1 = min(t.L, t.x0 + 128) - 1;
  SeqMask& mk = t.mk;
  mk.L = t.L; mk.G = p.target_group; mk.wl = p.win_left; mk.wr = p.win_right;
  mk.has_t = p.num_targets != nullptr; mk.has_c = p.num_contexts != nullptr;
  mk.seqlen_c = m.seqlen_c;
  mk.seqlen_h = t.L - m.num_t;
  // streamed 64-row tiles: regular range [y_lo, y_hi] plus (dKV only) leading context-row tiles that see all history keys
  int y_hi;
  t.n_ctx = 0;
  if (kIsDQ) {
    int n_end = (mk.wr >= 0) ? min(t.L, t.x1 + mk.wr + 1) : t.L;
    if (mk.has_c && t.x0 < mk.seqlen_c) n_end = max(n_end, mk.seqlen_h);
    t.y_lo = (mk.wl >= 0) ? max(0, t.x0 - mk.wl) / 64 : 0;
    y_hi = (n_end + 63) / 64 - 1;
  } else {
    t.y_lo = (mk.wr >= 0) ? max(0, t.x0 - mk.wr) / 64 : 0;
    y_hi = (mk.wl >= 0) ? min(t.L - 1, t.x1 + mk.wl) / 64 : (t.L - 1) / 64;
    if (mk.has_c && t.x0 < mk.seqlen_h) t.n_ctx = min(t.y_lo, (mk.seqlen_c + 63) / 64);
  }
  t.n_iter = t.n_ctx + (y_hi - t.y_lo + 1);
  return true;
}

// kIsDQ = false: X1 = K, X2 = V (stationary, 128 keys), Y1 = Q, Y2 = dO (streamed, 64 queries)
// kIsDQ = true : X1 = Q, X2 = dO (stationary, 128 queries), Y1 = K, Y2 = V (streamed, 64 keys)
// PERSISTENT: one CTA per SM pulls stationary tiles from a global counter; 16 warps:
//   0 streamed-tile producer (TMA) | 1 MMA issuer | 2 scheduler + TMEM allocator | 3 idle | 4-11 SiLU warps | 12-15 tile I/O warpgroup
//   (packs the next tile's stationary rows into tensor memory as soon as the last score GEMM of the current tile has retired, reads
//   the finished accumulators out and stores them) — the per-tile prologue / epilogue no longer sits on the MMA / SiLU critical path.
template <int D, bool kIsDQ>
__global__ void __launch_bounds__(512, 1) hstu_bwd_kernel(const __grid_constant__ CUtensorMap map_y1, const __grid_constant__ CUtensorMap map_y2, Params p) {
  using SM = Smem<D, kIsDQ>;
  constexpr int NH = D / 64;
  constexpr int NS = SM::kStages;
  constexpr int kRing = 4;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t x_full, x_free, y_full[NS], y_empty[NS], s_full[2], s_empty, pd_full[2], pd_empty[2], acc_full, acc_empty, tile_full[kRing], tile_empty[kRing];
  __shared__ TileMsg tile_ring[kRing];
  __shared__ uint32_t tmem_base_s;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(&x_full, 4); mbar_init(&x_free, 1); mbar_init(&acc_full, 1); mbar_init(&acc_empty, 4); mbar_init(&s_empty, 8);
    for (int i = 0; i < NS; ++i) { mbar_init(&y_full[i], 1); mbar_init(&y_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&s_full[i], 1); mbar_init(&pd_full[i], 8); mbar_init(&pd_empty[i], 1); }
    for (int i = 0; i < kRing; ++i) { mbar_init(&tile_full[i], 1); mbar_init(&tile_empty[i], 14); }
    fence_barrier_init();
    tma_prefetch_desc(&map_y1); tma_prefetch_desc(&map_y2);
  }
  if (warp == 2) tmem_alloc<512>(&tmem_base_s);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  // TMEM columns: X1 @0, X2 @64 (bf16x2 packed stationary operands, D/2 columns each); acc0 @128 (dV | dQ); then
  //   dKV: acc1 @256 (dK), S^T @384, dP^T @448 — single buffers: the SiLU warps pull a tile into registers and hand it straight back
  //   dQ : S[2] @256,320, dP[2] @384,448 — double buffers: dS is written back over dP and stays there until dQ += dS K has read it
  // Buffers and barriers are addressed arithmetically (tS0 + sb * 64, a_s_full + sb * 8): see hstu_fwd.cu.
  const uint32_t tX1 = tmem, tX2 = tmem + 64, tA0 = tmem + 128, tA1 = tmem + 256;
  const uint32_t tS0 = tmem + (kIsDQ ? 256 : 384), tDP0 = tmem + (kIsDQ ? 384 : 448);      // buffer sb at + sb * 64 (dQ only: sb in {0, 1})
  const uint32_t a_x_full = smem_u32(&x_full), a_x_free = smem_u32(&x_free), a_y_full = smem_u32(&y_full[0]), a_y_empty = smem_u32(&y_empty[0]),
                 a_s_full = smem_u32(&s_full[0]), a_s_empty = smem_u32(&s_empty), a_pd_full = smem_u32(&pd_full[0]), a_pd_empty = smem_u32(&pd_empty[0]),
                 a_acc_full = smem_u32(&acc_full), a_acc_empty = smem_u32(&acc_empty), a_tile_full = smem_u32(&tile_full[0]),
                 a_tile_empty = smem_u32(&tile_empty[0]);
  const uint32_t a_smem = smem_u32(smem);

  // every consumer walks the tile ring with its own cursor (warp_wide: all 32 lanes call it together)
  auto next_tile = [&](int& cursor, Tile& t, bool arrive_lane, bool warp_wide) -> bool {
    for (;;) {
      const int slot = cursor % kRing;
      mbar_wait(a_tile_full + slot * 8, (cursor / kRing) & 1);
      const TileMsg m = tile_ring[slot];
      ++cursor;
      if (m.w < 0) return false;                     // end marker: left in place, never released
      if (warp_wide) __syncwarp();
      if (arrive_lane) mbar_arrive(a_tile_empty + slot * 8);
      if (decode_tile<kIsDQ>(p, m, t)) return true;
    }
  };

  if (warp == 2) {
    // ------------------------------------------------------------------ tile scheduler
    if (elect_one()) {
      for (int c = 0;; ++c) {
        const int slot = c % kRing;
        mbar_wait(a_tile_empty + slot * 8, ((c / kRing) & 1) ^ 1);
        TileMsg m;
        m.w = atomicAdd(p.tile_counter, 1);
        if (m.w >= p.n_tiles) m.w = -1;
        if (m.w >= 0) {
          const int b = m.w / (p.n_x * p.H);
          m.seq_start = p.cu_seqlens[b];
          m.L = p.cu_seqlens[b + 1] - m.seq_start;
          m.seqlen_c = p.num_contexts ? p.num_contexts[b] : 0;
          m.num_t = p.num_targets ? p.num_targets[b] : 0;
        }
        const int w = m.w;
        tile_ring[slot] = m;
        mbar_arrive(a_tile_full + slot * 8);
        if (w < 0) break;
      }
    }
  } else if (warp == 0) {
    // ------------------------------------------------------------------ streamed-tile producer
    if (elect_one()) {          // elect.sync, not `lane == 0`: see hstu_fwd.cu
      int cursor = 0, c = 0;
      Tile t;
      while (next_tile(cursor, t, true, false)) {
        for (int j = 0; j < t.n_iter; ++j, ++c) {
          const int st = c % NS, ph = (c / NS) & 1;
          const int row = t.seq_start + t.y_tile_of(j) * 64;
          mbar_wait(a_y_empty + st * 8, ph ^ 1);
          mbar_arrive_expect_tx(a_y_full + st * 8, 2 * SM::kY);
          const uint32_t y1 = a_smem + SM::oY + st * 2 * SM::kY;
#pragma unroll
          for (int hf = 0; hf < NH; ++hf) {
            tma_load_3d(y1 + hf * 8192, &map_y1, a_y_full + st * 8, hf * 64, t.h, row);
            tma_load_3d(y1 + SM::kY + hf * 8192, &map_y2, a_y_full + st * 8, hf * 64, t.h, row);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (elect_one()) {
      int acc__[4] = {0, 0, 0, 0};
      const long long t_begin = p.prof ? clock64() : 0;
      constexpr uint32_t idesc_s = umma_idesc_bf16(128, 64, 0, 0);       // [128 x 64] = X (tensor memory) * Y^T (K-major)
      constexpr uint32_t idesc_acc = umma_idesc_bf16(128, D, 0, 1);      // [128 x D] += PD (K = 64) * Y (MN-major)
      int cursor = 0, yc = 0, it = 0, tc = 0, n_total = 0;               // running: streamed tiles issued for scores, iterations done, tiles done
      Tile t;
      while (next_tile(cursor, t, true, false)) {
        n_total += t.n_iter;
        const int it0 = it;
        auto issue_scores = [&](int j) {
          const int idx = it0 + j;                           // running iteration index
          const int sb = kIsDQ ? (idx & 1) : 0;              // S / dP buffer
          const int ys = yc % NS, yph = (yc / NS) & 1;       // streamed-tile ring
          BWD_T0();
          mbar_wait(a_y_full + ys * 8, yph);
          BWD_ACC(0);
          if (!kIsDQ && idx > 0) mbar_wait(a_s_empty, (idx - 1) & 1);   // SiLU(idx-1) holds S^T / dP^T in registers
          BWD_ACC(1);
          tc_fence_after();
          const uint32_t aY1 = a_smem + SM::oY + ys * 2 * SM::kY, aY2 = aY1 + SM::kY;
#pragma unroll
          for (int k = 0; k < D / 16; ++k)
            umma_ts(tS0 + sb * 64, tX1 + k * 8, umma_desc_sw128(aY1 + (k >> 2) * 8192 + (k & 3) * 32, 16, 1024), idesc_s, k > 0);
#pragma unroll
          for (int k = 0; k < D / 16; ++k)
            umma_ts(tDP0 + sb * 64, tX2 + k * 8, umma_desc_sw128(aY2 + (k >> 2) * 8192 + (k & 3) * 32, 16, 1024), idesc_s, k > 0);
          umma_commit(a_s_full + sb * 8);
          if (j == t.n_iter - 1) umma_commit(a_x_free);      // every MMA that reads this tile's X1 / X2 has been issued: they may be replaced
          ++yc;
        };
        mbar_wait(a_x_full, tc & 1);
        tc_fence_after();
        issue_scores(0);
        int yr = yc - 1;                                     // ring index of the streamed tile of iteration j (released after acc(j))
        for (int j = 0; j < t.n_iter; ++j, ++it, ++yr) {
          // dQ: scores(j+1) go to the other S/dP pair, whose previous tenant dS_{j-1} was consumed by an MMA issued earlier (in-order pipe)
          if (j + 1 < t.n_iter) issue_scores(j + 1);
          const int st = it & 1, ph = (it >> 1) & 1;
          const int ys = yr % NS;
          BWD_T0();
          mbar_wait(a_pd_full + st * 8, ph);
          BWD_ACC(2);
          if (j == 0 && tc > 0) mbar_wait(a_acc_empty, (tc - 1) & 1);    // the I/O warpgroup has pulled the previous tile's accumulators into registers
          tc_fence_after();
          const uint32_t aY1 = a_smem + SM::oY + ys * 2 * SM::kY, aY2 = aY1 + SM::kY;
          if (kIsDQ) {
#pragma unroll
            for (int k = 0; k < 4; ++k)    // A = dS from tensor memory: keys 0-31 packed in dP columns 0-15, keys 32-63 in columns 32-47
              umma_ts(tA0, tDP0 + st * 64 + (k >> 1) * 32 + (k & 1) * 8, umma_desc_sw128(aY1 + k * 2048, 8192, 1024), idesc_acc, (j > 0 || k > 0));
          } else {
            const uint32_t aDS = a_smem + SM::oDS + st * SM::kPD, aP = a_smem + SM::oP + st * SM::kPD;
#pragma unroll
            for (int k = 0; k < 4; ++k)    // K = 64 streamed rows: 4 steps of 16 rows (2048 B of the MN-major Y tile)
              umma_ss(tA1, umma_desc_sw128(aDS + k * 32, 16, 1024), umma_desc_sw128(aY1 + k * 2048, 8192, 1024), idesc_acc, (j > 0 || k > 0));
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_ss(tA0, umma_desc_sw128(aP + k * 32, 16, 1024), umma_desc_sw128(aY2 + k * 2048, 8192, 1024), idesc_acc, (j > 0 || k > 0));
            umma_commit(a_pd_empty + st * 8);
          }
          umma_commit(a_y_empty + ys * 8);
        }
        umma_commit(a_acc_full);
        ++tc;
      }
      if (p.prof) {
        acc__[3] = (int)((clock64() - t_begin) >> 4);
        BWD_FLUSH(0, 3);
        atomicAdd(const_cast<int*>(p.prof) + (kIsDQ ? 96 : 64) + 3, acc__[3]);
        atomicAdd(const_cast<int*>(p.prof) + (kIsDQ ? 96 : 64) + 15, n_total);
        atomicAdd(const_cast<int*>(p.prof) + (kIsDQ ? 96 : 64) + 14, tc);
      }
    }
  } else if (warp >= 12) {
    // ------------------------------------------------------------------ tile I/O warpgroup (one warp per TMEM lane quadrant, thread = row)
    const int wq = warp & 3;
    const int rit = wq * 32 + lane;
    const uint32_t lane_off = (uint32_t)(wq * 32) << 16;
    int cursor = 0, tc = 0;
    Tile t, tn;
    auto x_put = [&](const Tile& tt, bool first) {
      // whole rows of D bf16 straight from global into packed tensor-memory columns.  X1's loads are in flight while the warp waits for
      // the last score GEMM of the current tile (x_free) — only then may the single X1 / X2 buffers be overwritten.
      const int xi = tt.x0 + rit;
      const bool valid = xi < tt.L;
      const __nv_bfloat16* r1 = p.x1 + (int64_t)(tt.seq_start + xi) * p.x1_t + (int64_t)tt.h * p.x1_h;
      const __nv_bfloat16* r2 = p.x2 + (int64_t)(tt.seq_start + xi) * p.x2_t + (int64_t)tt.h * p.x2_h;
      uint4 xr[D / 8];
#pragma unroll
      for (int c = 0; c < D / 8; ++c) xr[c] = valid ? hstu_bwd_ldg(r1 + c * 8) : make_uint4(0, 0, 0, 0);
      if (!first) mbar_wait(a_x_free, (tc & 1));           // tile tc's score GEMMs have retired
      tc_fence_after();
#pragma unroll
      for (int c = 0; c < D / 16; ++c) {
        const uint32_t r[8] = {xr[2 * c].x, xr[2 * c].y, xr[2 * c].z, xr[2 * c].w, xr[2 * c + 1].x, xr[2 * c + 1].y, xr[2 * c + 1].z, xr[2 * c + 1].w};
        tmem_st8(tX1 + lane_off + c * 8, r);
      }
#pragma unroll
      for (int c = 0; c < D / 8; ++c) xr[c] = valid ? hstu_bwd_ldg(r2 + c * 8) : make_uint4(0, 0, 0, 0);
#pragma unroll
      for (int c = 0; c < D / 16; ++c) {
        const uint32_t r[8] = {xr[2 * c].x, xr[2 * c].y, xr[2 * c].z, xr[2 * c].w, xr[2 * c + 1].x, xr[2 * c + 1].y, xr[2 * c + 1].z, xr[2 * c + 1].w};
        tmem_st8(tX2 + lane_off + c * 8, r);
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(a_x_full);
    };
    bool have = next_tile(cursor, t, lane == 0, true);
    if (have) x_put(t, true);
    while (have) {
      const bool have_next = next_tile(cursor, tn, lane == 0, true);
      if (have_next) x_put(tn, false);
      const int xi = t.x0 + rit;
      mbar_wait(a_acc_full, tc & 1);
      tc_fence_after();
      // read-out: dKV: dV (acc0) then dK (acc1); dQ: acc0.  64 columns at a time; after the last load the accumulators go back (acc_empty).
      constexpr int kRounds = (kIsDQ ? 1 : 2) * (D / 64);
#pragma unroll
      for (int rd = 0; rd < kRounds; ++rd) {
        const int which = rd / (D / 64), c0 = (rd % (D / 64)) * 64;
        uint32_t o[64];
        tmem_ld32((which ? tA1 : tA0) + lane_off + c0, *reinterpret_cast<uint32_t(*)[32]>(&o[0]));
        tmem_ld32((which ? tA1 : tA0) + lane_off + c0 + 32, *reinterpret_cast<uint32_t(*)[32]>(&o[32]));
        tmem_ld_wait();
        if (rd == kRounds - 1) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(a_acc_empty);
        }
        if (xi < t.L) {
          const float sc = which ? p.scale1 : p.scale0;
          __nv_bfloat16* orow = (which ? p.out1 : p.out0) + ((int64_t)(t.seq_start + xi) * p.H + t.h) * D + c0;
#pragma unroll
          for (int q8 = 0; q8 < 8; ++q8) {
            uint4 v;
            v.x = pack_bf16x2(__uint_as_float(o[8 * q8 + 0]) * sc, __uint_as_float(o[8 * q8 + 1]) * sc);
            v.y = pack_bf16x2(__uint_as_float(o[8 * q8 + 2]) * sc, __uint_as_float(o[8 * q8 + 3]) * sc);
            v.z = pack_bf16x2(__uint_as_float(o[8 * q8 + 4]) * sc, __uint_as_float(o[8 * q8 + 5]) * sc);
            v.w = pack_bf16x2(__uint_as_float(o[8 * q8 + 6]) * sc, __uint_as_float(o[8 * q8 + 7]) * sc);
            *reinterpret_cast<uint4*>(orow + q8 * 8) = v;
          }
        }
      }
      ++tc;
      have = have_next;
      if (have) t = tn;
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ SiLU warps
    const int wq = warp & 3;                        // TMEM lane quadrant
    const int ch = (warp - 4) >> 2;                 // which 32-column half of the 64 streamed columns
    const int rit = wq * 32 + lane;                 // stationary row in tile
    const uint32_t lane_off = (uint32_t)(wq * 32) << 16;
    int acc__[4] = {0, 0, 0, 0};
    const f32x2 ha2 = pack2(p.half_alpha, p.half_alpha), one2 = pack2(1.f, 1.f), mone2 = pack2(-1.f, -1.f), mhalf2 = pack2(-0.5f, -0.5f);
    int cursor = 0, it = 0;
    Tile t;
    while (next_tile(cursor, t, lane == 0, true)) {
      const int xi = t.x0 + rit;                    // stationary index (dKV: key, dQ: query)
      const Intervals iv = kIsDQ ? hstu::cols_of_row(t.mk, xi) : hstu::rows_of_col(t.mk, xi);
      for (int j = 0; j < t.n_iter; ++j, ++it) {
        const int st = it & 1, ph = (it >> 1) & 1;
        const int sb = kIsDQ ? st : 0, sph = kIsDQ ? ph : (it & 1);
        const int y0 = t.y_tile_of(j) * 64;
        const bool full = kIsDQ ? t.mk.tile_full(t.x0, t.x1, y0, y0 + 63) : t.mk.tile_full(y0, y0 + 63, t.x0, t.x1);
        BWD_T0();
        if (!kIsDQ) mbar_wait(a_pd_empty + st * 8, ph ^ 1);            // the accumulate GEMMs of iteration it-2 have finished reading these operand buffers
        BWD_ACC(0);
        mbar_wait(a_s_full + sb * 8, sph);
        BWD_ACC(1);
        tc_fence_after();
        const uint32_t dds = a_smem + SM::oDS + st * SM::kPD + rit * 128;
        const uint32_t dpp = a_smem + SM::oP + st * SM::kPD + rit * 128;
        // Mask test hoisted out of the tile (see hstu_fwd.cu).  16 score columns at a time keeps registers free for the scheduler.
        auto tile = [&](auto masked_tag) {
          constexpr bool kMasked = decltype(masked_tag)::value;
          uint32_t s[16], dp[16];
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            tmem_ld16(tS0 + sb * 64 + lane_off + ch * 32 + c * 16, s);
            tmem_ld16(tDP0 + sb * 64 + lane_off + ch * 32 + c * 16, dp);
            tmem_ld_wait();
            if (!kIsDQ && c == 1) {                     // S^T / dP^T are in registers: the MMA warp may overwrite them with the next tile
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(a_s_empty);
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
            if (kIsDQ) {
              tmem_st8(tDP0 + sb * 64 + lane_off + ch * 32 + c * 8, pk_ds);       // over dP columns this warpgroup has already read
            } else {
#pragma unroll
              for (int q = 0; q < 2; ++q) {
                const uint32_t sw = (uint32_t)((ch * 4 + c * 2 + q) ^ (rit & 7)) << 4;
                sts128(dds + sw, pk_ds[4 * q], pk_ds[4 * q + 1], pk_ds[4 * q + 2], pk_ds[4 * q + 3]);
                sts128(dpp + sw, pk_p[4 * q], pk_p[4 * q + 1], pk_p[4 * q + 2], pk_p[4 * q + 3]);
              }
            }
          }
        };
        if (full) tile(std::false_type{}); else tile(std::true_type{});
        BWD_ACC(2);
        if (kIsDQ) { tmem_st_wait(); tc_fence_before(); }
        else fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(a_pd_full + st * 8);
        BWD_ACC(3);
      }
    }
    if (threadIdx.x == 128) BWD_FLUSH(8, 4);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc<512>(tmem);
}

template <int D, bool kIsDQ>
int launch(const CUtensorMap& y1, const CUtensorMap& y2, const Params& p_in, int B, int max_seqlen, int* tile_counter, cudaStream_t stream) {
  constexpr int smem = Smem<D, kIsDQ>::kTotal + 1024;
  static std::atomic<int> configured[devinfo::kMaxDevices];          // the attribute is per device
  cudaError_t e = devinfo::once_per_device(configured, [] { return cudaFuncSetAttribute(hstu_bwd_kernel<D, kIsDQ>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); });
  if (e != cudaSuccess) return -(int)e;
  Params p = p_in;
  p.n_x = (max_seqlen + 127) / 128;
  p.n_tiles = p.n_x * p.H * B;
  p.tile_counter = tile_counter;                                      // caller workspace (one counter per launch)
  e = cudaMemsetAsync(p.tile_counter, 0, sizeof(int), stream);
  if (e != cudaSuccess) return -(int)e;
  const int sms = devinfo::sm_count();
  const int grid = p.n_tiles < sms ? p.n_tiles : sms;
  hstu_bwd_kernel<D, kIsDQ><<<grid, 512, smem, stream>>>(y1, y2, p);
  e = cudaGetLastError();
  return e == cudaSuccess ? 0 : -(int)e;
}

}  // namespace hstu_bwd

extern "C" volatile int* hstu_get_debug_buffer();
extern "C" int64_t hstu_workspace_bytes();
extern "C" int hstu_bwd_sm100(const void* dout, const void* q, const void* k, const void* v, void* dq, void* dk, void* dv, const int32_t* cu_seqlens,
                              const int32_t* num_contexts, const int32_t* num_targets, int batch, int heads, int head_dim, int total_tokens,
                              int max_seqlen, int scaling_seqlen, int target_group_size, int window_left, int window_right, float alpha,
                              const int64_t* strides /*q_t,q_h,k_t,k_h,v_t,v_h,do_t,do_h*/, void* workspace, int64_t workspace_bytes,
                              void* stream_) {
  if (!workspace || workspace_bytes < hstu_workspace_bytes()) return HSTU_ERR_WORKSPACE;
  int* counters = (int*)workspace;
  if (batch <= 0 || total_tokens <= 0) return 0;
  if (head_dim != 64 && head_dim != 128) return HSTU_ERR_UNSUPPORTED;
  if (target_group_size < 1 || scaling_seqlen <= 0) return HSTU_ERR_ARG;
  for (int i = 0; i < 8; ++i) if (strides[i] % 8) return HSTU_ERR_ARG;
  cudaStream_t stream = (cudaStream_t)stream_;
  const void* ptr[4] = {q, k, v, dout};
  tma::bind_context(q);
  CUtensorMap small[4];     // 64-row boxes of q, k, v, dO (the streamed operands; the stationary ones are read row-wise into tensor memory)
  for (int i = 0; i < 4; ++i) {
    if (reinterpret_cast<uintptr_t>(ptr[i]) & 15) return HSTU_ERR_ARG;
    int rc;
    if ((rc = tma::make_map_3d(&small[i], ptr[i], head_dim, heads, total_tokens, strides[2 * i + 1] * 2, strides[2 * i] * 2, 64, 1, 64))) return rc;
  }
  auto bf = [](const void* x) { return static_cast<const __nv_bfloat16*>(x); };
  hstu_bwd::Params p;
  p.cu_seqlens = cu_seqlens; p.num_targets = num_targets; p.num_contexts = num_contexts;
  p.H = heads; p.half_alpha = 0.5f * alpha;
  p.target_group = target_group_size; p.win_left = window_left; p.win_right = window_right;
  const float invN = 1.0f / (float)scaling_seqlen;
  p.prof = hstu_get_debug_buffer();
  int rc;
  using hstu_bwd::launch;
  // dK, dV: X = (K, V), Y = (Q, dO)
  p.out0 = reinterpret_cast<__nv_bfloat16*>(dv); p.out1 = reinterpret_cast<__nv_bfloat16*>(dk); p.scale0 = invN; p.scale1 = alpha * invN;
  p.x1 = bf(k); p.x1_t = strides[2]; p.x1_h = strides[3]; p.x2 = bf(v); p.x2_t = strides[4]; p.x2_h = strides[5];
  rc = head_dim == 128 ? launch<128, false>(small[0], small[3], p, batch, max_seqlen, counters, stream) : launch<64, false>(small[0], small[3], p, batch, max_seqlen, counters, stream);
  if (rc) return rc;
  // dQ: X = (Q, dO), Y = (K, V)
  p.out0 = reinterpret_cast<__nv_bfloat16*>(dq); p.out1 = nullptr; p.scale0 = alpha * invN; p.scale1 = 0.f;
  p.x1 = bf(q); p.x1_t = strides[0]; p.x1_h = strides[1]; p.x2 = bf(dout); p.x2_t = strides[6]; p.x2_h = strides[7];
  rc = head_dim == 128 ? launch<128, true>(small[1], small[2], p, batch, max_seqlen, counters + 8, stream) : launch<64, true>(small[1], small[2], p, batch, max_seqlen, counters + 8, stream);
  return rc;
}
