// HSTU jagged attention forward for sm_100a:  O_i = (1/N) * sum_{j in mask(i)} silu(alpha * q_i.k_j) * v_j
// per (sequence b, head h), N = scaling_seqlen.  Replaces the reference's CuTe-DSL kernel
// HSTUAttentionForwardSm100 (third_party/FBGEMM/fbgemm_gpu/experimental/hstu/src/hstu_blackwell/hstu_fwd.py:33-2022,
// host wrapper hstu_ops_gpu.py:85-252); mask rule = hstu_blackwell/mask.py:61-127 (causal / local window,
// target groups, contexts).  Hand-written tcgen05 / TMA / TMEM:
//
//   PERSISTENT kernel, one 512-thread CTA per SM; work unit = one 128-row Q tile of one (b, h), pulled from a global counter; 16 warps:
//     warp 0    K producer    : K_j tiles (128 x D bf16, SWIZZLE_128B boxes of 64 columns) into a 3-stage ring
//     warp 3    V producer    : V_j tiles, own ring, own thread (a busy V slot must never delay the K tile the next QK^T needs)
//     warp 1    MMA issuer    : S_j = Q K_j^T and O += P_j V_j, BOTH with the A operand in tensor memory (.ts):
//                               Q is packed into TMEM once per tile, P_j is written back over S_j by the SiLU warps.  Measured on
//                               B200 (tools/ubench/umma_bench.cu): 128x128x16 .ss = 107 cycles (shared-memory operand bandwidth,
//                               ~75 B/clk), .ts = 74 cycles.  S is double-buffered so QK^T(j+1) overlaps the SiLU of tile j.
//                               The issuing thread is chosen with elect.sync: under `if (lane == 0)` ptxas wraps every UTCHMMA /
//                               UTMALDG in an ELECT + BRA.U.ANY loop (94 cycles per MMA issue).
//     warp 2    tile scheduler (atomicAdd on the tile counter, cu_seqlens loads, 4-deep ring of tile messages) + TMEM alloc/dealloc
//               (512 columns: S0 @0, S1 @128, O @256, Q0 @384, Q1 @448)
//     warps 4-11 two SiLU warpgroups (64 score columns each): thread = accumulator row; tcgen05.ld 16 columns at a time ->
//                               h + h*tanh.approx(h), h = alpha/2*s (packed FMUL2 / FFMA2 around one MUFU.TANH per score: the SFU
//                               is the floor, 1024 cycles per 128x128 tile) -> bf16x2 -> tcgen05.st into the first half of the
//                               warpgroup's own S columns (already read).  Nothing else: the loop is the critical path of the kernel.
//     warps 12-15 tile I/O warpgroup: next tile's Q rows -> the other TMEM Q buffer; finished O: TMEM -> regs -> *1/N -> bf16 -> global.
//   No shared-memory traffic for Q or P at all: shared memory only holds the K / V rings.
//   The 1/N scale is applied once to O (linear), not to every P element.
#include <cuda_bf16.h>

#include <type_traits>

#include "../../include/hstu_b200.h"
#include "hstu_mask.cuh"
#include "sm100_ptx.cuh"
#include "tma_host.cuh"
#include "device_info.cuh"

using namespace sm100;

namespace hstu {

struct FwdParams {
  const int32_t* cu_seqlens;
  const int32_t* num_targets;    // nullable
  const int32_t* num_contexts;   // nullable
  const __nv_bfloat16* q;        // [T, H, D] with element strides q_t / q_h (rows are read straight into tensor memory)
  int64_t q_t, q_h;
  __nv_bfloat16* out;            // [T, H, D] contiguous
  int H;
  float half_alpha;              // alpha / 2
  float inv_scale;               // 1 / scaling_seqlen
  int target_group;
  int win_left, win_right;       // -1 = unbounded
  int n_m, n_tiles;              // 128-query tiles per sequence (max_seqlen based), total tiles = n_m * H * B
  int* tile_counter;             // zeroed before every launch: the persistent CTAs pull tile indices from it
  volatile int* dbg;             // optional DEVICE buffer (hstu_set_debug_buffer), nullptr in production
};

// Cycle accounting (kProf instantiation only, selected when a debug buffer — DEVICE memory, zeroed by the caller — is installed):
// clock64() deltas are accumulated in REGISTERS and added to the buffer once per role per CTA, in units of 16 cycles, summed over ALL
// CTAs (a store + fence per mark costs microseconds and distorts exactly the pipeline it is meant to observe).
// Slots: 8 = iterations, 9 = CTAs, 10 = CTA lifetime, 11 = prologue (launch -> first S tile ready), 12 = epilogue (last P -> exit);
// 40.. MMA thread waits (k_full, -, v_full, p_full), 48.. SiLU thread 128 (s_full wait, -, math, -, arrive).
#define HSTU_T0() long long t__0 = kProf ? clock64() : 0
#define HSTU_ACC(i) do { if (kProf) { long long t__1 = clock64(); acc__[i] += (int)(t__1 - t__0); t__0 = t__1; } } while (0)
#define HSTU_FLUSH(base, n) do { if (kProf && p.dbg) { for (int i__ = 0; i__ < (n); ++i__) atomicAdd(const_cast<int*>(p.dbg) + (base) + i__, acc__[i__] >> 4); } } while (0)

template <int D>
struct FwdSmem {
  static constexpr int kTile = 128 * D * 2;         // bytes of one 128 x D bf16 tile
  static constexpr int kStages = 3;
  static constexpr int kK = 0;
  static constexpr int kV = kK + kStages * kTile;
  static constexpr int kTotal = kV + kStages * kTile;
};

__device__ __forceinline__ uint4 ldg_nc_u4(const void* ptr) {
  uint4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(ptr));
  return v;
}

// One unit of work: a 128-query tile of one (sequence, head).
struct FwdTile {
  int b, h, seq_start, L, r0, r1, nb0, n_iter;
  SeqMask mk;
};
// What the scheduler thread publishes per tile: the global loads (cu_seqlens, contexts, targets) are done ONCE, by the otherwise idle
// scheduler, instead of by each of the 11 consumers on its critical path.
struct TileMsg { int w, seq_start, L, seqlen_c, num_t; };
// tile index w (heavy-first inside each (b, h), the 32-ish tiles of one (b, h) adjacent so concurrently running CTAs share its K / V
// through L2) -> geometry.  false: the tile lies beyond the end of its (jagged) sequence.
__device__ __forceinline__ bool decode_tile(const FwdParams& p, const TileMsg& m, FwdTile& t) {
  const int w = m.w;
  const int bh = w / p.n_m, m_tile = p.n_m - 1 - (w - bh * p.n_m);
  t.b = bh / p.H; t.h = bh - t.b * p.H;
  t.seq_start = m.seq_start;
  t.L = m.L;
  t.r0 = m_tile * 128;
  if (t.r0 >= t.L) return false;
  t.r1 = min(t.L, t.r0 + 128) - 1;
  SeqMask& mk = t.mk;
  mk.L = t.L; mk.G = p.target_group; mk.wl = p.win_left; mk.wr = p.win_right;
  mk.has_t = p.num_targets != nullptr; mk.has_c = p.num_contexts != nullptr;
  mk.seqlen_c = m.seqlen_c;
  mk.seqlen_h = t.L - m.num_t;
  int n_end = (mk.wr >= 0) ? min(t.L, t.r1 + mk.wr + 1) : t.L;
  if (mk.has_c && t.r0 < mk.seqlen_c) n_end = max(n_end, mk.seqlen_h);
  t.nb0 = (mk.wl >= 0) ? max(0, t.r0 - mk.wl) / 128 : 0;
  t.n_iter = (n_end + 127) / 128 - t.nb0;
  return true;
}

// PERSISTENT kernel: one CTA per SM pulls tiles from a global counter (same order the hardware block scheduler used to give the
// one-tile-per-CTA version: measured 6400 cycles of prologue + 2600 of epilogue on a 33000-cycle CTA, 27 % of the SM's time with the
// tensor pipe idle).  Across tile boundaries the K / V producers keep prefetching, the MMA thread issues Q K^T of the next tile as
// soon as its Q is in tensor memory, and only the O read-out of the SiLU warps is exposed.
template <int D, bool kProf>
__global__ void __launch_bounds__(512, 1) hstu_fwd_kernel(const __grid_constant__ CUtensorMap map_k, const __grid_constant__ CUtensorMap map_v, FwdParams p) {
  using SM = FwdSmem<D>;
  constexpr int NH = D / 64;                         // 64-column (128-byte) halves per tile row
  constexpr int S = SM::kStages;
  constexpr int kRing = 4;                           // tile ring between the scheduler thread and the 15 consumers (K, V, MMA, 8 SiLU warps, 4 I/O warps)
  const long long t_cta0 = kProf ? clock64() : 0;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t q_full[2], k_full[S], k_empty[S], v_full[S], v_empty[S], s_full[2], p_full[2], o_full, o_empty, tile_full[kRing], tile_empty[kRing];
  __shared__ TileMsg tile_ring[kRing];
  __shared__ uint32_t tmem_base_s;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(&q_full[0], 4); mbar_init(&q_full[1], 4); mbar_init(&o_full, 1); mbar_init(&o_empty, 4);
    for (int i = 0; i < S; ++i) { mbar_init(&k_full[i], 1); mbar_init(&k_empty[i], 1); mbar_init(&v_full[i], 1); mbar_init(&v_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&s_full[i], 1); mbar_init(&p_full[i], 8); }
    for (int i = 0; i < kRing; ++i) { mbar_init(&tile_full[i], 1); mbar_init(&tile_empty[i], 15); }
    fence_barrier_init();
    tma_prefetch_desc(&map_k); tma_prefetch_desc(&map_v);
  }
  if (warp == 2) tmem_alloc<512>(&tmem_base_s);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  // TMEM columns: S0 @0, S1 @128, O @256, Q0 @384, Q1 @448 (bf16x2 packed Q tiles, D / 2 columns, double-buffered across tiles).
  // Buffers and barriers are addressed arithmetically (tS0 + st * 128, a_s_full + st * 8): a runtime-indexed local array is a
  // local-memory load, and smem_u32(&bar[st]) an S2R + address chain, in front of every wait of the inner loops.
  const uint32_t tS0 = tmem, tO = tmem + 256, tQ0 = tmem + 384;
  const uint32_t a_q_full = smem_u32(&q_full[0]), a_k_full = smem_u32(&k_full[0]), a_k_empty = smem_u32(&k_empty[0]), a_v_full = smem_u32(&v_full[0]),
                 a_v_empty = smem_u32(&v_empty[0]), a_s_full = smem_u32(&s_full[0]), a_p_full = smem_u32(&p_full[0]), a_o_full = smem_u32(&o_full),
                 a_o_empty = smem_u32(&o_empty), a_tile_full = smem_u32(&tile_full[0]), a_tile_empty = smem_u32(&tile_empty[0]);
  const uint32_t a_smem = smem_u32(smem);

  // every consumer walks the ring with its own cursor; returns the next tile that exists (skipping tiles past the end of their
  // sequence) or false when the scheduler has published the end marker
  // (warp_wide: all 32 lanes call it together — the SiLU warps; the single-thread roles must not execute a warp barrier)
  auto next_tile = [&](int& cursor, FwdTile& t, bool arrive_lane, bool warp_wide) -> bool {
    for (;;) {
      const int slot = cursor % kRing;
      mbar_wait(a_tile_full + slot * 8, (cursor / kRing) & 1);
      const TileMsg m = tile_ring[slot];
      ++cursor;
      if (m.w < 0) return false;                     // end marker: left in place, never released
      if (warp_wide) __syncwarp();                   // every lane has read the slot before lane 0 hands it back
      if (arrive_lane) mbar_arrive(a_tile_empty + slot * 8);
      if (decode_tile(p, m, t)) return true;
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
          const int b = m.w / (p.n_m * p.H);
          m.seq_start = p.cu_seqlens[b];
          m.L = p.cu_seqlens[b + 1] - m.seq_start;
          m.seqlen_c = p.num_contexts ? p.num_contexts[b] : 0;
          m.num_t = p.num_targets ? p.num_targets[b] : 0;
        }
        const int w = m.w;
        tile_ring[slot] = m;
        mbar_arrive(a_tile_full + slot * 8);         // release semantics: the store above is visible to the waiters
        if (w < 0) break;
      }
    }
  } else if (warp == 0 || warp == 3) {
    // ------------------------------------------------------------------ K (warp 0) / V (warp 3) producers
    if (elect_one()) {
      const CUtensorMap* map = warp == 0 ? &map_k : &map_v;
      const uint32_t ring = a_smem + (warp == 0 ? SM::kK : SM::kV);
      const uint32_t full = warp == 0 ? a_k_full : a_v_full;
      const uint32_t empty = warp == 0 ? a_k_empty : a_v_empty;
      int cursor = 0, c = 0;                         // c: running K / V tile count (ring slot and phase)
      FwdTile t;
      while (next_tile(cursor, t, true, false)) {
        for (int j = 0; j < t.n_iter; ++j, ++c) {
          const int st = c % S, ph = (c / S) & 1;
          const int row = t.seq_start + (t.nb0 + j) * 128;
          mbar_wait(empty + st * 8, ph ^ 1);
          mbar_arrive_expect_tx(full + st * 8, SM::kTile);
#pragma unroll
          for (int hf = 0; hf < NH; ++hf) tma_load_3d(ring + st * SM::kTile + hf * 16384, map, full + st * 8, hf * 64, t.h, row);
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (one elected thread)
    if (elect_one()) {
      int acc__[4] = {0, 0, 0, 0};
      constexpr uint32_t idesc_qk = umma_idesc_bf16(128, 128, 0, 0);
      constexpr uint32_t idesc_pv = umma_idesc_bf16(128, D, 0, 1);
      int cursor = 0, kc = 0, vc = 0, it = 0, tc = 0;   // running counts: K tiles issued, V tiles consumed, P tiles consumed, tiles
      FwdTile t;
      int n_total = 0;
      while (next_tile(cursor, t, true, false)) {
        n_total += t.n_iter;
        auto issue_qk = [&](int sidx) {                  // sidx = running index of the score tile (S double buffer)
          const int st = sidx & 1;
          const int ks = kc % S, kph = (kc / S) & 1;       // K ring
          HSTU_T0();
          mbar_wait(a_k_full + ks * 8, kph);
          HSTU_ACC(0);
          tc_fence_after();
          const uint32_t aK = a_smem + SM::kK + ks * SM::kTile;
#pragma unroll
          for (int k = 0; k < D / 16; ++k)                   // A = Q from tensor memory: 16 k values = 8 packed columns per step
            umma_ts(tS0 + st * 128, tQ0 + (tc & 1) * 64 + k * 8, umma_desc_sw128(aK + (k >> 2) * 16384 + (k & 3) * 32, 16, 1024), idesc_qk, k > 0);
          umma_commit(a_s_full + st * 8);
          umma_commit(a_k_empty + ks * 8);
          ++kc;
        };
        {
          HSTU_T0();
          mbar_wait(a_q_full + (tc & 1) * 8, (tc >> 1) & 1);       // Q of this tile is in tensor memory (written by the I/O warpgroup one tile ahead)
          HSTU_ACC(1);
        }
        tc_fence_after();
        issue_qk(it);
        for (int j = 0; j < t.n_iter; ++j, ++it, ++vc) {
          // S_{j+1} goes to the other S buffer; its previous tenant P_{j-1} was consumed by P_{j-1} V_{j-1}, issued before this MMA
          // (the tensor pipe executes one thread's MMAs in issue order), so no "S empty" barrier is needed
          if (j + 1 < t.n_iter) issue_qk(it + 1);
          const int st = it & 1, ph = (it >> 1) & 1;
          const int vs = vc % S, vph = (vc / S) & 1;
          HSTU_T0();
          mbar_wait(a_v_full + vs * 8, vph);
          HSTU_ACC(2);
          mbar_wait(a_p_full + st * 8, ph);
          HSTU_ACC(3);
          if (j == 0 && tc > 0) mbar_wait(a_o_empty, (tc - 1) & 1);   // the I/O warpgroup has pulled the previous tile's O into registers
          tc_fence_after();
          const uint32_t aV = a_smem + SM::kV + vs * SM::kTile;
#pragma unroll
          for (int k = 0; k < 8; ++k)                        // A = P_j: keys 0-63 packed in S columns 0-31, keys 64-127 in columns 64-95
            umma_ts(tO, tS0 + st * 128 + (k >> 2) * 64 + (k & 3) * 8, umma_desc_sw128(aV + k * 2048, 16384, 1024), idesc_pv, (j > 0 || k > 0));
          umma_commit(a_v_empty + vs * 8);
        }
        umma_commit(a_o_full);
        ++tc;
      }
      if (kProf && p.dbg) { atomicAdd(const_cast<int*>(p.dbg) + 8, n_total); atomicAdd(const_cast<int*>(p.dbg) + 9, tc); }
      HSTU_FLUSH(40, 4);
    }
  } else if (warp >= 12) {
    // ------------------------------------------------------------------ tile I/O warpgroup (warps 12-15, one per TMEM lane quadrant)
    // Keeps everything that happens once per tile OFF the SiLU warps' critical path (measured there: O read-out 3400 cycles, Q hand-over
    // 1500, per 16.5 iterations of 1250): packs the NEXT tile's Q rows into the other Q buffer, then reads the finished O accumulator,
    // hands it back (o_empty) and stores it.  thread = row.
    const int wq = warp & 3;
    const int rit = wq * 32 + lane;
    const uint32_t lane_off = (uint32_t)(wq * 32) << 16;
    int acc__[4] = {0, 0, 0, 0};
    int cursor = 0, tc = 0;
    FwdTile t, tn;
    auto q_put = [&](const FwdTile& tt, int buf) {   // whole row of D bf16 -> D/2 packed columns, then signal the MMA thread
      const int row = tt.r0 + rit;
      const __nv_bfloat16* qrow = p.q + (int64_t)(tt.seq_start + row) * p.q_t + (int64_t)tt.h * p.q_h;
      uint4 qreg[D / 8];
#pragma unroll
      for (int c = 0; c < D / 8; ++c) qreg[c] = row < tt.L ? ldg_nc_u4(qrow + c * 8) : make_uint4(0, 0, 0, 0);
#pragma unroll
      for (int c = 0; c < D / 16; ++c) {
        const uint32_t r[8] = {qreg[2 * c].x, qreg[2 * c].y, qreg[2 * c].z, qreg[2 * c].w, qreg[2 * c + 1].x, qreg[2 * c + 1].y, qreg[2 * c + 1].z, qreg[2 * c + 1].w};
        tmem_st8(tQ0 + buf * 64 + lane_off + c * 8, r);
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(a_q_full + buf * 8);
    };
    bool have = next_tile(cursor, t, lane == 0, true);
    if (have) q_put(t, 0);
    while (have) {
      const bool have_next = next_tile(cursor, tn, lane == 0, true);
      // buffer (tc+1)&1 last held the Q of tile tc-1, whose MMAs all completed before its o_full, which this warp waited for below
      if (have_next) q_put(tn, (tc + 1) & 1);
      const int row = t.r0 + rit;
      HSTU_T0();
      mbar_wait(a_o_full, tc & 1);
      HSTU_ACC(0);
      tc_fence_after();
      __nv_bfloat16* orow = p.out + ((int64_t)(t.seq_start + row) * p.H + t.h) * D;
#pragma unroll
      for (int half = 0; half < D / 64; ++half) {
        uint32_t o[64];
        tmem_ld32(tO + lane_off + half * 64, *reinterpret_cast<uint32_t(*)[32]>(&o[0]));
        tmem_ld32(tO + lane_off + half * 64 + 32, *reinterpret_cast<uint32_t(*)[32]>(&o[32]));
        tmem_ld_wait();
        if (half == D / 64 - 1) {                    // O is in registers: the next tile's first P V may overwrite it
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(a_o_empty);
        }
        if (row < t.L) {
#pragma unroll
          for (int q8 = 0; q8 < 8; ++q8) {
            uint4 v;
            v.x = pack_bf16x2(__uint_as_float(o[8 * q8 + 0]) * p.inv_scale, __uint_as_float(o[8 * q8 + 1]) * p.inv_scale);
            v.y = pack_bf16x2(__uint_as_float(o[8 * q8 + 2]) * p.inv_scale, __uint_as_float(o[8 * q8 + 3]) * p.inv_scale);
            v.z = pack_bf16x2(__uint_as_float(o[8 * q8 + 4]) * p.inv_scale, __uint_as_float(o[8 * q8 + 5]) * p.inv_scale);
            v.w = pack_bf16x2(__uint_as_float(o[8 * q8 + 6]) * p.inv_scale, __uint_as_float(o[8 * q8 + 7]) * p.inv_scale);
            *reinterpret_cast<uint4*>(orow + half * 64 + q8 * 8) = v;
          }
        }
      }
      HSTU_ACC(1);
      ++tc;
      have = have_next;
      if (have) t = tn;
    }
    if (threadIdx.x == 384) HSTU_FLUSH(56, 2);
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ SiLU warpgroups + epilogue
    // two SiLU warpgroups split the 128 score columns of every tile (warps 4-7: columns 0-63, warps 8-11: columns 64-127) so
    // two warps per SM sub-partition hide each other's TMEM-load / MUFU latency
    const int wq = warp & 3;                       // TMEM lane quadrant
    const int ch = (warp - 4) >> 2;                // column half
    const int rit = wq * 32 + lane;                // row in tile
    const uint32_t lane_off = (uint32_t)(wq * 32) << 16;
    const f32x2 ha2 = pack2(p.half_alpha, p.half_alpha);
    int acc__[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    long long t_mark = kProf ? clock64() : 0;       // untracked-time accounting: 5 = between iterations, 6 = Q hand-over, 7 = tile fetch + setup
#define HSTU_GAP(i) do { if (kProf) { long long t__g = clock64(); acc__[i] += (int)(t__g - t_mark); t_mark = t__g; } } while (0)
#define HSTU_MARK() do { if (kProf) t_mark = clock64(); } while (0)
    int cursor = 0, it = 0, tc = 0;
    FwdTile t;
    HSTU_MARK();
    while (next_tile(cursor, t, lane == 0, true)) {
      const int row = t.r0 + rit;
      const Intervals iv = cols_of_row(t.mk, row);
      int c_base = t.nb0 * 128 + ch * 64;
      bool full = t.mk.tile_full(t.r0, t.r1, c_base, c_base + 63);
      for (int j = 0; j < t.n_iter; ++j, ++it) {
        const int st = it & 1, ph = (it >> 1) & 1;
        const uint32_t t_s = tS0 + st * 128 + lane_off + ch * 64;
        HSTU_GAP(j == 0 ? 7 : 5);
        HSTU_T0();
        mbar_wait(a_s_full + st * 8, ph);
        HSTU_ACC(0);
        tc_fence_after();
        // 16 score columns at a time (tcgen05.ld x16, the next chunk in flight while this one is in the SFU); packed chunk c (8 columns)
        // overwrites S columns 8c..8c+7 of this warpgroup's half, which chunks <= c have already read.  The mask test is hoisted out of the
        // tile: a per-pair `if (!full)` split the unrolled loop into 32 basic blocks and ptxas could not cover the MUFU latency.
        auto tile = [&](auto masked_tag) {
          constexpr bool kMasked = decltype(masked_tag)::value;
          uint32_t sa[16], sb[16];
          tmem_ld16(t_s, sa);
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            uint32_t (&cur)[16] = (c & 1) ? sb : sa;
            uint32_t (&nxt)[16] = (c & 1) ? sa : sb;
            tmem_ld_wait();
            if (c < 3) tmem_ld16(t_s + 16 * (c + 1), nxt);
            f32x2 h2[8], t2[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) h2[i] = mul2(pack2(__uint_as_float(cur[2 * i]), __uint_as_float(cur[2 * i + 1])), ha2);   // h = alpha/2 s
#pragma unroll
            for (int i = 0; i < 8; ++i) t2[i] = tanh2(h2[i]);
            uint32_t pk[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              f32x2 p2 = fma2(h2[i], t2[i], h2[i]);                                      // silu(alpha s) = h + h tanh(h)
              if (kMasked) {
                const int col = c_base + c * 16 + 2 * i;
                float p0, p1; unpack2(p2, p0, p1);
                p2 = pack2(iv.has(col) ? p0 : 0.f, iv.has(col + 1) ? p1 : 0.f);
              }
              pk[i] = pack_bf16x2_v(p2);
            }
            tmem_st8(t_s + c * 8, pk);
          }
        };
        if (full) tile(std::false_type{}); else tile(std::true_type{});
        // next iteration's scalars, computed here (every PTX wrapper below is a compiler barrier: after the arrive they would sit on the
        // critical path between two tiles)
        c_base += 128;
        full = t.mk.tile_full(t.r0, t.r1, c_base, c_base + 63);
        HSTU_ACC(2);
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(a_p_full + st * 8);
        HSTU_ACC(4);
        HSTU_MARK();
      }
    }
    if (threadIdx.x == 128) HSTU_FLUSH(48, 9);
  }
  if (kProf && threadIdx.x == 128 && p.dbg) atomicAdd(const_cast<int*>(p.dbg) + 10, (int)((clock64() - t_cta0) >> 4));
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc<512>(tmem);
}

template <int D>
int launch_fwd(const CUtensorMap& mkk, const CUtensorMap& mv, const FwdParams& p, int B, int max_seqlen, int* tile_counter, cudaStream_t stream) {
  constexpr int smem = FwdSmem<D>::kTotal + 1024;
  static std::atomic<int> configured[devinfo::kMaxDevices];          // the attribute is per device
  cudaError_t e = devinfo::once_per_device(configured, [] {
    cudaError_t e1 = cudaFuncSetAttribute(hstu_fwd_kernel<D, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    return e1 != cudaSuccess ? e1 : cudaFuncSetAttribute(hstu_fwd_kernel<D, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  });
  if (e != cudaSuccess) return -(int)e;
  FwdParams q = p;
  q.n_m = (max_seqlen + 127) / 128;
  q.n_tiles = q.n_m * p.H * B;
  q.tile_counter = tile_counter;                                      // caller workspace: one counter per launch, nothing shared between streams / graphs
  e = cudaMemsetAsync(q.tile_counter, 0, sizeof(int), stream);
  if (e != cudaSuccess) return -(int)e;
  const int sms = devinfo::sm_count();
  const int grid = q.n_tiles < sms ? q.n_tiles : sms;                // one persistent CTA per SM
  if (p.dbg) hstu_fwd_kernel<D, true><<<grid, 512, smem, stream>>>(mkk, mv, q);
  else hstu_fwd_kernel<D, false><<<grid, 512, smem, stream>>>(mkk, mv, q);
  e = cudaGetLastError();
  return e == cudaSuccess ? 0 : -(int)e;
}

}  // namespace hstu

static volatile int* g_hstu_dbg = nullptr;
extern "C" int hstu_set_debug_buffer(int* host_mapped) { g_hstu_dbg = host_mapped; return 0; }
extern "C" volatile int* hstu_get_debug_buffer() { return g_hstu_dbg; }

extern "C" int64_t hstu_workspace_bytes() { return 64; }
extern "C" int hstu_fwd_sm100(const void* q, const void* k, const void* v, void* out, const int32_t* cu_seqlens, const int32_t* num_contexts,
                              const int32_t* num_targets, int batch, int heads, int head_dim, int total_tokens, int max_seqlen, int scaling_seqlen,
                              int target_group_size, int window_left, int window_right, float alpha, const int64_t* strides /*q_t,q_h,k_t,k_h,v_t,v_h (elements)*/,
                              void* workspace, int64_t workspace_bytes, void* stream) {
  if (!workspace || workspace_bytes < hstu_workspace_bytes()) return HSTU_ERR_WORKSPACE;
  if (batch <= 0 || total_tokens <= 0) return 0;
  if (head_dim != 64 && head_dim != 128) return HSTU_ERR_UNSUPPORTED;
  if (target_group_size < 1 || scaling_seqlen <= 0) return HSTU_ERR_ARG;
  for (int i = 0; i < 6; ++i) if (strides[i] % 8) return HSTU_ERR_ARG;                       // TMA: 16-byte aligned strides
  if ((reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(k) | reinterpret_cast<uintptr_t>(v) | reinterpret_cast<uintptr_t>(out)) & 15) return HSTU_ERR_ARG;
  tma::bind_context(q);
  CUtensorMap mk, mv;
  int rc;
  if ((rc = tma::make_map_3d(&mk, k, head_dim, heads, total_tokens, strides[3] * 2, strides[2] * 2, 64, 1, 128))) return rc;
  if ((rc = tma::make_map_3d(&mv, v, head_dim, heads, total_tokens, strides[5] * 2, strides[4] * 2, 64, 1, 128))) return rc;
  hstu::FwdParams p;
  p.cu_seqlens = cu_seqlens; p.num_targets = num_targets; p.num_contexts = num_contexts;
  p.q = static_cast<const __nv_bfloat16*>(q); p.q_t = strides[0]; p.q_h = strides[1];
  p.out = reinterpret_cast<__nv_bfloat16*>(out);
  p.H = heads; p.half_alpha = 0.5f * alpha; p.inv_scale = 1.0f / (float)scaling_seqlen;
  p.target_group = target_group_size; p.win_left = window_left; p.win_right = window_right;
  p.dbg = g_hstu_dbg;
  if (head_dim == 128) return hstu::launch_fwd<128>(mk, mv, p, batch, max_seqlen, (int*)workspace, (cudaStream_t)stream);
  return hstu::launch_fwd<64>(mk, mv, p, batch, max_seqlen, (int*)workspace, (cudaStream_t)stream);
}
