// DynamicEmb row path — sm_100a kernels + C-ABI: fused probe+gather(+pool) forward, row init,
// gather-by-slot, and the fused backward (segmented gradient reduce + sparse optimizer row update).
//
// Replaces reference corelib/dynamicemb/src/{lookup_forward.cu,lookup_kernel.cuh,lookup_backward.cu,
// dynamic_emb_op.cu (load/store flat table, gather_embedding[_pooled], reduce_grads),
// optimizer.cu, optimizer_kernel.cuh, initializer.cu}.  The reference runs probe -> flagged_compact ->
// load_from_flat -> gather as four passes with a staging copy of every unique row; here one warp
// probes 32 ids at a time (32 independent 16-B digest loads in flight) and then streams each found
// 512-B row straight from the value table into the output (or the bag accumulator) with 16-B
// no-allocate loads / streaming stores, 8 rows in flight per warp.  Backward sorts (unique idx, grad
// row) pairs once, then fixed 32-row tiles reduce and apply the optimizer in place in the same kernel.
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "../../include/dynamicemb_b200.h"
#include "demb_common.cuh"
#include "demb_init.cuh"
#include "sm100_ptx.cuh"
#include "demb_probe.cuh"
#include "demb_sort.cuh"

using namespace demb;

namespace {

constexpr int kWarpsPerBlock = 8;
constexpr int kBlock = kWarpsPerBlock * 32;
constexpr int kMaxChunks = 8;   // D <= 1024 (reference limit, lookup_kernel.cuh copy_multi_to_one)

inline int warp_grid(int64_t warps) {
  int64_t blocks = (warps + kWarpsPerBlock - 1) / kWarpsPerBlock;
  int64_t cap = (int64_t)sm_count() * 8 * 4;   // persistent-ish: multiple of the SM count x 8 resident CTAs
  return (int)(blocks < 1 ? 1 : (blocks > cap ? cap : blocks));
}

__device__ __forceinline__ int table_of(const int64_t* __restrict__ range, int T, int64_t i) {
  int lo = 0, hi = T;
  while (hi - lo > 1) { int mid = (lo + hi) >> 1; if (range[mid] <= i) lo = mid; else hi = mid; }
  return lo;
}

// ---- typed output store ---------------------------------------------------------------------------
__device__ __forceinline__ void store_out4(void* out, int dtype, int64_t elem_off, float4 v) {
  if (dtype == DEMB_F32) { st_cs_f4(reinterpret_cast<float*>(out) + elem_off, v); }
  else if (dtype == DEMB_BF16) {
    __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
    uint2 p; p.x = *reinterpret_cast<uint32_t*>(&a); p.y = *reinterpret_cast<uint32_t*>(&b);
    *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(out) + elem_off) = p;
  } else {
    __half2 a = __floats2half2_rn(v.x, v.y), b = __floats2half2_rn(v.z, v.w);
    uint2 p; p.x = *reinterpret_cast<uint32_t*>(&a); p.y = *reinterpret_cast<uint32_t*>(&b);
    *reinterpret_cast<uint2*>(reinterpret_cast<__half*>(out) + elem_off) = p;
  }
}

// Row source for the forward kernels: either probe the hash table per id (eval / inference path,
// `rows`==nullptr) or read a precomputed global row index through `inverse` (training path after
// prefetch: rows[unique idx], inverse[id] = unique idx).
struct RowSrc {
  Table t;
  const uint64_t* keys;       // probe mode
  const int64_t* table_range; // probe mode, ids grouped by table; nullable when T==1
  int T;
  const int64_t* row_base;    // [T] first value row of each table, nullable (0)
  const int64_t* rows;        // indirect mode: global value row per unique (or per id when inverse==nullptr), -1 = absent
  const int64_t* inverse;     // indirect mode, nullable
  uint8_t* founds;            // optional outputs (probe mode)
  int64_t* slots_out;
  const int64_t* n_dev;       // device-side id count (<= the launch bound n), nullable
};
__device__ __forceinline__ int64_t resolve_row(const RowSrc& s, int64_t i) {
  if (s.rows) { int64_t u = s.inverse ? s.inverse[i] : i; return s.rows[u]; }
  const uint64_t key = s.keys[i];
  const int tid = (s.T > 1 && s.table_range) ? table_of(s.table_range, s.T, i) : 0;
  Locus L = locate(s.t, key, tid);
  int64_t slot = -1;
  if (L.cap > 0) {
    int64_t it = probe_thread(s.t, s.t.bucket(L.bucket), key, L.h, nullptr);
    if (it >= 0) slot = (L.bucket - L.bkt_begin) * s.t.C + it;
  }
  if (s.founds) s.founds[i] = slot >= 0;
  if (s.slots_out) s.slots_out[i] = slot;
  return slot < 0 ? -1 : (s.row_base ? s.row_base[tid] : 0) + slot;
}

// ---- forward, sequence mode: out[i,:] = values[row(i), :D] -------------------------------------------
// A12 + A11(load) + A4 fused.  Absent rows produce `absent_value` (reference eval default: zeros).
template <int U>
__global__ void __launch_bounds__(kBlock) forward_seq_kernel(RowSrc s, const float* __restrict__ values, int64_t vdim, int D, int64_t n,
                                                             void* __restrict__ out, int out_dtype, float absent_value) {
  if (s.n_dev) { const int64_t v = *s.n_dev; n = v < 0 ? 0 : (v < n ? v : n); }
  const int lane = threadIdx.x & 31;
  const int D4 = D >> 2;
  const int64_t tiles = (n + 31) >> 5;
  const int64_t wstride = (int64_t)gridDim.x * kWarpsPerBlock;
  for (int64_t tile = (int64_t)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5); tile < tiles; tile += wstride) {
    const int64_t base = tile << 5;
    const int cnt = (int)((n - base) < 32 ? (n - base) : 32);
    int64_t row = -1;
    if (lane < cnt) row = resolve_row(s, base + lane);
    for (int j = 0; j < cnt; j += U) {
      int64_t r[U];
#pragma unroll
      for (int u = 0; u < U; ++u) r[u] = __shfl_sync(0xffffffffu, row, (j + u) & 31);
      for (int c = lane; c < D4; c += 32) {
        float4 v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          if (j + u < cnt && r[u] >= 0) v[u] = ld_nc_f4(values + r[u] * vdim + 4 * c);
          else v[u] = make_float4(absent_value, absent_value, absent_value, absent_value);
        }
#pragma unroll
        for (int u = 0; u < U; ++u)
          if (j + u < cnt) store_out4(out, out_dtype, (base + j + u) * (int64_t)D + 4 * c, v[u]);
      }
    }
  }
}

// ---- forward, sequence mode, fp32 output: TMA-staged variant ------------------------------------------------------------------
// Rows never pass through registers: after the 32-id probe pass each lane issues ONE bulk async copy (cp.async.bulk, 512 B for
// D=128) of its row into the warp's shared-memory stage, completion is tracked by an mbarrier, and because the 32 output rows of a
// tile are contiguous the whole stage leaves with a SINGLE 16 KB bulk store.  12 warps x 16 KB = 192 KB of loads in flight per SM
// (the register kernel above tops out at ~64 KB/SM, occupancy-bound at 98 registers) — a memcpy-shaped pipeline for the gather.
__global__ void __launch_bounds__(384) forward_seq_tma_kernel(RowSrc s, const float* __restrict__ values, int64_t vdim, int D, int64_t n,
                                                              float* __restrict__ out, float absent_value, int warps_per_block) {
  extern __shared__ __align__(128) uint8_t stage_raw[];
  __shared__ uint64_t bars[12];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  if (wib >= warps_per_block) return;
  if (s.n_dev) { const int64_t v = *s.n_dev; n = v < 0 ? 0 : (v < n ? v : n); }
  const uint32_t row_bytes = (uint32_t)D * 4u;
  uint8_t* buf = stage_raw + (size_t)wib * 32u * row_bytes;
  if (lane == 0) { sm100::mbar_init(&bars[wib], 1); sm100::fence_barrier_init(); }
  __syncwarp();
  uint32_t parity = 0;
  const int64_t tiles = (n + 31) >> 5;
  const int64_t wstride = (int64_t)gridDim.x * warps_per_block;
  int64_t tile = (int64_t)blockIdx.x * warps_per_block + wib;
  // Software pipeline: the row ids of tile t+1 (a chain of dependent loads: digest -> key in probe mode, inverse -> rows in gather
  // mode, ~2 us) are resolved while the bulk copies of tile t are in flight, not after them.
  int64_t row = -1;
  if (tile < tiles && (tile << 5) + lane < n) row = resolve_row(s, (tile << 5) + lane);
  for (; tile < tiles; tile += wstride) {
    const int64_t base = tile << 5;
    const int cnt = (int)((n - base) < 32 ? (n - base) : 32);
    const unsigned found = __ballot_sync(0xffffffffu, row >= 0);
    if (lane == 0) {
      sm100::bulk_wait_read0();                                  // the previous tile's bulk store has finished reading this stage
      sm100::mbar_arrive_expect_tx(&bars[wib], (uint32_t)__popc(found) * row_bytes);
    }
    __syncwarp();
    if (row >= 0) {
      sm100::bulk_load(buf + (size_t)lane * row_bytes, values + row * vdim, row_bytes, &bars[wib]);
    } else if (lane < cnt) {
      float4* d = reinterpret_cast<float4*>(buf + (size_t)lane * row_bytes);
      for (int c = 0; c < (D >> 2); ++c) d[c] = make_float4(absent_value, absent_value, absent_value, absent_value);
    }
    const int64_t nbase = (tile + wstride) << 5;
    int64_t next_row = -1;
    if (tile + wstride < tiles && nbase + lane < n) next_row = resolve_row(s, nbase + lane);
    sm100::mbar_wait(&bars[wib], parity);
    parity ^= 1;
    sm100::fence_proxy_async_smem();                             // absent rows were written through the generic proxy
    __syncwarp();
    if (lane == 0) {
      sm100::bulk_store(out + base * (int64_t)D, buf, (uint32_t)cnt * row_bytes);
      sm100::bulk_commit();
    }
    row = next_row;
  }
  if (lane == 0) sm100::bulk_wait0();                             // shared memory must outlive the last bulk store
}

// ---- forward, sequence mode, fp32 output, PROBE mode for 128-slot buckets: hash probe + row gather in one persistent kernel -----------
// The 128-d lookup path of the north star.  One CTA per SM, `warps` independent warps, each running a three-deep software pipeline over
// its tiles of 32 ids (demb_probe.cuh):
//     tile t+2   coalesced loads of the 32 digest lines (one 128-B line per id's bucket) into registers                 [issued]
//     tile t+1   scan the lines (8 lanes per line), confirm candidates by their keys -> 32 value-row ids                [one memory hop]
//     tile t     one cp.async.bulk per row (512 B for D=128) into the warp's shared-memory stage, one 16-KB bulk store to the output
// so a probe costs ONE exposed memory round trip (the candidate key loads), overlapped with the row copies of the previous tile, instead
// of the per-thread chain digest chunk -> key -> next chunk ... of the reference's probe.
constexpr int kProbeFwdWarps = 12;
struct ProbeFwdSmem { uint64_t bar_row[kProbeFwdWarps]; int slot[kProbeFwdWarps][32]; };
template <int GEN>                                               // probe generation (demb_probe.cuh, TileProbe<GEN>)
__global__ void __launch_bounds__(kProbeFwdWarps * 32) forward_seq_probe_kernel(RowSrc s, const float* __restrict__ values, int64_t vdim, int D, int64_t n,
                                                                                float* __restrict__ out, float absent_value, int warps_per_block) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  __shared__ ProbeFwdSmem sm;
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  if (wib >= warps_per_block) return;
  const uint32_t row_bytes = (uint32_t)D * 4u;
  uint8_t* rowbuf = smem_raw + (size_t)wib * 32u * row_bytes;
  if (lane == 0) { sm100::mbar_init(&sm.bar_row[wib], 1); sm100::fence_barrier_init(); }
  __syncwarp();
  const int64_t tiles = (n + 31) >> 5;
  const int64_t wstride = (int64_t)gridDim.x * warps_per_block;
  int64_t tile = (int64_t)blockIdx.x * warps_per_block + wib;
  TableCache tc = empty_table_cache();
  auto row_of = [&](const ProbeKey& p, int pos, int64_t tl) -> int64_t {
    const int64_t i = (tl << 5) + lane;
    const int64_t slot = pos >= 0 ? p.slot_base + pos : -1;
    if (tl < tiles && i < n) { if (s.founds) s.founds[i] = slot >= 0; if (s.slots_out) s.slots_out[i] = slot; }
    return slot < 0 ? -1 : (s.row_base ? s.row_base[p.tid] : 0) + slot;
  };
  uint32_t par_row = 0;
  // Software pipeline, four tiles deep.  The row stage is the scarce resource (in-flight bytes per SM), so the row copies of tile t are
  // issued FIRST in an iteration — right after the previous tile's store — and all probe work for later tiles is done while they fly:
  //   iteration t:  issue row copies (t)
  //                 | compare candidate keys (t+1) [issued an iteration ago] -> its row ids
  //                 | digest masks (t+2) [lines arrived] -> issue its candidate key loads
  //                 | ids (t+3) [arrived] -> hash -> issue its digest-line loads | issue id load (t+4)
  //                 | WAIT rows (t) (the one exposed wait), bulk-store them
  auto load_id = [&](int64_t tl) -> uint64_t {
    const int64_t i = (tl << 5) + lane;
    return (tl < tiles && i < n) ? s.keys[i] : kEmptyKey;          // EmptyKey is a reserved (invalid) key: the lane takes no part
  };
  auto key_of = [&](int64_t tl, uint64_t id) -> ProbeKey {
    const int64_t i = (tl << 5) + lane;
    if (tl >= tiles || i >= n) return ProbeKey{0, 0, 0, 0, false};
    const int tid = (s.T > 1 && s.table_range) ? table_of(s.table_range, s.T, i) : 0;
    return make_probe_key(s.t, id, tid, tc);
  };
  // prologue: tile t probed completely; candidates of t+1 issued; digest lines of t+2 issued; ids of t+3 fetched
  using TP = TileProbe<GEN>;
  typename TP::Dig dg;
  typename TP::Cand cand;
  ProbeKey k1, k2;
  int64_t row;
  {
    const ProbeKey k0 = key_of(tile, load_id(tile));
    TP::load(s.t, k0, dg, lane);
    TP::issue(s.t, k0, dg, cand, lane);
    row = row_of(k0, TP::finish(s.t, k0, cand, sm.slot[wib], lane), tile);
    k1 = key_of(tile + wstride, load_id(tile + wstride));
    TP::load(s.t, k1, dg, lane);
    TP::issue(s.t, k1, dg, cand, lane);
    k2 = key_of(tile + 2 * wstride, load_id(tile + 2 * wstride));
    TP::load(s.t, k2, dg, lane);
  }
  uint64_t id3 = load_id(tile + 3 * wstride);
  for (; tile < tiles; tile += wstride) {
    const int64_t base = tile << 5;
    const int cnt = (int)((n - base) < 32 ? (n - base) : 32);
    const unsigned found = __ballot_sync(0xffffffffu, row >= 0);
    if (lane == 0) {
      sm100::bulk_wait_read0();                                  // the previous tile's bulk store has finished reading the row stage
      sm100::mbar_arrive_expect_tx(&sm.bar_row[wib], (uint32_t)__popc(found) * row_bytes);
    }
    __syncwarp();
    if (row >= 0) {
      sm100::bulk_load(rowbuf + (size_t)lane * row_bytes, values + row * vdim, row_bytes, &sm.bar_row[wib]);
    } else if (lane < cnt) {
      float4* d = reinterpret_cast<float4*>(rowbuf + (size_t)lane * row_bytes);
      for (int c = 0; c < (D >> 2); ++c) d[c] = make_float4(absent_value, absent_value, absent_value, absent_value);
    }
    const int64_t next_row = row_of(k1, TP::finish(s.t, k1, cand, sm.slot[wib], lane), tile + wstride);   // tile t+1
    TP::issue(s.t, k2, dg, cand, lane);                                                                    // tile t+2
    const ProbeKey k3 = key_of(tile + 3 * wstride, id3);                                                           // tile t+3
    TP::load(s.t, k3, dg, lane);
    id3 = load_id(tile + 4 * wstride);                                                                             // tile t+4
    sm100::mbar_wait(&sm.bar_row[wib], par_row);                  // the one exposed wait
    par_row ^= 1;
    sm100::fence_proxy_async_smem();                             // absent rows were written through the generic proxy
    __syncwarp();
    if (lane == 0) {
      sm100::bulk_store(out + base * (int64_t)D, rowbuf, (uint32_t)cnt * row_bytes);
      sm100::bulk_commit();
    }
    row = next_row; k1 = k2; k2 = k3;
  }
  if (lane == 0) sm100::bulk_wait0();                             // shared memory must outlive the last bulk store
}

// ---- the same fused lookup with SPECIALISED warps: 8 probe warps resolve row ids and hand them through a shared-memory ring to 12 copy
// warps that run exactly the loop of the indirect gather kernel (bulk row copies into the warp's stage, one bulk store per tile).  The
// copy warps never wait on a probe: their iteration is the gather kernel's, so the kernel's throughput is the gather's as long as the
// probe warps keep up (one tile per ~5 us per warp).  Tiles of a CTA are numbered i = 0, 1, ...; tile i is probed by warp 12 + i % 8 and
// copied by warp i % 12; ring slot i % 24 therefore always has the same producer and the same consumer.
constexpr int kCopyWarps = 12, kProbeWarps = 8, kRing = 24;
struct Probe2Smem { uint64_t bar_row[kCopyWarps]; uint64_t full[kRing]; uint64_t empty[kRing]; int slot[kProbeWarps][32]; int64_t rowq[kRing][32]; };
__global__ void __launch_bounds__((kCopyWarps + kProbeWarps) * 32, 1) forward_seq_probe2_kernel(RowSrc s, const float* __restrict__ values, int64_t vdim, int D,
                                                                                              int64_t n, float* __restrict__ out, float absent_value) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  __shared__ Probe2Smem sm;
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const uint32_t row_bytes = (uint32_t)D * 4u;
  if (threadIdx.x == 0) {
    for (int i = 0; i < kCopyWarps; ++i) sm100::mbar_init(&sm.bar_row[i], 1);
    for (int i = 0; i < kRing; ++i) { sm100::mbar_init(&sm.full[i], 1); sm100::mbar_init(&sm.empty[i], 1); }
    sm100::fence_barrier_init();
  }
  __syncthreads();
  const int64_t tiles = (n + 31) >> 5;
  const int64_t my_tiles = tiles > (int64_t)blockIdx.x ? (tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;   // tiles blockIdx.x + i * gridDim.x
  if (wib >= kCopyWarps) {
    // ------------------------------------------------------------------ probe warps (producers): two warpgroups, take the registers the
    // copy warpgroups give up (12 x 32 x 64 + 8 x 32 x 136 = 59392 <= 65536)
    asm volatile("setmaxnreg.inc.sync.aligned.u32 136;");
    const int p = wib - kCopyWarps;
    TableCache tc = empty_table_cache();
    auto load_key = [&](int64_t i) -> ProbeKey {
      const int64_t tl = (int64_t)blockIdx.x + i * gridDim.x;
      const int64_t id = (tl << 5) + lane;
      if (i >= my_tiles || id >= n) return ProbeKey{0, 0, 0, 0, false};
      const int tid = (s.T > 1 && s.table_range) ? table_of(s.table_range, s.T, id) : 0;
      return make_probe_key(s.t, s.keys[id], tid, tc);
    };
    ProbeKey k0 = load_key(p);
    DigRegs d0, d1;
    tile_load_digests(s.t, k0, d0, lane);
    for (int64_t i = p; i < my_tiles; i += kProbeWarps) {
      const ProbeKey k1 = load_key(i + kProbeWarps);
      tile_load_digests(s.t, k1, d1, lane);
      const int pos = tile_probe(s.t, k0, d0, sm.slot[p], lane);
      const int64_t id = ((((int64_t)blockIdx.x + i * gridDim.x)) << 5) + lane;
      const int64_t slot = pos >= 0 ? k0.slot_base + pos : -1;
      if (id < n) { if (s.founds) s.founds[id] = slot >= 0; if (s.slots_out) s.slots_out[id] = slot; }
      const int64_t row = slot < 0 ? -1 : (s.row_base ? s.row_base[k0.tid] : 0) + slot;
      const int q = (int)(i % kRing);
      const uint32_t use = (uint32_t)(i / kRing);                 // how many times this slot has been used before
      if (use > 0) sm100::mbar_wait(&sm.empty[q], (use - 1) & 1u);
      sm.rowq[q][lane] = row;
      __syncwarp();
      if (lane == 0) sm100::mbar_arrive(&sm.full[q]);
      k0 = k1; d0 = d1;
    }
    return;
  }
  // -------------------------------------------------------------------- copy warps (consumers): the gather kernel's loop
  asm volatile("setmaxnreg.dec.sync.aligned.u32 64;");
  uint8_t* buf = smem_raw + (size_t)wib * 32u * row_bytes;
  uint32_t par_row = 0;
  for (int64_t i = wib; i < my_tiles; i += kCopyWarps) {
    const int64_t tile = (int64_t)blockIdx.x + i * gridDim.x;
    const int64_t base = tile << 5;
    const int cnt = (int)((n - base) < 32 ? (n - base) : 32);
    const int q = (int)(i % kRing);
    sm100::mbar_wait(&sm.full[q], (uint32_t)(i / kRing) & 1u);
    const int64_t row = lane < cnt ? sm.rowq[q][lane] : -1;
    __syncwarp();
    if (lane == 0) sm100::mbar_arrive(&sm.empty[q]);
    const unsigned found = __ballot_sync(0xffffffffu, row >= 0);
    if (lane == 0) {
      sm100::bulk_wait_read0();
      sm100::mbar_arrive_expect_tx(&sm.bar_row[wib], (uint32_t)__popc(found) * row_bytes);
    }
    __syncwarp();
    if (row >= 0) {
      sm100::bulk_load(buf + (size_t)lane * row_bytes, values + row * vdim, row_bytes, &sm.bar_row[wib]);
    } else if (lane < cnt) {
      float4* d = reinterpret_cast<float4*>(buf + (size_t)lane * row_bytes);
      for (int c = 0; c < (D >> 2); ++c) d[c] = make_float4(absent_value, absent_value, absent_value, absent_value);
    }
    sm100::mbar_wait(&sm.bar_row[wib], par_row);
    par_row ^= 1;
    sm100::fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0) {
      sm100::bulk_store(out + base * (int64_t)D, buf, (uint32_t)cnt * row_bytes);
      sm100::bulk_commit();
    }
  }
  if (lane == 0) sm100::bulk_wait0();
}

// ---- third variant: the specialised layout again, but with MANY small probe warps instead of a few large ones.  probe2's eight probe
// warps resolve ~1.5 tiles/us per SM — exactly the rate the twelve copy warps consume at gather speed, so the copy side starved.  Here a
// CTA is 12 copy warps + 20 probe warps = 1024 threads at 64 registers (the whole register file, no setmaxnreg); a probe warp carries
// nothing between tiles except the next tile's hashed keys, whose digest lines it has already pulled into L2 (prefetch.global.L2), and
// probes with tile_probe_small.  Ring: tile i of the CTA -> slot i % 60, produced by probe warp i % 20, consumed by copy warp i % 12;
// 60 is a multiple of both, so every slot has ONE producer and ONE consumer, each of which sees the slot's barrier phases in order (with
// a ring of 32 two different probe warps shared a slot, one could overtake the other by a whole phase and the parity wait aliased: hang).
template <int kProbe3Warps> struct Ring3 { static constexpr int value = kProbe3Warps == 20 ? 60 : 48; };   // a multiple of both warp counts
template <int kProbe3Warps>
struct Probe3Smem {
  static constexpr int kRing3 = Ring3<kProbe3Warps>::value; uint64_t bar_row[kCopyWarps]; uint64_t full[kRing3]; uint64_t empty[kRing3]; int slot[kProbe3Warps][32]; int64_t rowq[kRing3][32]; };
template <int kProbe3Warps>
__global__ void __launch_bounds__((kCopyWarps + kProbe3Warps) * 32, 1) forward_seq_probe3_kernel(RowSrc s, const float* __restrict__ values, int64_t vdim, int D,
                                                                                               int64_t n, float* __restrict__ out, float absent_value) {
  constexpr int kRing3 = Ring3<kProbe3Warps>::value;
  static_assert(kRing3 % kProbe3Warps == 0 && kRing3 % kCopyWarps == 0, "a ring slot must always have the same producer and the same consumer");
  extern __shared__ __align__(128) uint8_t smem_raw[];
  __shared__ Probe3Smem<kProbe3Warps> sm;
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const uint32_t row_bytes = (uint32_t)D * 4u;
  if (threadIdx.x == 0) {
    for (int i = 0; i < kCopyWarps; ++i) sm100::mbar_init(&sm.bar_row[i], 1);
    for (int i = 0; i < kRing3; ++i) { sm100::mbar_init(&sm.full[i], 1); sm100::mbar_init(&sm.empty[i], 1); }
    sm100::fence_barrier_init();
  }
  __syncthreads();
  const int64_t tiles = (n + 31) >> 5;
  const int64_t my_tiles = tiles > (int64_t)blockIdx.x ? (tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;   // tiles blockIdx.x + i * gridDim.x
  if (wib >= kCopyWarps) {
    // ------------------------------------------------------------------ probe warps (producers)
    const int p = wib - kCopyWarps;
    TableCache tc = empty_table_cache();
    auto load_id = [&](int64_t i) -> uint64_t {
      const int64_t id = ((((int64_t)blockIdx.x + i * gridDim.x)) << 5) + lane;
      return (i < my_tiles && id < n) ? s.keys[id] : kEmptyKey;
    };
    auto key_of = [&](int64_t i, uint64_t key) -> ProbeKey {
      const int64_t id = ((((int64_t)blockIdx.x + i * gridDim.x)) << 5) + lane;
      if (i >= my_tiles || id >= n) return ProbeKey{0, 0, 0, 0, false};
      const int tid = (s.T > 1 && s.table_range) ? table_of(s.table_range, s.T, id) : 0;
      const ProbeKey k = make_probe_key(s.t, key, tid, tc);
      if (k.valid) asm volatile("prefetch.global.L2 [%0];" ::"l"(s.t.digests(s.t.bucket(k.bucket))));
      return k;
    };
    ProbeKey k0 = key_of(p, load_id(p));
    uint64_t id1 = load_id(p + kProbe3Warps);
    for (int64_t i = p; i < my_tiles; i += kProbe3Warps) {
      const ProbeKey k1 = key_of(i + kProbe3Warps, id1);          // next tile: hash, digest lines towards L2
      id1 = load_id(i + 2 * kProbe3Warps);
      const int pos = tile_probe_small(s.t, k0, sm.slot[p], lane);
      const int64_t id = ((((int64_t)blockIdx.x + i * gridDim.x)) << 5) + lane;
      const int64_t slot = pos >= 0 ? k0.slot_base + pos : -1;
      if (id < n) { if (s.founds) s.founds[id] = slot >= 0; if (s.slots_out) s.slots_out[id] = slot; }
      const int64_t row = slot < 0 ? -1 : (s.row_base ? s.row_base[k0.tid] : 0) + slot;
      const int q = (int)(i % kRing3);
      const uint32_t use = (uint32_t)(i / kRing3);                // how many times this slot has been used before
      if (use > 0) sm100::mbar_wait(&sm.empty[q], (use - 1) & 1u);
      sm.rowq[q][lane] = row;
      __syncwarp();
      if (lane == 0) sm100::mbar_arrive(&sm.full[q]);
      k0 = k1;
    }
    return;
  }
  // -------------------------------------------------------------------- copy warps (consumers): the gather kernel's loop
  uint8_t* buf = smem_raw + (size_t)wib * 32u * row_bytes;
  uint32_t par_row = 0;
  for (int64_t i = wib; i < my_tiles; i += kCopyWarps) {
    const int64_t tile = (int64_t)blockIdx.x + i * gridDim.x;
    const int64_t base = tile << 5;
    const int cnt = (int)((n - base) < 32 ? (n - base) : 32);
    const int q = (int)(i % kRing3);
    sm100::mbar_wait(&sm.full[q], (uint32_t)(i / kRing3) & 1u);
    const int64_t row = lane < cnt ? sm.rowq[q][lane] : -1;
    __syncwarp();
    if (lane == 0) sm100::mbar_arrive(&sm.empty[q]);
    const unsigned found = __ballot_sync(0xffffffffu, row >= 0);
    if (lane == 0) {
      sm100::bulk_wait_read0();
      sm100::mbar_arrive_expect_tx(&sm.bar_row[wib], (uint32_t)__popc(found) * row_bytes);
    }
    __syncwarp();
    if (row >= 0) {
      sm100::bulk_load(buf + (size_t)lane * row_bytes, values + row * vdim, row_bytes, &sm.bar_row[wib]);
    } else if (lane < cnt) {
      float4* d = reinterpret_cast<float4*>(buf + (size_t)lane * row_bytes);
      for (int c = 0; c < (D >> 2); ++c) d[c] = make_float4(absent_value, absent_value, absent_value, absent_value);
    }
    sm100::mbar_wait(&sm.bar_row[wib], par_row);
    par_row ^= 1;
    sm100::fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0) {
      sm100::bulk_store(out + base * (int64_t)D, buf, (uint32_t)cnt * row_bytes);
      sm100::bulk_commit();
    }
  }
  if (lane == 0) sm100::bulk_wait0();
}

// ---- forward, pooled mode: ids feature-major (offsets index f*B+b, lookup_forward.cu:53-59);
// out[b, f*D : (f+1)*D] = SUM or MEAN of the bag's rows, fp32 accumulation in id order (lookup_kernel.cuh:901-962).
// A13 + A11 + A4 fused.  A warp takes `bpw` consecutive bags per pass (bpw ~ 32 / average bag length, chosen by the host):
// their ids are contiguous in the id stream, so ONE probe pass keeps all 32 lanes busy (the reference runs one warp per bag:
// 10 active lanes at hotness 10) and the row loads of several bags are in flight together.  Rows are accumulated in id order and
// flushed at every bag boundary, so results do not depend on bpw.
template <int NCHUNK>
__global__ void __launch_bounds__(kBlock) forward_pool_kernel(RowSrc s, const float* __restrict__ values, int64_t vdim, int D, int64_t B, int F,
                                                              const int64_t* __restrict__ offsets, int combiner, void* __restrict__ out,
                                                              int out_dtype, int64_t total_D, int bpw, float absent_value) {
  __shared__ int pool_slot_sm[kWarpsPerBlock][32];
  TableCache pool_tc = empty_table_cache();
  const int lane = threadIdx.x & 31;
  const int D4 = D >> 2;
  const int64_t bags = B * (int64_t)F;
  const int64_t groups = (bags + bpw - 1) / bpw;
  const int64_t wstride = (int64_t)gridDim.x * kWarpsPerBlock;
  for (int64_t grp = (int64_t)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5); grp < groups; grp += wstride) {
    const int64_t g0 = grp * bpw;
    const int nb = (int)((bags - g0) < bpw ? (bags - g0) : bpw);
    // lane l < nb+1 holds offsets[g0 + l]
    const int64_t my_off = lane <= nb ? offsets[g0 + lane] : 0;
    const int64_t beg = __shfl_sync(0xffffffffu, my_off, 0), end = __shfl_sync(0xffffffffu, my_off, nb);
    float4 acc[NCHUNK];
#pragma unroll
    for (int k = 0; k < NCHUNK; ++k) acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    int cur = 0;                                             // bag (within the group) being accumulated
    int64_t cur_end = __shfl_sync(0xffffffffu, my_off, 1);
    auto flush = [&]() {
      const int64_t g = g0 + cur;
      const int64_t f = g / B, b = g - f * B;
      const int64_t len = cur_end - __shfl_sync(0xffffffffu, my_off, cur);
      if (combiner == 1 && len > 0) {
        const float L = (float)len;
#pragma unroll
        for (int k = 0; k < NCHUNK; ++k) { acc[k].x = __fdiv_rn(acc[k].x, L); acc[k].y = __fdiv_rn(acc[k].y, L); acc[k].z = __fdiv_rn(acc[k].z, L); acc[k].w = __fdiv_rn(acc[k].w, L); }
      }
#pragma unroll
      for (int k = 0; k < NCHUNK; ++k) {
        const int c = lane + 32 * k;
        if (c < D4) store_out4(out, out_dtype, b * total_D + f * (int64_t)D + 4 * c, acc[k]);
        acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
      ++cur;
      cur_end = __shfl_sync(0xffffffffu, my_off, (cur + 1) & 31);
    };
    for (int64_t base = beg; base < end; base += 32) {
      const int cnt = (int)((end - base) < 32 ? (end - base) : 32);
      int64_t row = -1;
      if (!s.rows && s.t.C == kProbeC) {
        // probe mode, 128-slot buckets: the warp-tile probe (coalesced digest lines, all candidate keys in flight together)
        ProbeKey pk{0, 0, 0, 0, false};
        if (lane < cnt) pk = make_probe_key(s.t, s.keys[base + lane], (s.T > 1 && s.table_range) ? table_of(s.table_range, s.T, base + lane) : 0, pool_tc);
        DigRegs dg;
        tile_load_digests(s.t, pk, dg, lane);
        const int pos = tile_probe(s.t, pk, dg, pool_slot_sm[threadIdx.x >> 5], lane);
        const int64_t slot = pos >= 0 ? pk.slot_base + pos : -1;
        if (lane < cnt) { if (s.founds) s.founds[base + lane] = slot >= 0; if (s.slots_out) s.slots_out[base + lane] = slot; }
        row = slot < 0 ? -1 : (s.row_base ? s.row_base[pk.tid] : 0) + slot;
      } else if (lane < cnt) {
        row = resolve_row(s, base + lane);
      }
      constexpr int U = 8;
      for (int j = 0; j < cnt; j += U) {
        int64_t r[U];
#pragma unroll
        for (int u = 0; u < U; ++u) r[u] = __shfl_sync(0xffffffffu, row, (j + u) & 31);
        float4 v[U][NCHUNK];
#pragma unroll
        for (int u = 0; u < U; ++u)
#pragma unroll
          for (int k = 0; k < NCHUNK; ++k) {
            const int c = lane + 32 * k;
            // absent ids contribute the eval initializer's constant (reference: eval initializer runs before pooling)
            v[u][k] = (j + u < cnt && c < D4) ? (r[u] >= 0 ? ld_nc_f4(values + r[u] * vdim + 4 * c) : make_float4(absent_value, absent_value, absent_value, absent_value))
                                              : make_float4(0.f, 0.f, 0.f, 0.f);
          }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          if (j + u >= cnt) break;
          while (base + j + u >= cur_end) flush();           // empty bags flush zeros
#pragma unroll
          for (int k = 0; k < NCHUNK; ++k) {                  // in id order => bit-reproducible
            acc[k].x = __fadd_rn(acc[k].x, v[u][k].x); acc[k].y = __fadd_rn(acc[k].y, v[u][k].y);
            acc[k].z = __fadd_rn(acc[k].z, v[u][k].z); acc[k].w = __fadd_rn(acc[k].w, v[u][k].w);
          }
        }
      }
    }
    while (cur < nb) flush();
  }
}

// ---- initializer + store (A10 + A11 store fused): device helpers live in demb_init.cuh (shared with demb_train.cu)
// warp per new row: rows[i] (global value row, <0 = skip), keys[i]; optional `emb_out[i,:D]` copy of the
// initialised embedding (for non-admitted / eval-miss ids that are not stored).
__global__ void __launch_bounds__(kBlock) init_rows_kernel(float* __restrict__ values, int64_t vdim, int D, int64_t n, const int64_t* __restrict__ rows,
                                                           const uint64_t* __restrict__ keys, InitArgs a0, const int64_t* __restrict__ tids,
                                                           const InitArgs* __restrict__ table_init, float state_init,
                                                           const uint8_t* __restrict__ only_if /*nullable: init only where !=0*/,
                                                           float* __restrict__ emb_out) {
  const int lane = threadIdx.x & 31;
  const int D4 = D >> 2, V4 = (int)(vdim >> 2);
  const int64_t wstride = (int64_t)gridDim.x * kWarpsPerBlock;
  for (int64_t i = (int64_t)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5); i < n; i += wstride) {
    if (only_if && !only_if[i]) continue;
    const int64_t r = rows ? rows[i] : -1;
    const uint64_t key = keys[i];
    const InitArgs a = table_init ? table_init[tids ? tids[i] : 0] : a0;
    for (int c = lane; c < D4; c += 32) {
      float4 v = init4(a, key, c);
      if (r >= 0) st_f4(values + r * vdim + 4 * c, v);
      if (emb_out) st_f4(emb_out + i * (int64_t)D + 4 * c, v);
    }
    if (r >= 0)
      for (int c = D4 + lane; c < V4; c += 32) st_f4(values + r * vdim + 4 * c, make_float4(state_init, state_init, state_init, state_init));
  }
}

// rows[i] = row_base[tid[i]] + slot[i]  (or -1)
__global__ void rows_from_slots_kernel(int64_t n, const int64_t* __restrict__ slots, const int64_t* __restrict__ tids,
                                       const int64_t* __restrict__ row_base, int64_t* __restrict__ rows) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t s = slots[i];
    rows[i] = s < 0 ? -1 : ((row_base ? row_base[tids ? tids[i] : 0] : 0) + s);
  }
}

// A11 standalone: copy [emb | state] rows between the value table and a dense [n, width] staging.
__global__ void __launch_bounds__(kBlock) copy_rows_kernel(float* __restrict__ values, int64_t vdim, int width, int64_t n, const int64_t* __restrict__ rows,
                                                           float* __restrict__ dense, int64_t dense_stride, int to_table) {
  const int lane = threadIdx.x & 31;
  const int W4 = width >> 2;
  const int64_t wstride = (int64_t)gridDim.x * kWarpsPerBlock;
  for (int64_t i = (int64_t)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5); i < n; i += wstride) {
    const int64_t r = rows[i];
    if (r < 0) {
      if (!to_table) for (int c = lane; c < W4; c += 32) st_f4(dense + i * dense_stride + 4 * c, make_float4(0.f, 0.f, 0.f, 0.f));
      continue;
    }
    for (int c = lane; c < W4; c += 32) {
      if (to_table) st_f4(values + r * vdim + 4 * c, ld_f4(dense + i * dense_stride + 4 * c));
      else st_f4(dense + i * dense_stride + 4 * c, ld_f4(values + r * vdim + 4 * c));
    }
  }
}

// ---- backward -----------------------------------------------------------------------------------------
struct OptArgs { int type; float lr, eps, beta1, beta2, weight_decay, bc1, bc2; };

// One row update, lane owns float4 chunk(s).  Formulas: optimizer_kernel.cuh:41-404 (IEEE fp32 here; the
// reference compiles with --use_fast_math).  `g` holds the reduced gradient chunks of this lane.
template <int NCHUNK>
__device__ __forceinline__ void apply_row(const OptArgs& o, float* __restrict__ w, int D, const float4 (&g)[NCHUNK], int lane) {
  const int D4 = D >> 2;
  if (o.type == DEMB_OPT_SGD) {
#pragma unroll
    for (int k = 0; k < NCHUNK; ++k) { int c = lane + 32 * k; if (c < D4) {
      float4 x = ld_f4(w + 4 * c);
      x.x -= o.lr * g[k].x; x.y -= o.lr * g[k].y; x.z -= o.lr * g[k].z; x.w -= o.lr * g[k].w;
      st_f4(w + 4 * c, x); } }
  } else if (o.type == DEMB_OPT_ADAGRAD) {
#pragma unroll
    for (int k = 0; k < NCHUNK; ++k) { int c = lane + 32 * k; if (c < D4) {
      float4 x = ld_f4(w + 4 * c), s = ld_f4(w + D + 4 * c);
      s.x += g[k].x * g[k].x; s.y += g[k].y * g[k].y; s.z += g[k].z * g[k].z; s.w += g[k].w * g[k].w;
      x.x -= o.lr * g[k].x / (sqrtf(s.x) + o.eps); x.y -= o.lr * g[k].y / (sqrtf(s.y) + o.eps);
      x.z -= o.lr * g[k].z / (sqrtf(s.z) + o.eps); x.w -= o.lr * g[k].w / (sqrtf(s.w) + o.eps);
      st_f4(w + D + 4 * c, s); st_f4(w + 4 * c, x); } }
  } else if (o.type == DEMB_OPT_ADAM) {
#pragma unroll
    for (int k = 0; k < NCHUNK; ++k) { int c = lane + 32 * k; if (c < D4) {
      float4 x = ld_f4(w + 4 * c), m = ld_f4(w + D + 4 * c), v = ld_f4(w + 2 * D + 4 * c);
      float* xp = &x.x; float* mp = &m.x; float* vp = &v.x; const float* gp = &g[k].x;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        mp[e] = o.beta1 * mp[e] + (1.0f - o.beta1) * gp[e];
        vp[e] = o.beta2 * vp[e] + (1.0f - o.beta2) * gp[e] * gp[e];
        float mh = mp[e] / o.bc1, vh = vp[e] / o.bc2;
        xp[e] -= o.lr * (mh / (sqrtf(vh) + o.eps) + o.weight_decay * xp[e]);
      }
      st_f4(w + D + 4 * c, m); st_f4(w + 2 * D + 4 * c, v); st_f4(w + 4 * c, x); } }
  } else if (o.type == DEMB_OPT_ROWWISE_ADAGRAD) {
    float ss = 0.f;
#pragma unroll
    for (int k = 0; k < NCHUNK; ++k) { int c = lane + 32 * k; if (c < D4) ss += g[k].x * g[k].x + g[k].y * g[k].y + g[k].z * g[k].z + g[k].w * g[k].w; }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, d);
    float acc = w[D] + ss / (float)D;
    __syncwarp();
    if (lane == 0) w[D] = acc;
    const float mult = o.lr / (sqrtf(acc) + o.eps);
#pragma unroll
    for (int k = 0; k < NCHUNK; ++k) { int c = lane + 32 * k; if (c < D4) {
      float4 x = ld_f4(w + 4 * c);
      x.x -= mult * g[k].x; x.y -= mult * g[k].y; x.z -= mult * g[k].z; x.w -= mult * g[k].w;
      st_f4(w + 4 * c, x); } }
  }
}

struct BwdArgs;
__device__ __forceinline__ int64_t bwd_n(const BwdArgs& a);
struct BwdArgs {
  const float* grads; int64_t grad_stride;   // row r of the gradient matrix is grads + r*grad_stride
  int D; int pooled; int combiner; int64_t B; int F; const int64_t* offsets;   // pooled: row id = b*F+f, MEAN scale 1/len(bag f*B+b)
  const int32_t* skey; const int32_t* sval; int64_t n;   // sorted (unique idx, gradient row id); n = launch bound
  const int64_t* n_dev;     // device-side element count (<= n), nullable => n
  const int64_t* ug_addr;   // optional per-unique destination ADDRESS of the reduced gradient row (may be peer memory), overrides unique_grads
  const int64_t* rows;      // unique idx -> global value row (<0: skip), nullable => emit only
  float* values; int64_t vdim;
  float* unique_grads;      // optional [n_unique, D] output of the reduced gradients (reduce_grads op), nullable
  float* part_cont; float* part_start;   // [tiles, D] each
  OptArgs opt;
};

__device__ __forceinline__ int64_t bwd_n(const BwdArgs& a) {
  if (!a.n_dev) return a.n;
  const int64_t v = *a.n_dev;
  return v < 0 ? 0 : (v > a.n ? a.n : v);
}

// r = a.rows[u] when the caller already holds it (kRowKnown), else it is loaded here
template <int NCHUNK, bool kRowKnown = false>
__device__ __forceinline__ void finish_segment(const BwdArgs& a, int32_t u, const float4 (&acc)[NCHUNK], int lane, int64_t r_known = -1) {
  const int D4 = a.D >> 2;
  if (a.ug_addr) {                                              // reduced row goes straight to its owner (possibly over NVLink)
    float* dst = reinterpret_cast<float*>(a.ug_addr[u]);
    if (dst) {
#pragma unroll
      for (int k = 0; k < NCHUNK; ++k) { int c = lane + 32 * k; if (c < D4) st_f4(dst + 4 * c, acc[k]); }
    }
  } else if (a.unique_grads) {
#pragma unroll
    for (int k = 0; k < NCHUNK; ++k) { int c = lane + 32 * k; if (c < D4) st_f4(a.unique_grads + (int64_t)u * a.D + 4 * c, acc[k]); }
  }
  if (a.rows && a.opt.type != DEMB_OPT_NONE) {
    const int64_t r = kRowKnown ? r_known : a.rows[u];
    if (r >= 0) apply_row<NCHUNK>(a.opt, a.values + r * a.vdim, a.D, acc, lane);
  }
}

// Stage 1: warp per 32-row tile of the sorted pair list.  Segments (runs of equal unique idx) that lie
// inside the tile are finished here (reduced gradient -> optimizer update in place).  A segment that
// enters from the previous tile leaves its partial in part_cont[tile]; one that starts here and runs
// past the tile end leaves it in part_start[tile].  Summation order is fixed: ascending sorted position
// inside a tile, tiles in ascending order in stage 2 => bit-reproducible (oracle: oracle/dynamicemb.py).
template <int NCHUNK>
__global__ void __launch_bounds__(kBlock) backward_tiles_kernel(BwdArgs a) {
  const int lane = threadIdx.x & 31;
  const int D4 = a.D >> 2;
  const int64_t n_act = bwd_n(a);
  const int64_t tiles = (n_act + 31) >> 5;
  const int64_t wstride = (int64_t)gridDim.x * kWarpsPerBlock;
  for (int64_t tile = (int64_t)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5); tile < tiles; tile += wstride) {
    const int64_t base = tile << 5;
    const int cnt = (int)((n_act - base) < 32 ? (n_act - base) : 32);
    int32_t myu = -1, myr = 0; float mys = 1.f;
    if (lane < cnt) {
      myu = a.skey[base + lane]; myr = a.sval[base + lane];
      if (a.pooled && a.combiner == 1) {
        int64_t b = myr / a.F, f = myr - b * a.F;
        int64_t len = a.offsets[f * a.B + b + 1] - a.offsets[f * a.B + b];
        mys = 1.0f / (float)len;                                   // lookup_backward.cu:209-241
      }
    }
    // The optimizer's read-modify-write of a value row sits at the END of a dependent chain (pair -> rows[u] -> row); resolve the
    // row id per lane up front and pull the row (embedding + optimizer state) into L2 now, so the RMW after the gradient rows have
    // been summed finds it there instead of paying a DRAM round trip per segment.
    int64_t myrow = -1;
    if (lane < cnt && a.rows && a.opt.type != DEMB_OPT_NONE) {
      myrow = a.rows[myu];
      const int32_t up = __shfl_up_sync(__activemask(), myu, 1);
      if (myrow >= 0 && (lane == 0 || up != myu)) {
        const char* pr = reinterpret_cast<const char*>(a.values + myrow * a.vdim);
        for (int64_t b = 0; b < a.vdim * 4; b += 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(pr + b));
      }
    }
    const int32_t prev_u = base > 0 ? a.skey[base - 1] : -1;
    const int32_t next_u = base + 32 < n_act ? a.skey[base + 32] : -2;
    float4 acc[NCHUNK];
#pragma unroll
    for (int k = 0; k < NCHUNK; ++k) acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    bool incoming = (__shfl_sync(0xffffffffu, myu, 0) == prev_u);
    constexpr int U = 4;
    for (int j = 0; j < cnt; j += U) {
      int32_t uu[U + 1]; int32_t rr[U]; float sc[U]; int64_t vr[U];
#pragma unroll
      for (int u = 0; u < U; ++u) { uu[u] = __shfl_sync(0xffffffffu, myu, (j + u) & 31); rr[u] = __shfl_sync(0xffffffffu, myr, (j + u) & 31); sc[u] = __shfl_sync(0xffffffffu, mys, (j + u) & 31); vr[u] = __shfl_sync(0xffffffffu, myrow, (j + u) & 31); }
      uu[U] = __shfl_sync(0xffffffffu, myu, (j + U) & 31);
      float4 v[U][NCHUNK];
#pragma unroll
      for (int u = 0; u < U; ++u)
#pragma unroll
        for (int k = 0; k < NCHUNK; ++k) {
          const int c = lane + 32 * k;
          v[u][k] = (j + u < cnt && c < D4) ? ld_nc_f4(a.grads + (int64_t)rr[u] * a.grad_stride + 4 * c) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (j + u >= cnt) break;
#pragma unroll
        for (int k = 0; k < NCHUNK; ++k) {
          acc[k].x = __fadd_rn(acc[k].x, __fmul_rn(v[u][k].x, sc[u])); acc[k].y = __fadd_rn(acc[k].y, __fmul_rn(v[u][k].y, sc[u]));
          acc[k].z = __fadd_rn(acc[k].z, __fmul_rn(v[u][k].z, sc[u])); acc[k].w = __fadd_rn(acc[k].w, __fmul_rn(v[u][k].w, sc[u]));
        }
        const bool last_in_tile = (j + u == cnt - 1);
        const int32_t nxt = last_in_tile ? next_u : ((u + 1 < U) ? uu[u + 1] : uu[U]);
        if (nxt != uu[u]) {                       // segment ends at this row (really ends: next row differs)
          if (incoming) {
#pragma unroll
            for (int k = 0; k < NCHUNK; ++k) { int c = lane + 32 * k; if (c < D4) st_f4(a.part_cont + tile * a.D + 4 * c, acc[k]); }
          } else {
            finish_segment<NCHUNK, true>(a, uu[u], acc, lane, vr[u]);
          }
#pragma unroll
          for (int k = 0; k < NCHUNK; ++k) acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
          incoming = false;
        } else if (last_in_tile) {                // runs past the tile end
          float* dst = incoming ? a.part_cont : a.part_start;
#pragma unroll
          for (int k = 0; k < NCHUNK; ++k) { int c = lane + 32 * k; if (c < D4) st_f4(dst + tile * a.D + 4 * c, acc[k]); }
        }
      }
    }
  }
}

// The same stage 1 with the 32 gradient rows of a tile STAGED THROUGH SHARED MEMORY by bulk async copies (one cp.async.bulk per row,
// mbarrier-tracked): the register version keeps 4 rows (2 KB) in flight per warp; here every warp has its whole tile (16 KB for D=128)
// in flight, 12 warps per SM — the pipeline that took the forward gather from 0.55 to 0.79 of the HBM roofline.  Arithmetic and its order
// are unchanged (rows are added in sorted order from the stage), so results are bit-identical to backward_tiles_kernel.
constexpr int kBwdTmaWarps = 12;
template <int NCHUNK>
__global__ void __launch_bounds__(kBwdTmaWarps * 32, 1) backward_tiles_tma_kernel(BwdArgs a, int warps_per_block) {
  extern __shared__ __align__(128) uint8_t stage_raw[];
  __shared__ uint64_t bars[kBwdTmaWarps];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  if (wib >= warps_per_block) return;
  const int D4 = a.D >> 2;
  const uint32_t row_bytes = (uint32_t)a.D * 4u;
  const float* stage = reinterpret_cast<const float*>(stage_raw + (size_t)wib * 32u * row_bytes);
  if (lane == 0) { sm100::mbar_init(&bars[wib], 1); sm100::fence_barrier_init(); }
  __syncwarp();
  uint32_t parity = 0;
  const int64_t n_act = bwd_n(a);
  const int64_t tiles = (n_act + 31) >> 5;
  const int64_t wstride = (int64_t)gridDim.x * warps_per_block;
  for (int64_t tile = (int64_t)blockIdx.x * warps_per_block + wib; tile < tiles; tile += wstride) {
    const int64_t base = tile << 5;
    const int cnt = (int)((n_act - base) < 32 ? (n_act - base) : 32);
    int32_t myu = -1, myr = 0; float mys = 1.f;
    if (lane < cnt) {
      myu = a.skey[base + lane]; myr = a.sval[base + lane];
      if (a.pooled && a.combiner == 1) {
        int64_t b = myr / a.F, f = myr - b * a.F;
        int64_t len = a.offsets[f * a.B + b + 1] - a.offsets[f * a.B + b];
        mys = 1.0f / (float)len;                                   // lookup_backward.cu:209-241
      }
    }
    // gradient rows of the tile -> shared memory (all 32 copies in flight at once)
    if (lane == 0) sm100::mbar_arrive_expect_tx(&bars[wib], (uint32_t)cnt * row_bytes);
    __syncwarp();
    if (lane < cnt) sm100::bulk_load(const_cast<float*>(stage) + (size_t)lane * a.D, a.grads + (int64_t)myr * a.grad_stride, row_bytes, &bars[wib]);
    // value rows (embedding + optimizer state) of the segments that END in this tile: pull them into L2 now
    int64_t myrow = -1;
    if (lane < cnt && a.rows && a.opt.type != DEMB_OPT_NONE) {
      myrow = a.rows[myu];
      const int32_t up = __shfl_up_sync(__activemask(), myu, 1);
      if (myrow >= 0 && (lane == 0 || up != myu)) {
        const char* pr = reinterpret_cast<const char*>(a.values + myrow * a.vdim);
        for (int64_t b = 0; b < a.vdim * 4; b += 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(pr + b));
      }
    }
    const int32_t prev_u = base > 0 ? a.skey[base - 1] : -1;
    const int32_t next_u = base + 32 < n_act ? a.skey[base + 32] : -2;
    float4 acc[NCHUNK];
#pragma unroll
    for (int k = 0; k < NCHUNK; ++k) acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    bool incoming = (__shfl_sync(0xffffffffu, myu, 0) == prev_u);
    sm100::mbar_wait(&bars[wib], parity);
    parity ^= 1;
    for (int j = 0; j < cnt; ++j) {
      const int32_t uj = __shfl_sync(0xffffffffu, myu, j);
      const int32_t un = __shfl_sync(0xffffffffu, myu, (j + 1) & 31);
      const float sc = __shfl_sync(0xffffffffu, mys, j);
      const int64_t vr = __shfl_sync(0xffffffffu, myrow, j);
#pragma unroll
      for (int k = 0; k < NCHUNK; ++k) {
        const int c = lane + 32 * k;
        if (c < D4) {
          const float4 v = *reinterpret_cast<const float4*>(stage + (size_t)j * a.D + 4 * c);
          acc[k].x = __fadd_rn(acc[k].x, __fmul_rn(v.x, sc)); acc[k].y = __fadd_rn(acc[k].y, __fmul_rn(v.y, sc));
          acc[k].z = __fadd_rn(acc[k].z, __fmul_rn(v.z, sc)); acc[k].w = __fadd_rn(acc[k].w, __fmul_rn(v.w, sc));
        }
      }
      const bool last_in_tile = (j == cnt - 1);
      const int32_t nxt = last_in_tile ? next_u : un;
      if (nxt != uj) {                          // segment ends at this row
        if (incoming) {
#pragma unroll
          for (int k = 0; k < NCHUNK; ++k) { int c = lane + 32 * k; if (c < D4) st_f4(a.part_cont + tile * a.D + 4 * c, acc[k]); }
        } else {
          finish_segment<NCHUNK, true>(a, uj, acc, lane, vr);
        }
#pragma unroll
        for (int k = 0; k < NCHUNK; ++k) acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
        incoming = false;
      } else if (last_in_tile) {                // runs past the tile end
        float* dst = incoming ? a.part_cont : a.part_start;
#pragma unroll
        for (int k = 0; k < NCHUNK; ++k) { int c = lane + 32 * k; if (c < D4) st_f4(dst + tile * a.D + 4 * c, acc[k]); }
      }
    }
    __syncwarp();                               // every lane is done reading the stage before the next tile's copies land in it
  }
}

// Ordered sum of `count` consecutive [D]-float partial rows starting at `src` into acc: loads 8 deep, adds in order.
template <int NCHUNK>
__device__ __forceinline__ void add_partials(float4 (&acc)[NCHUNK], const float* __restrict__ src, int count, int D, int lane) {
  const int D4 = D >> 2;
  for (int i0 = 0; i0 < count; i0 += 8) {
    float4 pp[8][NCHUNK];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int k = 0; k < NCHUNK; ++k) {
        const int c = lane + 32 * k;
        pp[i][k] = (i0 + i < count && c < D4) ? ld_f4(src + (int64_t)(i0 + i) * D + 4 * c) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (i0 + i >= count) break;
#pragma unroll
      for (int k = 0; k < NCHUNK; ++k) {
        acc[k].x = __fadd_rn(acc[k].x, pp[i][k].x); acc[k].y = __fadd_rn(acc[k].y, pp[i][k].y);
        acc[k].z = __fadd_rn(acc[k].z, pp[i][k].z); acc[k].w = __fadd_rn(acc[k].w, pp[i][k].w);
      }
    }
  }
}

// Stage 2a: a Zipf-hot id spans thousands of tiles; walking its continuation partials with one warp is a serial chain.
// So first every WINDOW of 32 tiles (1024 sorted rows) that begins inside a segment sums that segment's leading run of
// continuation partials (part_cont) into win_cont[window] — one warp per window, all windows in parallel.
template <int NCHUNK>
__global__ void __launch_bounds__(kBlock) backward_windows_kernel(BwdArgs a, float* __restrict__ win_cont) {
  const int lane = threadIdx.x & 31;
  const int D4 = a.D >> 2;
  const int64_t tiles = (bwd_n(a) + 31) >> 5;
  const int64_t windows = (tiles + 31) >> 5;
  const int64_t wstride = (int64_t)gridDim.x * kWarpsPerBlock;
  for (int64_t w = 1 + (int64_t)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5); w < windows; w += wstride) {
    const int64_t t0 = w << 5;
    const int32_t u = a.skey[t0 << 5];
    if (a.skey[(t0 << 5) - 1] != u) continue;                   // window does not begin inside a segment
    const int64_t tt = t0 + lane;
    const unsigned m = __ballot_sync(0xffffffffu, tt < tiles && a.skey[tt << 5] == u);
    const int run = (m == 0xffffffffu) ? 32 : (__ffs(~m) - 1);
    float4 acc[NCHUNK];
#pragma unroll
    for (int k = 0; k < NCHUNK; ++k) acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    add_partials<NCHUNK>(acc, a.part_cont + t0 * a.D, run, a.D, lane);
#pragma unroll
    for (int k = 0; k < NCHUNK; ++k) { int c = lane + 32 * k; if (c < D4) st_f4(win_cont + w * a.D + 4 * c, acc[k]); }
  }
}

// Stage 2b: warp per tile that owns a START partial: total = start[tile] + cont partials up to the end of its own window
// + win_cont of every following window the segment reaches; then finish the segment.  Fixed order => bit-reproducible.
template <int NCHUNK>
__global__ void __launch_bounds__(kBlock) backward_spans_kernel(BwdArgs a, const float* __restrict__ win_cont) {
  const int lane = threadIdx.x & 31;
  const int D4 = a.D >> 2;
  const int64_t tiles = (bwd_n(a) + 31) >> 5;
  const int64_t windows = (tiles + 31) >> 5;
  const int64_t wstride = (int64_t)gridDim.x * kWarpsPerBlock;
  for (int64_t tile = (int64_t)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5); tile + 1 < tiles; tile += wstride) {
    const int64_t base = tile << 5;
    const int32_t u = a.skey[base + 31];
    if (a.skey[base + 32] != u) continue;                       // last segment does not spill
    if (a.skey[base] == u && base > 0 && a.skey[base - 1] == u) continue;   // whole tile is a continuation
    float4 acc[NCHUNK];
#pragma unroll
    for (int k = 0; k < NCHUNK; ++k) { int c = lane + 32 * k; acc[k] = c < D4 ? ld_f4(a.part_start + tile * a.D + 4 * c) : make_float4(0.f, 0.f, 0.f, 0.f); }
    // continuation tiles inside this tile's own window
    const int64_t w0 = tile >> 5;
    const int64_t wend = ((w0 + 1) << 5) < tiles ? ((w0 + 1) << 5) : tiles;
    {
      const int64_t tt = tile + 1 + lane;
      const unsigned m = __ballot_sync(0xffffffffu, tt < wend && a.skey[tt << 5] == u);
      const int run = (m == 0xffffffffu) ? 32 : (__ffs(~m) - 1);
      add_partials<NCHUNK>(acc, a.part_cont + (tile + 1) * a.D, run, a.D, lane);
    }
    // following windows that begin inside this segment
    for (int64_t w = w0 + 1; w < windows; w += 32) {
      const int64_t ww = w + lane;
      const unsigned m = __ballot_sync(0xffffffffu, ww < windows && a.skey[ww << 10] == u);
      const int run = (m == 0xffffffffu) ? 32 : (__ffs(~m) - 1);
      add_partials<NCHUNK>(acc, win_cont + w * a.D, run, a.D, lane);
      if (run < 32) break;
    }
    finish_segment<NCHUNK>(a, u, acc, lane);
  }
}

// sort keys / payload for backward: key = inverse[i] (unique idx); payload = gradient row id
// (sequence: i; pooled: b*F+f of the bag holding id i — generate_gather_ids_pooled_kernel, lookup_backward.cu).
// n_dev (nullable): real element count; positions beyond it get the key `pad_key` (> every unique idx) so they sort to the end.
// grad_row_of (nullable, sequence mode): gradient row id of id i when the gradient rows are not stored in id order.
__global__ void backward_pairs_kernel(int64_t n, const int64_t* __restrict__ n_dev, int32_t pad_key, const int64_t* __restrict__ inverse,
                                      const int64_t* __restrict__ grad_row_of, int pooled, int64_t B, int F, const int64_t* __restrict__ offsets,
                                      int32_t* __restrict__ key, int32_t* __restrict__ val) {
  int64_t n_act = n;
  if (n_dev) { n_act = *n_dev; n_act = n_act < 0 ? 0 : (n_act > n ? n : n_act); }
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    if (i >= n_act) { key[i] = pad_key; val[i] = 0; continue; }
    key[i] = (int32_t)inverse[i];
    if (!pooled) { val[i] = grad_row_of ? (int32_t)grad_row_of[i] : (int32_t)i; continue; }
    int64_t lo = 0, hi = B * F;   // bag g with offsets[g] <= i < offsets[g+1]
    while (hi - lo > 1) { int64_t mid = (lo + hi) >> 1; if (offsets[mid] <= i) lo = mid; else hi = mid; }
    int64_t f = lo / B, b = lo - f * B;
    val[i] = (int32_t)(b * F + f);
  }
}

// A15 standalone: warp per unique row with a dense [n, D] gradient.
template <int NCHUNK>
__global__ void __launch_bounds__(kBlock) update_rows_kernel(float* __restrict__ values, int64_t vdim, int D, int64_t n, const int64_t* __restrict__ rows,
                                                             const float* __restrict__ grads, int64_t grad_stride, OptArgs o) {
  const int lane = threadIdx.x & 31;
  const int D4 = D >> 2;
  const int64_t wstride = (int64_t)gridDim.x * kWarpsPerBlock;
  for (int64_t i = (int64_t)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5); i < n; i += wstride) {
    const int64_t r = rows[i];
    if (r < 0) continue;
    float4 g[NCHUNK];
#pragma unroll
    for (int k = 0; k < NCHUNK; ++k) { int c = lane + 32 * k; g[k] = c < D4 ? ld_nc_f4(grads + i * grad_stride + 4 * c) : make_float4(0.f, 0.f, 0.f, 0.f); }
    apply_row<NCHUNK>(o, values + r * vdim, D, g, lane);
  }
}

// ---- flat multi-table load / store / update through per-table base pointers (dynamic_emb_op.cu:295-490, optimizer.cu) ---------------
// The reference keeps one value buffer per table and hands kernels an int64 array of base pointers; rows may have different widths
// per table (mixed embedding dims).  region: 0 = contiguous prefix of min(value_dim, dense_dim) floats, 1 = embedding only,
// 2 = [emb | pad to max_emb_dim | optimizer state].  Warp per row; indices < 0 are skipped.
struct FlatArgs {
  const int64_t* table_ptrs; const int64_t* table_ids; int64_t scalar_table_id; const int64_t* indices; int64_t n;
  const int64_t* value_dims; const int64_t* emb_dims; int64_t max_emb_dim;
  float* dense; int64_t dense_stride; int64_t dense_dim; int region; int to_table;
};
__device__ __forceinline__ void copy_span(const float* __restrict__ src, float* __restrict__ dst, int64_t len, int lane) {
  const bool vec = ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15) == 0;
  const int64_t v4 = vec ? (len >> 2) : 0;
  for (int64_t c = lane; c < v4; c += 32) st_f4(dst + 4 * c, ld_f4(src + 4 * c));
  for (int64_t i = 4 * v4 + lane; i < len; i += 32) dst[i] = src[i];
}
__global__ void __launch_bounds__(kBlock) flat_table_copy_kernel(FlatArgs a) {
  const int lane = threadIdx.x & 31;
  const int64_t wstride = (int64_t)gridDim.x * kWarpsPerBlock;
  for (int64_t i = (int64_t)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5); i < a.n; i += wstride) {
    const int64_t idx = a.indices[i];
    if (idx < 0) continue;
    const int64_t t = a.table_ids ? a.table_ids[i] : a.scalar_table_id;
    const int64_t vdim = a.value_dims[t], edim = a.emb_dims[t];
    float* row = reinterpret_cast<float*>(a.table_ptrs[t]) + idx * vdim;
    float* d = a.dense + i * a.dense_stride;
    int64_t len0, off_t1 = 0, off_d1 = 0, len1 = 0;
    if (a.region == 0) len0 = vdim < a.dense_dim ? vdim : a.dense_dim;
    else if (a.region == 1) len0 = edim < a.dense_dim ? edim : a.dense_dim;
    else { len0 = edim; off_t1 = edim; off_d1 = a.max_emb_dim; len1 = vdim - edim; }
    if (a.to_table) { copy_span(d, row, len0, lane); if (len1 > 0) copy_span(d + off_d1, row + off_t1, len1, lane); }
    else { copy_span(row, d, len0, lane); if (len1 > 0) copy_span(row + off_t1, d + off_d1, len1, lane); }
  }
}
template <int NCHUNK>
__global__ void __launch_bounds__(kBlock) flat_table_update_kernel(FlatArgs a, const float* __restrict__ grads, int64_t grad_stride, OptArgs o) {
  const int lane = threadIdx.x & 31;
  const int64_t wstride = (int64_t)gridDim.x * kWarpsPerBlock;
  for (int64_t i = (int64_t)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5); i < a.n; i += wstride) {
    const int64_t idx = a.indices[i];
    if (idx < 0) continue;
    const int64_t t = a.table_ids ? a.table_ids[i] : a.scalar_table_id;
    const int D = (int)a.emb_dims[t];
    float* row = reinterpret_cast<float*>(a.table_ptrs[t]) + idx * a.value_dims[t];
    float4 g[NCHUNK];
#pragma unroll
    for (int k = 0; k < NCHUNK; ++k) { const int c = lane + 32 * k; g[k] = c < (D >> 2) ? ld_nc_f4(grads + i * grad_stride + 4 * c) : make_float4(0.f, 0.f, 0.f, 0.f); }
    apply_row<NCHUNK>(o, row, D, g, lane);
  }
}

size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }
inline int nchunk_of(int D) { return (D / 4 + 31) / 32; }

// bench.py's roofline needs per-kernel device time of the multi-kernel backward: optional CUDA events around its stages.
int g_dev_opts[8] = {1, 1, 1, 1, 1, 0, 3, 1};   // [2] route histogram in smem, [3] full-bucket eviction shortcut, [4] multiply-shift in unique, [5] TMA-staged gather_to_peers,
                                                // [6] CTAs per SM of evict / init when a consumer is hooked on the prefetch's lookup stage, [7] CTAs per SM of gather_to_peers part 1
int g_dev_opts_pad_;  //   // demb_set_option(2..7): development toggles read by other translation units (demb_get_option)
bool g_bwd_tma = false;         // demb_set_option(1, v): gradient rows of the backward staged through shared memory (1) or registers (0, default:
                                // measured 0.293 ms against 0.367 ms — 12 resident warps cannot hide the per-segment row read-modify-write chain)
int g_probe_kernel = 1;         // demb_set_option(0, v): 1 = one probe + copy pipeline per warp (default), 2 = specialised probe / copy warps (measured slower), 0 = round-1 thread-per-key probe
bool g_prof_on = false;
cudaEvent_t g_prof_ev[4] = {nullptr, nullptr, nullptr, nullptr};

}  // namespace

#define DISPATCH_NCHUNK(D, ...)                                          \
  switch (nchunk_of(D)) {                                                \
    case 1: { constexpr int NC = 1; __VA_ARGS__; break; }                \
    case 2: { constexpr int NC = 2; __VA_ARGS__; break; }                \
    case 3: case 4: { constexpr int NC = 4; __VA_ARGS__; break; }        \
    default: { constexpr int NC = 8; __VA_ARGS__; break; }               \
  }

extern "C" {

// bags a warp takes per pass: as many average-length bags as fit in one 32-lane probe pass (1..16)
static int pool_bags_per_warp(int64_t n_ids, int64_t bags) {
  if (bags <= 0 || n_ids <= 0) return 1;
  int64_t avg = (n_ids + bags - 1) / bags;
  int64_t bpw = 32 / (avg < 1 ? 1 : avg);
  return (int)(bpw < 1 ? 1 : (bpw > 16 ? 16 : bpw));
}
static int launch_seq_tma(const RowSrc& s, const float* values, int64_t value_dim, int emb_dim, int64_t n, float* out, float absent_value,
                          cudaStream_t stream) {
  const size_t stage = 32u * (size_t)emb_dim * 4u;                 // bytes per warp stage
  int warps = (int)((216u * 1024u) / stage);
  if (warps > 12) warps = 12;
  if (warps < 1) return DEMB_ERR_ARG;
  const int smem = (int)(warps * stage);
  static std::atomic<int> configured[kMaxDevices];
  cudaError_t ce = once_per_device(configured, [] { return cudaFuncSetAttribute(forward_seq_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 216 * 1024); });
  if (ce != cudaSuccess) return -(int)ce;
  const int64_t tiles = (n + 31) / 32;
  int64_t blocks = (tiles + warps - 1) / warps;
  if (blocks > sm_count()) blocks = sm_count();                    // one CTA per SM owns the whole shared memory; persistent over tiles
  forward_seq_tma_kernel<<<(int)blocks, 384, smem, stream>>>(s, values, value_dim, emb_dim, n, out, absent_value, warps);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : -(int)e;
}
static int launch_seq_probe(int gen, const RowSrc& s, const float* values, int64_t value_dim, int emb_dim, int64_t n, float* out, float absent_value,
                            cudaStream_t stream) {
  const size_t per_warp = 32u * (size_t)emb_dim * 4u;
  int warps = (int)((216u * 1024u) / per_warp);
  if (warps > kProbeFwdWarps) warps = kProbeFwdWarps;
  if (warps < 1) return DEMB_ERR_ARG;
  const int smem = (int)(warps * per_warp);
  static std::atomic<int> configured[kMaxDevices];
  cudaError_t ce = once_per_device(configured, [] {
    cudaError_t e = cudaFuncSetAttribute(forward_seq_probe_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 216 * 1024);
    return e != cudaSuccess ? e : cudaFuncSetAttribute(forward_seq_probe_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 216 * 1024);
  });
  if (ce != cudaSuccess) return -(int)ce;
  const int64_t tiles = (n + 31) / 32;
  int64_t blocks = (tiles + warps - 1) / warps;
  if (blocks > sm_count()) blocks = sm_count();
  if (gen == 2) forward_seq_probe_kernel<2><<<(int)blocks, kProbeFwdWarps * 32, smem, stream>>>(s, values, value_dim, emb_dim, n, out, absent_value, warps);
  else forward_seq_probe_kernel<1><<<(int)blocks, kProbeFwdWarps * 32, smem, stream>>>(s, values, value_dim, emb_dim, n, out, absent_value, warps);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : -(int)e;
}
static int launch_seq_probe2(const RowSrc& s, const float* values, int64_t value_dim, int emb_dim, int64_t n, float* out, float absent_value,
                             cudaStream_t stream) {
  const int smem = kCopyWarps * 32 * emb_dim * 4;
  static std::atomic<int> configured[kMaxDevices];
  cudaError_t ce = once_per_device(configured, [] { return cudaFuncSetAttribute(forward_seq_probe2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); });
  if (ce != cudaSuccess) return -(int)ce;
  const int64_t tiles = (n + 31) / 32;
  int64_t blocks = (tiles + kCopyWarps - 1) / kCopyWarps;
  if (blocks > sm_count()) blocks = sm_count();
  forward_seq_probe2_kernel<<<(int)blocks, (kCopyWarps + kProbeWarps) * 32, smem, stream>>>(s, values, value_dim, emb_dim, n, out, absent_value);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : -(int)e;
}
static int launch_seq_probe3(int pw, const RowSrc& s, const float* values, int64_t value_dim, int emb_dim, int64_t n, float* out, float absent_value,
                             cudaStream_t stream) {
  const int smem = kCopyWarps * 32 * emb_dim * 4;
  static std::atomic<int> configured[kMaxDevices];
  cudaError_t ce = once_per_device(configured, [] {
    cudaError_t e = cudaFuncSetAttribute(forward_seq_probe3_kernel<20>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    return e != cudaSuccess ? e : cudaFuncSetAttribute(forward_seq_probe3_kernel<12>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  });
  if (ce != cudaSuccess) return -(int)ce;
  const int64_t tiles = (n + 31) / 32;
  int64_t blocks = (tiles + kCopyWarps - 1) / kCopyWarps;
  if (blocks > sm_count()) blocks = sm_count();
  if (pw == 20) forward_seq_probe3_kernel<20><<<(int)blocks, (kCopyWarps + 20) * 32, smem, stream>>>(s, values, value_dim, emb_dim, n, out, absent_value);
  else forward_seq_probe3_kernel<12><<<(int)blocks, (kCopyWarps + 12) * 32, smem, stream>>>(s, values, value_dim, emb_dim, n, out, absent_value);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : -(int)e;
}
static int check_dims(int D, int64_t vdim) { return (D <= 0 || (D & 3) || D > 128 * kMaxChunks || (vdim & 3) || vdim < D) ? DEMB_ERR_ARG : 0; }

int demb_lookup_forward(void* storage, const int64_t* table_bucket_offsets, int64_t bucket_capacity, int num_scores, const float* values,
                        int64_t value_dim, int emb_dim, const int64_t* row_base, int64_t n, const void* keys, const int64_t* table_range,
                        int num_tables, const int64_t* offsets, int64_t batch_size, int num_features, int combiner, void* out, int out_dtype,
                        float absent_value, uint8_t* founds, int64_t* slots_out, void* stream) {
  if (check_dims(emb_dim, value_dim)) return DEMB_ERR_ARG;
  RowSrc s{Table{(uint8_t*)storage, table_bucket_offsets, bucket_capacity, num_scores}, (const uint64_t*)keys, table_range, num_tables, row_base,
           nullptr, nullptr, founds, slots_out, nullptr};
  if (combiner < 0) {
    if (n <= 0) return 0;
    if (out_dtype == DEMB_F32 && bucket_capacity == kProbeC && g_probe_kernel == 4 && kCopyWarps * 32 * emb_dim * 4 <= 200 * 1024)
      return launch_seq_probe3(12, s, values, value_dim, emb_dim, n, (float*)out, absent_value, (cudaStream_t)stream);
    if (out_dtype == DEMB_F32 && bucket_capacity == kProbeC && g_probe_kernel == 3 && kCopyWarps * 32 * emb_dim * 4 <= 200 * 1024)
      return launch_seq_probe3(20, s, values, value_dim, emb_dim, n, (float*)out, absent_value, (cudaStream_t)stream);
    if (out_dtype == DEMB_F32 && bucket_capacity == kProbeC && g_probe_kernel == 2 && kCopyWarps * 32 * emb_dim * 4 <= 200 * 1024)
      return launch_seq_probe2(s, values, value_dim, emb_dim, n, (float*)out, absent_value, (cudaStream_t)stream);
    if (out_dtype == DEMB_F32 && bucket_capacity == kProbeC && emb_dim <= 1024 && g_probe_kernel)
      return launch_seq_probe(g_probe_kernel == 5 ? 2 : 1, s, values, value_dim, emb_dim, n, (float*)out, absent_value, (cudaStream_t)stream);
    if (out_dtype == DEMB_F32) return launch_seq_tma(s, values, value_dim, emb_dim, n, (float*)out, absent_value, (cudaStream_t)stream);
    forward_seq_kernel<8><<<warp_grid((n + 31) / 32), kBlock, 0, (cudaStream_t)stream>>>(s, values, value_dim, emb_dim, n, out, out_dtype, absent_value);
  } else {
    if (batch_size <= 0 || num_features <= 0) return 0;
    if (emb_dim > 512) return DEMB_ERR_ARG;
    int64_t bags = batch_size * num_features;
    const int bpw = pool_bags_per_warp(n, bags);
    DISPATCH_NCHUNK(emb_dim, forward_pool_kernel<NC><<<warp_grid((bags + bpw - 1) / bpw), kBlock, 0, (cudaStream_t)stream>>>(
                                 s, values, value_dim, emb_dim, batch_size, num_features, offsets, combiner, out, out_dtype,
                                 (int64_t)num_features * emb_dim, bpw, absent_value));
  }
  DEMB_CHECK_LAST();
  return 0;
}

int demb_gather_forward(const float* values, int64_t value_dim, int emb_dim, int64_t n, const int64_t* rows, const int64_t* inverse,
                        const int64_t* offsets, int64_t batch_size, int num_features, int combiner, void* out, int out_dtype, const int64_t* n_dev,
                        void* stream) {
  if (check_dims(emb_dim, value_dim)) return DEMB_ERR_ARG;
  if (n_dev && combiner >= 0) return DEMB_ERR_ARG;                  // a device-side count only makes sense for the sequence layout
  RowSrc s{Table{nullptr, nullptr, 0, 1}, nullptr, nullptr, 1, nullptr, rows, inverse, nullptr, nullptr, n_dev};
  if (combiner < 0) {
    if (n <= 0) return 0;
    if (out_dtype == DEMB_F32) return launch_seq_tma(s, values, value_dim, emb_dim, n, (float*)out, 0.f, (cudaStream_t)stream);
    forward_seq_kernel<8><<<warp_grid((n + 31) / 32), kBlock, 0, (cudaStream_t)stream>>>(s, values, value_dim, emb_dim, n, out, out_dtype, 0.f);
  } else {
    if (batch_size <= 0 || num_features <= 0) return 0;
    if (emb_dim > 512) return DEMB_ERR_ARG;
    int64_t bags = batch_size * num_features;
    const int bpw = pool_bags_per_warp(n, bags);
    DISPATCH_NCHUNK(emb_dim, forward_pool_kernel<NC><<<warp_grid((bags + bpw - 1) / bpw), kBlock, 0, (cudaStream_t)stream>>>(
                                 s, values, value_dim, emb_dim, batch_size, num_features, offsets, combiner, out, out_dtype,
                                 (int64_t)num_features * emb_dim, bpw, 0.f));
  }
  DEMB_CHECK_LAST();
  return 0;
}

int demb_rows_from_slots(int64_t n, const int64_t* slots, const int64_t* table_ids, const int64_t* row_base, int64_t* rows, void* stream) {
  if (n <= 0) return 0;
  int grid = (int)((n + 255) / 256);
  rows_from_slots_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(n, slots, table_ids, row_base, rows);
  DEMB_CHECK_LAST();
  return 0;
}

int demb_init_rows(float* values, int64_t value_dim, int emb_dim, int64_t n, const int64_t* rows, const void* keys, int mode, float p0, float p1,
                   float p2, float p3, uint64_t seed, const int64_t* table_ids, const demb_init_args_t* table_init, float state_init,
                   const uint8_t* only_if, float* emb_out, void* stream) {
  if (check_dims(emb_dim, value_dim)) return DEMB_ERR_ARG;
  if (n <= 0) return 0;
  InitArgs a{mode, p0, p1, p2, p3, seed};
  init_rows_kernel<<<warp_grid(n), kBlock, 0, (cudaStream_t)stream>>>(values, value_dim, emb_dim, n, rows, (const uint64_t*)keys, a, table_ids,
                                                                      reinterpret_cast<const InitArgs*>(table_init), state_init, only_if, emb_out);
  DEMB_CHECK_LAST();
  return 0;
}

int demb_copy_rows(float* values, int64_t value_dim, int width, int64_t n, const int64_t* rows, float* dense, int64_t dense_stride, int to_table,
                   void* stream) {
  if (width <= 0 || (width & 3) || width > value_dim || (dense_stride & 3)) return DEMB_ERR_ARG;
  if (n <= 0) return 0;
  copy_rows_kernel<<<warp_grid(n), kBlock, 0, (cudaStream_t)stream>>>(values, value_dim, width, n, rows, dense, dense_stride, to_table);
  DEMB_CHECK_LAST();
  return 0;
}

int64_t demb_backward_workspace_bytes(int64_t n, int emb_dim) {
  if (n <= 0) return 256;
  const size_t tmp = rsort::workspace_bytes(n);
  size_t tiles = ((size_t)n + 31) / 32;
  return (int64_t)(4 * align256(4 * (size_t)n) + 2 * align256(tiles * (size_t)emb_dim * 4) + align256(((tiles + 31) / 32 + 1) * (size_t)emb_dim * 4) + align256(tmp) + 256);
}

// phase 0: everything.  phase 1: only the gradient-independent part (pair list + radix sort by unique index).  phase 2: the rest (same
// workspace).  All on `stream`: a caller that wants phase 1 to overlap the forward gather launches it on a stream of its own and orders
// phase 2 behind it with events (dynamicemb_extensions.BackwardPrep does, with torch streams so the caching allocator knows).
static int backward_impl(float* values, int64_t value_dim, int emb_dim, int64_t n, const int64_t* inverse, int64_t num_unique_bound, const int64_t* rows,
                         const float* grads, int64_t grad_stride, const int64_t* offsets, int64_t batch_size, int num_features, int combiner,
                         int opt_type, float lr, float eps, float beta1, float beta2, float weight_decay, float bias_correction1,
                         float bias_correction2, float* unique_grads, const int64_t* n_dev, const int64_t* grad_row_of, const int64_t* ug_addr,
                         void* workspace, int64_t workspace_bytes, cudaStream_t stream, int phase) {
  if (check_dims(emb_dim, value_dim > 0 ? value_dim : emb_dim)) return DEMB_ERR_ARG;
  if (n <= 0) return 0;
  if (n >= (1ll << 31) || num_unique_bound >= (1ll << 31)) return DEMB_ERR_ARG;
  if (workspace_bytes < demb_backward_workspace_bytes(n, emb_dim)) return DEMB_ERR_WORKSPACE;
  const int pooled = combiner >= 0;
  size_t tiles = ((size_t)n + 31) / 32;
  uint8_t* w = (uint8_t*)workspace;
  int32_t* k0 = (int32_t*)w; w += align256(4 * (size_t)n);
  int32_t* v0 = (int32_t*)w; w += align256(4 * (size_t)n);
  int32_t* k1 = (int32_t*)w; w += align256(4 * (size_t)n);
  int32_t* v1 = (int32_t*)w; w += align256(4 * (size_t)n);
  float* pc = (float*)w; w += align256(tiles * (size_t)emb_dim * 4);
  float* ps = (float*)w; w += align256(tiles * (size_t)emb_dim * 4);
  float* wc = (float*)w; w += align256(((tiles + 31) / 32 + 1) * (size_t)emb_dim * 4);
  size_t tmp_bytes = (size_t)((uint8_t*)workspace + workspace_bytes - w);
  int end_bit = 1; while (end_bit < 31 && (1ll << end_bit) < num_unique_bound + (n_dev ? 1 : 0)) ++end_bit;
  const bool in_first = n <= 1 || (rsort::num_passes(end_bit) % 2 == 0);                            // where the sorted pairs end up
  const int32_t* skey = in_first ? k0 : k1;
  const int32_t* sval = in_first ? v0 : v1;
  if (phase != 2) {
    cudaStream_t s1 = stream;
    if (g_prof_on && phase == 0) cudaEventRecord(g_prof_ev[0], stream);
    backward_pairs_kernel<<<(int)((n + 255) / 256), 256, 0, s1>>>(n, n_dev, (int32_t)num_unique_bound, inverse, grad_row_of, pooled, batch_size, num_features,
                                                                  offsets, k0, v0);
    int32_t *sk = nullptr, *sv = nullptr;
    const int rc = rsort::sort_pairs(k0, v0, k1, v1, n, end_bit, w, tmp_bytes, s1, &sk, &sv);       // own stable LSD radix sort (demb_sort.cuh)
    if (rc) return rc;
    if (phase == 1) {
      DEMB_CHECK_LAST();
      return 0;
    }
  } else {
    if (g_prof_on) cudaEventRecord(g_prof_ev[0], stream);
  }
  BwdArgs a{grads, grad_stride, emb_dim, pooled, combiner, batch_size, num_features, offsets, skey, sval, n, n_dev, ug_addr, rows, values, value_dim, unique_grads,
            pc, ps, OptArgs{opt_type, lr, eps, beta1, beta2, weight_decay, bias_correction1, bias_correction2}};
  if (g_prof_on) cudaEventRecord(g_prof_ev[1], stream);
  DISPATCH_NCHUNK(emb_dim, {
    const size_t stage_b = 32u * (size_t)emb_dim * 4u;
    int tw = (int)((216u * 1024u) / stage_b); if (tw > kBwdTmaWarps) tw = kBwdTmaWarps;
    if (g_bwd_tma && tw >= 4 && (((uintptr_t)grads | (uintptr_t)(grad_stride * 4)) & 15) == 0) {
      static std::atomic<int> configured[kMaxDevices];
      cudaError_t ce = once_per_device(configured, [] { return cudaFuncSetAttribute(backward_tiles_tma_kernel<NC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 216 * 1024); });
      if (ce != cudaSuccess) return -(int)ce;
      int64_t blocks = ((int64_t)tiles + tw - 1) / tw;
      if (blocks > sm_count()) blocks = sm_count();
      backward_tiles_tma_kernel<NC><<<(int)blocks, kBwdTmaWarps * 32, (size_t)tw * stage_b, stream>>>(a, tw);
    } else {
      backward_tiles_kernel<NC><<<warp_grid((int64_t)tiles), kBlock, 0, stream>>>(a);
    }
    if (g_prof_on) cudaEventRecord(g_prof_ev[2], stream);
    if (tiles > 32) backward_windows_kernel<NC><<<warp_grid((int64_t)(tiles + 31) / 32), kBlock, 0, stream>>>(a, wc);
    if (tiles > 1) backward_spans_kernel<NC><<<warp_grid((int64_t)tiles - 1), kBlock, 0, stream>>>(a, wc);
    if (g_prof_on) cudaEventRecord(g_prof_ev[3], stream);
  });
  DEMB_CHECK_LAST();
  return 0;
}

// Fused backward: reduce gradients per unique id and apply the sparse optimizer to the value rows.
//  grads: sequence mode [n, D] (row i = gradient of id i); pooled mode [B, F*D] viewed as [B*F, D].
//  inverse[n]: id -> unique idx in [0, num_unique_bound).  rows[u]: global value row of unique u (<0 skip).
//  unique_grads (nullable): also emit the reduced gradients [num_unique, D] (reference op reduce_grads).
int demb_backward(float* values, int64_t value_dim, int emb_dim, int64_t n, const int64_t* inverse, int64_t num_unique_bound, const int64_t* rows,
                  const float* grads, int64_t grad_stride, const int64_t* offsets, int64_t batch_size, int num_features, int combiner,
                  int opt_type, float lr, float eps, float beta1, float beta2, float weight_decay, float bias_correction1,
                  float bias_correction2, float* unique_grads, const int64_t* n_dev, const int64_t* grad_row_of, const int64_t* unique_grad_addr,
                  void* workspace, int64_t workspace_bytes, void* stream_) {
  return backward_impl(values, value_dim, emb_dim, n, inverse, num_unique_bound, rows, grads, grad_stride, offsets, batch_size, num_features, combiner,
                       opt_type, lr, eps, beta1, beta2, weight_decay, bias_correction1, bias_correction2, unique_grads, n_dev, grad_row_of, unique_grad_addr,
                       workspace, workspace_bytes, (cudaStream_t)stream_, 0);
}
// The part of demb_backward that does not need the gradients (pair list + sort).  Launch it EARLY — right after the prefetch, on a stream
// of the caller's — so it overlaps the forward gather instead of sitting in front of the gradient reduction.
int demb_backward_sort(int emb_dim, int64_t n, const int64_t* inverse, int64_t num_unique_bound, const int64_t* offsets, int64_t batch_size,
                       int num_features, int combiner, const int64_t* n_dev, const int64_t* grad_row_of, void* workspace, int64_t workspace_bytes,
                       void* stream_) {
  return backward_impl(nullptr, 0, emb_dim, n, inverse, num_unique_bound, nullptr, nullptr, 0, offsets, batch_size, num_features, combiner, 0, 0.f, 0.f, 0.f,
                       0.f, 0.f, 1.f, 1.f, nullptr, n_dev, grad_row_of, nullptr, workspace, workspace_bytes, (cudaStream_t)stream_, 1);
}
// demb_backward after demb_backward_sort(... same n / inverse / workspace ...); the caller orders it behind the sort
int demb_backward_apply(float* values, int64_t value_dim, int emb_dim, int64_t n, const int64_t* inverse, int64_t num_unique_bound,
                        const int64_t* rows, const float* grads, int64_t grad_stride, const int64_t* offsets, int64_t batch_size, int num_features,
                        int combiner, int opt_type, float lr, float eps, float beta1, float beta2, float weight_decay, float bias_correction1,
                        float bias_correction2, float* unique_grads, const int64_t* n_dev, const int64_t* unique_grad_addr, void* workspace,
                        int64_t workspace_bytes, void* stream_) {
  return backward_impl(values, value_dim, emb_dim, n, inverse, num_unique_bound, rows, grads, grad_stride, offsets, batch_size, num_features, combiner,
                       opt_type, lr, eps, beta1, beta2, weight_decay, bias_correction1, bias_correction2, unique_grads, n_dev, nullptr, unique_grad_addr,
                       workspace, workspace_bytes, (cudaStream_t)stream_, 2);
}

// development / measurement switches.  option 0 = fused lookup forward kernel: 2 (default) / 1 / 0, see g_probe_kernel
int demb_set_option(int option, int value) {
  if (option == 0) { g_probe_kernel = value; return 0; }
  if (option == 1) { g_bwd_tma = value != 0; return 0; }
  if (option >= 2 && option < 8) { g_dev_opts[option] = value; return 0; }
  return DEMB_ERR_ARG;
}

int demb_get_option(int option) { return (option >= 2 && option < 8) ? g_dev_opts[option] : -1; }

int demb_profile_enable(int on) {
  if (on && !g_prof_ev[0]) for (int i = 0; i < 4; ++i) if (cudaEventCreate(&g_prof_ev[i]) != cudaSuccess) return DEMB_ERR_ARG;
  g_prof_on = on != 0;
  return 0;
}
// stage times (ms) of the most recent demb_backward: [pairs + sort, tiles kernel, spans kernel]; synchronises on the last event
int demb_profile_read(float* ms3) {
  if (!g_prof_ev[0]) return DEMB_ERR_ARG;
  if (cudaEventSynchronize(g_prof_ev[3]) != cudaSuccess) return DEMB_ERR_ARG;
  for (int i = 0; i < 3; ++i) if (cudaEventElapsedTime(&ms3[i], g_prof_ev[i], g_prof_ev[i + 1]) != cudaSuccess) return DEMB_ERR_ARG;
  return 0;
}

// load_from_flat_table_{contiguous,emb,value} / store_to_flat_table_{contiguous,value} (dynamic_emb_op.cu:295-490): see FlatArgs above.
int demb_flat_table_copy(const int64_t* table_ptrs, const int64_t* table_ids, int64_t scalar_table_id, const int64_t* indices, int64_t n,
                         const int64_t* table_value_dims, const int64_t* table_emb_dims, int64_t max_emb_dim, float* dense, int64_t dense_stride,
                         int64_t dense_dim, int region, int to_table, void* stream) {
  if (region < 0 || region > 2 || !table_ptrs || !indices || !dense) return DEMB_ERR_ARG;
  if (n <= 0) return 0;
  FlatArgs a{table_ptrs, table_ids, scalar_table_id, indices, n, table_value_dims, table_emb_dims, max_emb_dim, dense, dense_stride, dense_dim, region, to_table};
  flat_table_copy_kernel<<<warp_grid(n), kBlock, 0, (cudaStream_t)stream>>>(a);
  DEMB_CHECK_LAST();
  return 0;
}
// {sgd,adam,adagrad,rowwise_adagrad}_update_for_flat_table (optimizer.cu:416-447): grads[n, max_emb_dim] applied to rows addressed through
// per-table base pointers; embedding dims must be multiples of 4 and <= 1024.
int demb_flat_table_update(const int64_t* table_ptrs, const int64_t* table_ids, const int64_t* indices, int64_t n, const int64_t* table_value_dims,
                           const int64_t* table_emb_dims, int64_t max_emb_dim, const float* grads, int64_t grad_stride, int opt_type, float lr,
                           float eps, float beta1, float beta2, float weight_decay, float bias_correction1, float bias_correction2, void* stream) {
  if (!table_ptrs || !indices || !grads || max_emb_dim <= 0 || (max_emb_dim & 3) || max_emb_dim > 128 * kMaxChunks) return DEMB_ERR_ARG;
  if (n <= 0) return 0;
  FlatArgs a{table_ptrs, table_ids, 0, indices, n, table_value_dims, table_emb_dims, max_emb_dim, nullptr, 0, 0, 0, 0};
  OptArgs o{opt_type, lr, eps, beta1, beta2, weight_decay, bias_correction1, bias_correction2};
  DISPATCH_NCHUNK((int)max_emb_dim, flat_table_update_kernel<NC><<<warp_grid(n), kBlock, 0, (cudaStream_t)stream>>>(a, grads, grad_stride, o));
  DEMB_CHECK_LAST();
  return 0;
}

int demb_update_rows(float* values, int64_t value_dim, int emb_dim, int64_t n, const int64_t* rows, const float* grads, int64_t grad_stride,
                     int opt_type, float lr, float eps, float beta1, float beta2, float weight_decay, float bias_correction1,
                     float bias_correction2, void* stream) {
  if (check_dims(emb_dim, value_dim)) return DEMB_ERR_ARG;
  if (n <= 0) return 0;
  OptArgs o{opt_type, lr, eps, beta1, beta2, weight_decay, bias_correction1, bias_correction2};
  DISPATCH_NCHUNK(emb_dim, update_rows_kernel<NC><<<warp_grid(n), kBlock, 0, (cudaStream_t)stream>>>(values, value_dim, emb_dim, n, rows, grads, grad_stride, o));
  DEMB_CHECK_LAST();
  return 0;
}

}  // extern "C"
