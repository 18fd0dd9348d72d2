// Row-wise sharded exchange over NVLink peer memory — replaces, for the DynamicEmb path, what the reference gets from TorchRec:
// RwSparseFeaturesDist (block_bucketize + KJTAllToAll of lengths and ids, corelib/dynamicemb/dynamicemb/input_dist.py:225-285),
// SequenceEmbeddingsAllToAll of the rows and its mirrored gradient exchange (shard/embedding.py:183-340, planner/rw_sharding.py:83),
// including the two host synchronisations of that path (split sizes, dedup count).
//
// B200 design: every rank owns one SYMMETRIC buffer (peer-mapped on all ranks of the NVSwitch domain); all counts stay on the device and
// travel with the data, so a whole training step is a fixed sequence of launches (CUDA-graph capturable):
//   requester   unique ids (per table) -> demb_shard_route: stable partition by owning rank; each id is STORED STRAIGHT INTO ITS OWNER'S
//               `ids_in[me]` segment over NVLink, with the per-table counts and the base offset of the rows the owner has to send back
//   barrier     demb_peer_barrier (flag exchange in the symmetric buffers, one 32-thread launch)
//   owner       demb_shard_recv_plan + _compact: received segments -> one table-major id list + for every id the peer ADDRESS its row
//               must be written to; local fused prefetch (dedup across sources, probe, insert/evict, init); demb_shard_gather_to_peers
//               copies each row from the value table directly into the requester's `rows_back` buffer (no staging, no all_to_all)
//   barrier
//   requester   ordinary gather from rows_back undoes partition + dedup (or pools: EmbeddingBagCollection semantics at the requester)
//   backward    the requester's gradient reduce (demb_backward, unique_grad_addr) writes every reduced row straight into the owner's
//               `grads_in[me]` segment; barrier; the owner's fused backward reads them through grad_row_of.
// Layout of a rank's symmetric buffer (byte offsets from demb_shard_layout): flags | meta_in[W] | ids_in[W][pair_cap] | rows_back[n_cap][D]
// | grads_in[W][pair_cap][D].
#include "../../include/dynamicemb_b200.h"
#include "demb_common.cuh"
#include "demb_scan.cuh"
#include "sm100_ptx.cuh"

#include <string.h>

using namespace demb;

namespace {
constexpr int kMaxW = 16;
constexpr int kMaxT = 60;
constexpr int kMetaWords = 64;                   // int64 words per (dst <- src) meta record: total, base, counts[kMaxT]
constexpr int kRouteTile = 1024;

struct Layout { int64_t flags, meta, ids_in, rows_back, grads_in, total; };
__host__ __device__ inline Layout make_layout(int W, int64_t pair_cap, int64_t n_cap, int D) {
  Layout L;
  int64_t o = 0;
  L.flags = o; o += 4096;
  L.meta = o; o += (int64_t)W * kMetaWords * 8;
  o = (o + 255) & ~(int64_t)255;
  L.ids_in = o; o += (int64_t)W * pair_cap * 8;
  o = (o + 255) & ~(int64_t)255;
  L.rows_back = o; o += n_cap * (int64_t)D * 4;
  o = (o + 255) & ~(int64_t)255;
  L.grads_in = o; o += (int64_t)W * pair_cap * D * 4;
  L.total = (o + 255) & ~(int64_t)255;
  return L;
}

struct Shard {
  int W, me, T, D;
  int64_t pair_cap, n_cap, recv_cap;
  const int64_t* peers;          // device [W]: base address of every rank's symmetric buffer as mapped HERE
  Layout L;
  int32_t* err;                  // device flag: 1 = a capacity was exceeded (ids dropped), 2 = barrier timeout
  __device__ __forceinline__ uint8_t* base(int r) const { return reinterpret_cast<uint8_t*>(peers[r]); }
};

// sparse_block_bucketize_features.cu:254-259 + :30-37: owning rank and the id the owner sees
// x % W for a small W with 32-bit arithmetic only (a 64-bit modulo by a run-time value is ~150 instructions): exact, since
// x = hi * 2^32 + lo  =>  x mod W = ((hi mod W) * (2^32 mod W) + lo mod W) mod W, every term < 2^32 for W <= 16
__device__ __forceinline__ int mod_small(uint64_t x, int W) {
  const uint32_t w = (uint32_t)W, hi = (uint32_t)(x >> 32), lo = (uint32_t)x;
  const uint32_t r32 = (uint32_t)(0x100000000ull % w);
  return (int)(((hi % w) * r32 + lo % w) % w);
}
__device__ __forceinline__ void owner_of(uint64_t key, int dist_type, int64_t blk, int W, int& p, uint64_t& nid) {
  if (dist_type == 1) { p = mod_small(key, W); nid = key; }
  else if (dist_type == 2) { p = mod_small(fmix64(key), W); nid = key; }
  else if (key < (uint64_t)blk * (uint64_t)W) { p = (int)(key / (uint64_t)blk); nid = key % (uint64_t)blk; }
  else { p = (int)(key % (uint64_t)W); nid = key / (uint64_t)W; }
}

struct RouteArgs {
  Shard s; int per_id_atomics;
  const uint64_t* ukeys; const int64_t* utids; const int64_t* n_u;   // unique ids grouped by table, device count
  int64_t n_max;
  const int32_t* dist_type; const int64_t* block_sizes;               // per table
  int32_t* tile_cnt;       // [tiles][W]
  int64_t* pair_table_cnt; // [W][T] (zeroed by the plan kernel of the PREVIOUS call / init)
  int64_t* dest_base;      // [W+1]
  int64_t* send_pos;       // [n_max] position of unique u in rows_back (-1 = dropped)
  int64_t* ug_addr;        // [n_max] address of unique u's gradient row in its owner's grads_in (0 = dropped)
};

__global__ void __launch_bounds__(kRouteTile) route_count_kernel(RouteArgs a) {
  __shared__ int cnt[kMaxW];
  __shared__ int pt[kMaxW * kMaxT];            // per-(owner, table) counts of this tile: one global atomic per non-zero entry, not one per id
  const int64_t n = *a.n_u < a.n_max ? *a.n_u : a.n_max;
  const int64_t u = (int64_t)blockIdx.x * kRouteTile + threadIdx.x;
  if (threadIdx.x < kMaxW) cnt[threadIdx.x] = 0;
  for (int i = threadIdx.x; i < a.s.W * a.s.T; i += kRouteTile) pt[i] = 0;
  __syncthreads();
  int p = -1;
  if (u < n) {
    const int t = a.utids ? (int)a.utids[u] : 0;
    uint64_t nid;
    owner_of(a.ukeys[u], a.dist_type[t], a.block_sizes[t], a.s.W, p, nid);
    if (a.per_id_atomics) atomicAdd(reinterpret_cast<unsigned long long*>(a.pair_table_cnt + p * a.s.T + t), 1ull);
    else if (a.s.T > 1) atomicAdd(&pt[p * a.s.T + t], 1);
  }
  for (int d = 0; d < a.s.W; ++d) {
    const unsigned m = __ballot_sync(0xffffffffu, p == d);
    if ((threadIdx.x & 31) == 0 && m) atomicAdd(&cnt[d], __popc(m));
  }
  __syncthreads();
  if (threadIdx.x < a.s.W) a.tile_cnt[(int64_t)blockIdx.x * a.s.W + threadIdx.x] = cnt[threadIdx.x];
  if (a.per_id_atomics) return;
  if (a.s.T == 1) { if (threadIdx.x < a.s.W && cnt[threadIdx.x]) atomicAdd(reinterpret_cast<unsigned long long*>(a.pair_table_cnt + threadIdx.x), (unsigned long long)cnt[threadIdx.x]); }
  else for (int i = threadIdx.x; i < a.s.W * a.s.T; i += kRouteTile) if (pt[i]) atomicAdd(reinterpret_cast<unsigned long long*>(a.pair_table_cnt + i), (unsigned long long)pt[i]);
}

// one block: exclusive scan of the tile counts per destination, capacities, meta records to the owners
__global__ void __launch_bounds__(1024) route_plan_kernel(RouteArgs a, int64_t n_tiles) {
  __shared__ int64_t tot[kMaxW + 1];
  const int W = a.s.W;
  for (int d = 0; d < W; ++d) {
    int64_t carry = 0;
    for (int64_t t0 = 0; t0 < n_tiles; t0 += blockDim.x) {
      const int64_t t = t0 + threadIdx.x;
      const int v = t < n_tiles ? a.tile_cnt[t * W + d] : 0;
      int total = 0;
      const int incl = block_inclusive_scan(v, total);
      if (t < n_tiles) a.tile_cnt[t * W + d] = (int32_t)(carry + incl - v);         // exclusive base of this tile for destination d
      carry += total;
      __syncthreads();
    }
    if (threadIdx.x == 0) tot[d] = carry;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int64_t b = 0;
    for (int d = 0; d < W; ++d) {
      a.dest_base[d] = b;
      int64_t c = tot[d];
      if (c > a.s.pair_cap) { c = a.s.pair_cap; atomicExch(a.s.err, 1); }
      b += tot[d];
    }
    a.dest_base[W] = b;
    if (b > a.s.n_cap) atomicExch(a.s.err, 1);
  }
  __syncthreads();
  // meta record for every owner d: peer d's meta_in[me] = {total, base of its rows in my rows_back, per-table counts}
  for (int i = threadIdx.x; i < W * kMetaWords; i += blockDim.x) {
    const int d = i / kMetaWords, w = i % kMetaWords;
    int64_t v = 0;
    if (w == 0) v = tot[d] < a.s.pair_cap ? tot[d] : a.s.pair_cap;
    else if (w == 1) v = a.dest_base[d];
    else if (w - 2 < a.s.T) v = a.pair_table_cnt[d * a.s.T + (w - 2)];
    int64_t* meta = reinterpret_cast<int64_t*>(a.s.base(d) + a.s.L.meta) + (int64_t)a.s.me * kMetaWords;
    meta[w] = v;
  }
}

__global__ void __launch_bounds__(kRouteTile) route_scatter_kernel(RouteArgs a) {
  __shared__ int wcnt[kRouteTile / 32][kMaxW];
  const int64_t n = *a.n_u < a.n_max ? *a.n_u : a.n_max;
  const int64_t u = (int64_t)blockIdx.x * kRouteTile + threadIdx.x;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int p = -1; uint64_t nid = 0;
  if (u < n) {
    const int t = a.utids ? (int)a.utids[u] : 0;
    owner_of(a.ukeys[u], a.dist_type[t], a.block_sizes[t], a.s.W, p, nid);
  }
  int rank_in_warp = 0;
  for (int d = 0; d < a.s.W; ++d) {
    const unsigned m = __ballot_sync(0xffffffffu, p == d);
    if (p == d) rank_in_warp = __popc(m & ((1u << lane) - 1u));
    if (lane == 0) wcnt[w][d] = __popc(m);
  }
  __syncthreads();
  if (u < n) {
    int before = 0;
    for (int ww = 0; ww < w; ++ww) before += wcnt[ww][p];
    const int64_t j = (int64_t)a.tile_cnt[(int64_t)blockIdx.x * a.s.W + p] + before + rank_in_warp;   // index inside my segment at owner p
    if (j < a.s.pair_cap && a.dest_base[p] + j < a.s.n_cap) {
      uint8_t* ob = a.s.base(p);
      reinterpret_cast<uint64_t*>(ob + a.s.L.ids_in)[(int64_t)a.s.me * a.s.pair_cap + j] = nid;          // NVLink store
      a.send_pos[u] = a.dest_base[p] + j;
      a.ug_addr[u] = (int64_t)reinterpret_cast<uintptr_t>(reinterpret_cast<float*>(ob + a.s.L.grads_in) + ((int64_t)a.s.me * a.s.pair_cap + j) * a.s.D);
    } else {
      a.send_pos[u] = -1;
      a.ug_addr[u] = 0;
    }
  }
  // leave the per-(owner, table) counters clean for the next call: the plan kernel has consumed them
  if (blockIdx.x == 0) for (int i = threadIdx.x; i < a.s.W * a.s.T; i += blockDim.x) a.pair_table_cnt[i] = 0;
}

// ---- owner side ------------------------------------------------------------------------------------------------------------------
struct RecvArgs {
  Shard s;
  int64_t* start;          // [T][W] first position (table-major list) of source s's ids of table t
  int64_t* cum;            // [W][T+1] per-source prefix over tables
  int64_t* table_range;    // [T+1]
  int64_t* n_recv;         // [1]
  int64_t* src_total;      // [W]
  uint64_t* ids_recv; int64_t* src_pos; int64_t* dst_addr;    // [recv_cap]
};

__global__ void recv_plan_kernel(RecvArgs a) {
  // W * T <= 16 * 60: one thread does the bookkeeping
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const int W = a.s.W, T = a.s.T;
  const int64_t* meta = reinterpret_cast<const int64_t*>(a.s.base(a.s.me) + a.s.L.meta);
  int64_t pos = 0;
  for (int s = 0; s < W; ++s) {
    a.src_total[s] = meta[s * kMetaWords];
    int64_t c = 0;
    for (int t = 0; t < T; ++t) { a.cum[s * (T + 1) + t] = c; c += meta[s * kMetaWords + 2 + t]; }
    a.cum[s * (T + 1) + T] = c;
  }
  for (int t = 0; t < T; ++t) {
    a.table_range[t] = pos;
    for (int s = 0; s < W; ++s) {
      a.start[t * W + s] = pos;
      // ids a source dropped for capacity never arrived: only the first src_total[s] of its list are present
      int64_t lo = a.cum[s * (T + 1) + t], hi = a.cum[s * (T + 1) + t + 1];
      if (hi > a.src_total[s]) hi = a.src_total[s];
      if (lo > hi) lo = hi;
      pos += hi - lo;
    }
  }
  a.table_range[T] = pos;
  if (pos > a.s.recv_cap) { atomicExch(a.s.err, 1); pos = a.s.recv_cap; }
  *a.n_recv = pos;
}

__global__ void recv_compact_kernel(RecvArgs a) {
  const int W = a.s.W, T = a.s.T;
  const uint8_t* mine = a.s.base(a.s.me);
  const uint64_t* ids_in = reinterpret_cast<const uint64_t*>(mine + a.s.L.ids_in);
  const int64_t* meta = reinterpret_cast<const int64_t*>(mine + a.s.L.meta);
  for (int s = 0; s < W; ++s) {
    const int64_t tot = a.src_total[s];
    const int64_t back_base = meta[s * kMetaWords + 1];
    float* rows_back = reinterpret_cast<float*>(a.s.base(s) + a.s.L.rows_back);
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < tot; j += (int64_t)gridDim.x * blockDim.x) {
      int lo = 0, hi = T;                                    // table of the j-th id of source s
      const int64_t* cum = a.cum + s * (T + 1);
      while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (cum[mid] <= j) lo = mid; else hi = mid; }
      const int64_t k = a.start[lo * W + s] + (j - cum[lo]);
      if (k >= a.s.recv_cap) continue;
      a.ids_recv[k] = ids_in[(int64_t)s * a.s.pair_cap + j];
      a.src_pos[k] = (int64_t)s * a.s.pair_cap + j;          // gradient row id inside grads_in
      a.dst_addr[k] = (int64_t)reinterpret_cast<uintptr_t>(rows_back + (back_base + j) * a.s.D);
    }
  }
}

// copy the row of every received id from the local value table straight into the requester's rows_back (peer stores over NVLink)
template <int U>
__global__ void __launch_bounds__(256) gather_to_peers_kernel(const float* __restrict__ values, int64_t vdim, int D, int64_t n_max, const int64_t* __restrict__ n_dev,
                                                              const int64_t* __restrict__ rows, const int64_t* __restrict__ inverse,
                                                              const int64_t* __restrict__ dst_addr, const int8_t* __restrict__ hit_flags, int part) {
  int64_t n = *n_dev; n = n < n_max ? n : n_max;
  const int lane = threadIdx.x & 31;
  const int D4 = D >> 2;
  const int64_t tiles = (n + 31) >> 5;
  const int64_t wstride = (int64_t)gridDim.x * 8;
  for (int64_t tile = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5); tile < tiles; tile += wstride) {
    const int64_t base = tile << 5;
    const int cnt = (int)((n - base) < 32 ? (n - base) : 32);
    int64_t row = -1, dst = 0;
    if (lane < cnt) {
      const int64_t u = inverse[base + lane];
      // part 1: only the ids whose key the prefetch's lookup stage found (this launch may run BEFORE insert / evict / init have finished,
      // so rows[] of the others is not read at all); part 2: only the others; part 0: all
      const bool take = !hit_flags || part == 0 || (hit_flags[u] != 0) == (part == 1);
      if (take) { row = rows[u]; dst = dst_addr[base + lane]; }
    }
    if (__ballot_sync(0xffffffffu, dst != 0) == 0u) continue;
    for (int j = 0; j < cnt; j += U) {
      int64_t r[U], d[U];
#pragma unroll
      for (int q = 0; q < U; ++q) { r[q] = __shfl_sync(0xffffffffu, row, (j + q) & 31); d[q] = __shfl_sync(0xffffffffu, dst, (j + q) & 31); }
      for (int c = lane; c < D4; c += 32) {
        float4 v[U];
#pragma unroll
        for (int q = 0; q < U; ++q) v[q] = (j + q < cnt && r[q] >= 0) ? ld_nc_f4(values + r[q] * vdim + 4 * c) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int q = 0; q < U; ++q) if (j + q < cnt && d[q]) st_f4(reinterpret_cast<float*>(d[q]) + 4 * c, v[q]);
      }
    }
  }
}

// The same copy with the rows STAGED THROUGH SHARED MEMORY by bulk async copies: 32 cp.async.bulk loads (value table -> stage), then 32
// cp.async.bulk stores (stage -> the requester's rows_back, peer memory): the copy engine keeps many more NVLink writes in flight than
// the register version's 8 rows per warp.  One persistent CTA per SM, 12 warps x 16 KB.
constexpr int kG2PWarps = 12;
__global__ void __launch_bounds__(kG2PWarps * 32, 1) gather_to_peers_tma_kernel(const float* __restrict__ values, int64_t vdim, int D, int64_t n_max,
                                                                             const int64_t* __restrict__ n_dev, const int64_t* __restrict__ rows,
                                                                             const int64_t* __restrict__ inverse, const int64_t* __restrict__ dst_addr,
                                                                             const int8_t* __restrict__ hit_flags, int part, int warps_per_block) {
  extern __shared__ __align__(128) uint8_t stage_raw[];
  __shared__ uint64_t bars[kG2PWarps];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  if (wib >= warps_per_block) return;
  int64_t n = *n_dev; n = n < n_max ? n : n_max;
  const uint32_t row_bytes = (uint32_t)D * 4u;
  uint8_t* buf = stage_raw + (size_t)wib * 32u * row_bytes;
  if (lane == 0) { sm100::mbar_init(&bars[wib], 1); sm100::fence_barrier_init(); }
  __syncwarp();
  uint32_t parity = 0;
  const int64_t tiles = (n + 31) >> 5;
  const int64_t wstride = (int64_t)gridDim.x * warps_per_block;
  for (int64_t tile = (int64_t)blockIdx.x * warps_per_block + wib; tile < tiles; tile += wstride) {
    const int64_t base = tile << 5;
    const int cnt = (int)((n - base) < 32 ? (n - base) : 32);
    int64_t row = -1, dst = 0;
    if (lane < cnt) {
      const int64_t u = inverse[base + lane];
      const bool take = !hit_flags || part == 0 || (hit_flags[u] != 0) == (part == 1);   // see gather_to_peers_kernel
      if (take) { row = rows[u]; dst = dst_addr[base + lane]; }
    }
    if (__ballot_sync(0xffffffffu, dst != 0) == 0u) continue;
    const unsigned found = __ballot_sync(0xffffffffu, row >= 0);
    if (lane == 0) {
      sm100::bulk_wait_read0();                                  // the previous tile's stores have finished reading the stage
      sm100::mbar_arrive_expect_tx(&bars[wib], (uint32_t)__popc(found) * row_bytes);
    }
    __syncwarp();
    if (row >= 0) {
      sm100::bulk_load(buf + (size_t)lane * row_bytes, values + row * vdim, row_bytes, &bars[wib]);
    } else if (lane < cnt && dst) {
      float4* d = reinterpret_cast<float4*>(buf + (size_t)lane * row_bytes);
      for (int c = 0; c < (D >> 2); ++c) d[c] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    sm100::mbar_wait(&bars[wib], parity);
    parity ^= 1;
    sm100::fence_proxy_async_smem();
    __syncwarp();
    // consecutive received ids of one source go to consecutive rows of that source's rows_back: the common case is ONE 16-KB store per tile
    const int64_t dst0 = __shfl_sync(0xffffffffu, dst, 0);
    const bool contig = dst0 != 0 && __all_sync(0xffffffffu, lane >= cnt || dst == dst0 + (int64_t)lane * row_bytes);
    if (contig) { if (lane == 0) sm100::bulk_store(reinterpret_cast<void*>(dst0), buf, (uint32_t)cnt * row_bytes); }
    else if (lane < cnt && dst) sm100::bulk_store(reinterpret_cast<void*>(dst), buf + (size_t)lane * row_bytes, row_bytes);
    sm100::bulk_commit();                                          // every lane commits its own group (bulk_wait_read0 above is lane 0's)
    // the other lanes must also have their stores drained before the stage is overwritten: wait per lane here (reads only)
    sm100::bulk_wait_read0();
    __syncwarp();
  }
  sm100::bulk_wait0();
}

// flag exchange barrier: every rank bumps its epoch, stores it into flag[channel][me] of every peer and waits until its own
// flag[channel][*] all carry that epoch.  Writes of earlier kernels of this stream (to peers) are ordered before the signal.
__global__ void peer_barrier_kernel(Shard s, int channel, unsigned long long* epochs) {
  const int lane = threadIdx.x;
  unsigned long long e = 0;
  if (lane == 0) e = epochs[channel] + 1;
  e = __shfl_sync(0xffffffffu, e, 0);
  __threadfence_system();
  if (lane < s.W) {
    volatile unsigned long long* theirs = reinterpret_cast<volatile unsigned long long*>(s.base(lane) + s.L.flags) + channel * kMaxW + s.me;
    *theirs = e;
    __threadfence_system();
    volatile unsigned long long* mine = reinterpret_cast<volatile unsigned long long*>(s.base(s.me) + s.L.flags) + channel * kMaxW + lane;
    const long long t0 = clock64();
    while (*mine < e) {
      if (clock64() - t0 > (1ll << 35)) { atomicExch(s.err, 2); break; }       // ~17 s at 2 GHz: a peer is gone; do not hang the GPU
    }
  }
  __syncwarp();
  __threadfence_system();
  if (lane == 0) epochs[channel] = e;
}

__global__ void zero_i64_kernel(int64_t* p, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] = 0;
}

int make_shard(Shard& s, int world, int rank, int num_tables, int emb_dim, int64_t pair_cap, int64_t n_cap, int64_t recv_cap, const int64_t* peers, int32_t* err) {
  if (world < 1 || world > kMaxW || rank < 0 || rank >= world || num_tables < 1 || num_tables > kMaxT || emb_dim <= 0 || (emb_dim & 3)) return DEMB_ERR_ARG;
  s.W = world; s.me = rank; s.T = num_tables; s.D = emb_dim; s.pair_cap = pair_cap; s.n_cap = n_cap; s.recv_cap = recv_cap; s.peers = peers; s.err = err;
  s.L = make_layout(world, pair_cap, n_cap, emb_dim);
  return 0;
}
size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }
}  // namespace

extern "C" {

// byte offsets of the regions of a rank's symmetric buffer: out[6] = flags, meta, ids_in, rows_back, grads_in, total size
int demb_shard_layout(int world, int64_t pair_cap, int64_t n_cap, int emb_dim, int64_t* out) {
  if (world < 1 || world > kMaxW || pair_cap <= 0 || n_cap <= 0 || emb_dim <= 0) return DEMB_ERR_ARG;
  const Layout L = make_layout(world, pair_cap, n_cap, emb_dim);
  out[0] = L.flags; out[1] = L.meta; out[2] = L.ids_in; out[3] = L.rows_back; out[4] = L.grads_in; out[5] = L.total;
  return 0;
}

int64_t demb_shard_route_workspace_bytes(int64_t n_max, int world, int num_tables) {
  const int64_t tiles = (n_max + kRouteTile - 1) / kRouteTile;
  return (int64_t)(align256(4 * (size_t)(tiles * world)) + align256(8 * (size_t)(world * num_tables)) + align256(8 * (size_t)(world + 1)) + 256);
}

// Requester: route the unique ids (grouped by table; count on the device) to their owners.  `state` = persistent device memory of
// demb_shard_route_workspace_bytes, zeroed once by the caller.  Outputs send_pos[n_max], unique_grad_addr[n_max] (see file header).
int demb_shard_route(int world, int rank, int num_tables, int emb_dim, int64_t pair_cap, int64_t n_cap, const int64_t* peers, int32_t* err,
                     int64_t n_max, const int64_t* n_unique_dev, const void* unique_keys, const int64_t* unique_table_ids,
                     const int32_t* dist_type_per_table, const int64_t* block_size_per_table, int64_t* send_pos, int64_t* unique_grad_addr,
                     void* state, int64_t state_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  Shard s;
  if (int rc = make_shard(s, world, rank, num_tables, emb_dim, pair_cap, n_cap, 0, peers, err)) return rc;
  if (n_max <= 0) return DEMB_ERR_ARG;
  if (state_bytes < demb_shard_route_workspace_bytes(n_max, world, num_tables)) return DEMB_ERR_WORKSPACE;
  const int64_t tiles = (n_max + kRouteTile - 1) / kRouteTile;
  uint8_t* w = (uint8_t*)state;
  RouteArgs a;
  a.s = s; a.per_id_atomics = demb_get_option(2) == 0; a.ukeys = (const uint64_t*)unique_keys; a.utids = num_tables > 1 ? unique_table_ids : nullptr; a.n_u = n_unique_dev; a.n_max = n_max;
  a.dist_type = dist_type_per_table; a.block_sizes = block_size_per_table;
  a.tile_cnt = (int32_t*)w; w += align256(4 * (size_t)(tiles * world));
  a.pair_table_cnt = (int64_t*)w; w += align256(8 * (size_t)(world * num_tables));
  a.dest_base = (int64_t*)w;
  a.send_pos = send_pos; a.ug_addr = unique_grad_addr;
  route_count_kernel<<<(int)tiles, kRouteTile, 0, stream>>>(a);
  route_plan_kernel<<<1, 1024, 0, stream>>>(a, tiles);
  route_scatter_kernel<<<(int)tiles, kRouteTile, 0, stream>>>(a);
  DEMB_CHECK_LAST();
  return 0;
}

int64_t demb_shard_recv_workspace_bytes(int world, int num_tables) {
  return (int64_t)(align256(8 * (size_t)(world * num_tables)) + align256(8 * (size_t)(world * (num_tables + 1))) + align256(8 * (size_t)world) + 256);
}

// Owner (after the barrier): received segments -> table-major id list ids_recv[recv_cap] (+ table_range[T+1], n_recv on the device),
// src_pos[k] = gradient row id of id k inside grads_in, dst_addr[k] = peer address its row has to be written to.
int demb_shard_recv(int world, int rank, int num_tables, int emb_dim, int64_t pair_cap, int64_t n_cap, int64_t recv_cap, const int64_t* peers,
                    int32_t* err, void* ids_recv, int64_t* table_range, int64_t* n_recv, int64_t* src_pos, int64_t* dst_addr, void* workspace,
                    int64_t workspace_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  Shard s;
  if (int rc = make_shard(s, world, rank, num_tables, emb_dim, pair_cap, n_cap, recv_cap, peers, err)) return rc;
  if (workspace_bytes < demb_shard_recv_workspace_bytes(world, num_tables)) return DEMB_ERR_WORKSPACE;
  uint8_t* w = (uint8_t*)workspace;
  RecvArgs a;
  a.s = s;
  a.start = (int64_t*)w; w += align256(8 * (size_t)(world * num_tables));
  a.cum = (int64_t*)w; w += align256(8 * (size_t)(world * (num_tables + 1)));
  a.src_total = (int64_t*)w;
  a.table_range = table_range; a.n_recv = n_recv; a.ids_recv = (uint64_t*)ids_recv; a.src_pos = src_pos; a.dst_addr = dst_addr;
  recv_plan_kernel<<<1, 32, 0, stream>>>(a);
  int64_t blocks = (recv_cap + 255) / 256;
  const int64_t cap = (int64_t)sm_count() * 8;
  recv_compact_kernel<<<(int)(blocks < 1 ? 1 : (blocks > cap ? cap : blocks)), 256, 0, stream>>>(a);
  DEMB_CHECK_LAST();
  return 0;
}

// Owner: out row of received id k = values[rows[inverse[k]]] (zeros when absent), written to dst_addr[k] (peer memory).
// hit_flags (nullable, per UNIQUE id, from demb_train_prefetch_hook) with part 1 / 2 splits the copy: part 1 = the ids whose key the
// prefetch's lookup stage found — may be launched on another stream as soon as that stage is done, concurrently with insert / evict /
// row init — and part 2 = the rest, after the prefetch.  part 0 (or no flags) copies everything.
static int gather_to_peers_impl(const float* values, int64_t value_dim, int emb_dim, int64_t n_max, const int64_t* n_dev, const int64_t* rows,
                                const int64_t* inverse, const int64_t* dst_addr, const int8_t* hit_flags, int part, void* stream) {
  if (emb_dim <= 0 || (emb_dim & 3) || (value_dim & 3) || value_dim < emb_dim || !n_dev || part < 0 || part > 2) return DEMB_ERR_ARG;
  if (n_max <= 0) return 0;
  int64_t blocks = ((n_max + 31) / 32 + 7) / 8;
  // part 1 runs beside the owner's insert / evict / init kernels: one CTA per SM (option 7) fits into the registers those leave free
  const int64_t cap = (int64_t)sm_count() * ((part == 1 && demb_get_option(7) > 0) ? demb_get_option(7) : 8);
  const size_t stage = 32u * (size_t)emb_dim * 4u;
  int tw = (int)((200u * 1024u) / stage); if (tw > kG2PWarps) tw = kG2PWarps;
  if (demb_get_option(5) != 0 && tw >= 2 && part != 1) {
    static std::atomic<int> configured[kMaxDevices];
    cudaError_t ce = once_per_device(configured, [] { return cudaFuncSetAttribute(gather_to_peers_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); });
    if (ce != cudaSuccess) return -(int)ce;
    int64_t tb = ((n_max + 31) / 32 + tw - 1) / tw;
    if (tb > sm_count()) tb = sm_count();
    gather_to_peers_tma_kernel<<<(int)tb, kG2PWarps * 32, (size_t)tw * stage, (cudaStream_t)stream>>>(values, value_dim, emb_dim, n_max, n_dev, rows, inverse, dst_addr,
                                                                                                      hit_flags, part, tw);
  } else {
    gather_to_peers_kernel<8><<<(int)(blocks > cap ? cap : blocks), 256, 0, (cudaStream_t)stream>>>(values, value_dim, emb_dim, n_max, n_dev, rows, inverse, dst_addr,
                                                                                                     hit_flags, part);
  }
  DEMB_CHECK_LAST();
  return 0;
}
int demb_shard_gather_to_peers(const float* values, int64_t value_dim, int emb_dim, int64_t n_max, const int64_t* n_dev, const int64_t* rows,
                               const int64_t* inverse, const int64_t* dst_addr, void* stream) {
  return gather_to_peers_impl(values, value_dim, emb_dim, n_max, n_dev, rows, inverse, dst_addr, nullptr, 0, stream);
}
int demb_shard_gather_to_peers_part(const float* values, int64_t value_dim, int emb_dim, int64_t n_max, const int64_t* n_dev, const int64_t* rows,
                                    const int64_t* inverse, const int64_t* dst_addr, const int8_t* hit_flags, int part, void* stream) {
  return gather_to_peers_impl(values, value_dim, emb_dim, n_max, n_dev, rows, inverse, dst_addr, hit_flags, part, stream);
}

// All ranks call this the same number of times per channel (0..3), in the same order.  epochs: 4 uint64 of LOCAL device memory, zero at start.
int demb_peer_barrier(int world, int rank, int64_t pair_cap, int64_t n_cap, int emb_dim, const int64_t* peers, int32_t* err, int channel,
                      uint64_t* epochs, void* stream) {
  Shard s;
  if (int rc = make_shard(s, world, rank, 1, emb_dim, pair_cap, n_cap, 0, peers, err)) return rc;
  if (channel < 0 || channel >= 4) return DEMB_ERR_ARG;
  peer_barrier_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(s, channel, (unsigned long long*)epochs);
  DEMB_CHECK_LAST();
  return 0;
}

// ---- one-time setup of the symmetric buffer (not on the step): cudaMalloc + legacy CUDA IPC, the mechanism NCCL's own P2P transport
// uses inside one box.  handle = 64 bytes (cudaIpcMemHandle_t) to be exchanged by the host side (torch.distributed all_gather).
int demb_ipc_alloc(int64_t bytes, void** ptr, void* handle64) {
  if (bytes <= 0 || !ptr || !handle64) return DEMB_ERR_ARG;
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
  cudaError_t e = cudaMalloc(ptr, (size_t)bytes);
  if (e != cudaSuccess) return -(int)e;
  e = cudaMemset(*ptr, 0, (size_t)bytes);
  if (e == cudaSuccess) e = cudaIpcGetMemHandle(reinterpret_cast<cudaIpcMemHandle_t*>(handle64), *ptr);
  if (e != cudaSuccess) { cudaFree(*ptr); *ptr = nullptr; return -(int)e; }
  return 0;
}
int demb_ipc_open(const void* handle64, void** ptr) {
  if (!ptr || !handle64) return DEMB_ERR_ARG;
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, sizeof(h));
  cudaError_t e = cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess);
  return e == cudaSuccess ? 0 : -(int)e;
}
int demb_ipc_close(void* ptr) { return ptr ? -(int)cudaIpcCloseMemHandle(ptr) : 0; }
int demb_ipc_free(void* ptr) { return ptr ? -(int)cudaFree(ptr) : 0; }

int demb_zero_i64(int64_t* p, int64_t n, void* stream) {
  if (n <= 0) return 0;
  int64_t blocks = (n + 255) / 256;
  zero_i64_kernel<<<(int)(blocks > 4096 ? 4096 : blocks), 256, 0, (cudaStream_t)stream>>>(p, n);
  DEMB_CHECK_LAST();
  return 0;
}

}  // extern "C"
