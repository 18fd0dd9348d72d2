// Fused training prefetch for the HBM-direct tier: dedup -> probe -> insert + init of the missing keys -> pin, with NO host
// synchronisation and no sort.  Replaces the reference's dynamicemb_prefetch / _prefetch_hbm_direct_path
// (corelib/dynamicemb/dynamicemb/batched_dynamicemb_function.py:559-830: segmented_unique (+host sync), table_lookup,
// flagged_compact (+host sync), initializer, table_insert (+unlock), store_to_flat, increment_counter x2, expand_table_ids).
//
//   1. demb_segmented_unique                      unique keys in first-occurrence order, count stays on the device
//   2. train_lookup_kernel (thread per unique)    probe; hit: score update + pin + slot/row out
//                                                 miss: push the key on its bucket's list (atomicExch on heads[bucket]); the first
//                                                 pusher records the bucket in `touched`
//   3. train_insert_thread_kernel (THREAD per touched bucket, device-side count)
//                                                 a bucket that cannot overflow (size + new keys <= capacity — all of them until the table
//                                                 fills up): walks the bucket's list in key order (the reference's deterministic order:
//                                                 (bucket, key), scored_hashtable.py:1451-1557) and inserts with thread_insert_one; marks the
//                                                 new keys.  ncu on the warp-per-bucket form: 500 warp instructions per inserted key with one
//                                                 useful lane in most of them (116 M warp instructions per step, issue-bound at 211 us).
//      train_insert_kernel (warp per bucket)      only the buckets that may have to evict (cooperative 128-slot victim scan); inserts,
//                                                 initialises and pins inline as before.
//   4. train_init_rows_kernel (warp per new key)  [emb | optimizer state] of the rows inserted by the thread kernel: pure streaming writes.
// The resulting table image is identical to lookup + demb_table_insert + demb_init_rows run op by op (tests/test_demb_module_gpu.py::test_fused_prefetch_matches_op_by_op; against the CPU oracle: tests/test_train_oracle_gpu.py).
#include "../../include/dynamicemb_b200.h"
#include "demb_common.cuh"
#include "demb_insert.cuh"
#include "demb_init.cuh"
#include "demb_probe.cuh"

using namespace demb;

namespace {
constexpr int kBlock = 256;

struct TrainArgs {
  Table t;
  int32_t* bucket_sizes; int32_t* counter; int32_t* heads;
  float* values; int64_t vdim; int D; const int64_t* row_base;
  const uint64_t* ukeys; const int64_t* utids; const int64_t* n_u;
  int pol; const uint64_t* table_scores; const int64_t* freq; uint64_t ts; int key_is_signed;
  int full_shortcut; InitArgs init; const InitArgs* table_init; float state_init;   // table_init: per-table initializer (device, [T]), nullable => init
  int64_t* slots; int64_t* rows; int32_t* next; int32_t* touched; unsigned long long* n_touched;
  int32_t* init_list; unsigned long long* n_init;   // uniques inserted by the thread kernel: their rows are initialised by train_init_rows_kernel
  int8_t* hit_flags;                                // nullable: 1 = found by the lookup stage (its row is valid and pinned from then on), 0 = not
};

__device__ __forceinline__ uint64_t train_score(const TrainArgs& a, int64_t u, int64_t tid) {
  if (a.pol == kConst) return 0;
  if (a.pol == kGlobalTimer) return a.ts ? a.ts : globaltimer();
  if (a.pol == kAccumulate || a.pol == kLruLfu) return a.freq ? (uint64_t)a.freq[u] : 1;
  return a.table_scores ? a.table_scores[tid] : 0;
}

__global__ void train_lookup_kernel(TrainArgs a) {
  const int64_t n = *a.n_u;
  const int lane = threadIdx.x & 31;
  // warp-uniform trip count (lanes past the end idle inside): the `touched` append below is warp-aggregated — one atomicAdd on the shared
  // counter per warp instead of one per bucket (~150 K same-address atomics per step serialised this kernel at ~100 us)
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t u0 = (int64_t)blockIdx.x * blockDim.x + (threadIdx.x & ~31); u0 < n; u0 += stride) {
    const int64_t u = u0 + lane;
    int32_t first_bucket = -1;                                                    // >= 0: this lane pushed the FIRST key of that bucket
    if (u < n) {
      const uint64_t key = a.ukeys[u];
      const int64_t tid = a.utids ? a.utids[u] : 0;
      Locus L = locate(a.t, key, tid);
      int64_t slot = -1, row = -1;
      if (L.cap > 0) {
        uint8_t* bk = a.t.bucket(L.bucket);
        const int64_t it = probe_thread(a.t, bk, key, L.h, nullptr);
        if (it >= 0) {
          if (a.pol != kConst) policy_update(a.pol, a.t.scores(bk, it), train_score(a, u, tid), a.ts, false);
          slot = (L.bucket - L.bkt_begin) * a.t.C + it;
          row = (a.row_base ? a.row_base[tid] : 0) + slot;
          atomicAdd(a.counter + L.bucket * a.t.C + it, 1);                        // pin (increment_counter, :607)
        } else {
          const int old = atomicExch(a.heads + L.bucket, (int)u);
          a.next[u] = old;                                                        // >= -1: list link
          if (old == -1) first_bucket = (int32_t)L.bucket;
        }
      }
      a.slots[u] = slot;
      a.rows[u] = row;
      if (a.hit_flags) a.hit_flags[u] = slot >= 0;
      if (slot >= 0 || L.cap <= 0) a.next[u] = -3;                                // not on any list (train_init_rows_kernel reads next[u] of every u)
    }
    const unsigned m = __ballot_sync(0xffffffffu, first_bucket >= 0);
    if (m) {
      const int leader = __ffs(m) - 1;
      unsigned long long base = 0;
      if (lane == leader) base = atomicAdd(a.n_touched, (unsigned long long)__popc(m));
      base = __shfl_sync(0xffffffffu, base, leader);
      if (first_bucket >= 0) a.touched[base + __popc(m & ((1u << lane) - 1u))] = first_bucket;
    }
  }
}

// The same lookup for 128-slot buckets on the warp-tile probe (demb_probe.cuh): two memory hops per tile (the coalesced digest lines,
// then all candidate key loads at once) instead of the per-thread chain of up to ten.
constexpr int kLookupWarps = 8;
__global__ void __launch_bounds__(kLookupWarps * 32, 3) train_lookup_tile_kernel(TrainArgs a) {
  __shared__ int slot_sm[kLookupWarps][32];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int64_t n = *a.n_u;
  const int64_t tiles = (n + 31) >> 5;
  const int64_t wstride = (int64_t)gridDim.x * kLookupWarps;
  int64_t tile = (int64_t)blockIdx.x * kLookupWarps + wib;
  TableCache tc = empty_table_cache();
  auto load_key = [&](int64_t tl) -> ProbeKey {
    const int64_t u = (tl << 5) + lane;
    if (tl >= tiles || u >= n) return ProbeKey{0, 0, 0, 0, false};
    return make_probe_key(a.t, a.ukeys[u], a.utids ? (int)a.utids[u] : 0, tc);
  };
  // no register double-buffering of the digest lines here: 24 warps per SM (3 CTAs) with one tile each in flight hide the latency, and
  // the 32 registers a second DigRegs costs would drop the kernel to 2 CTAs per SM (ncu: 20 % warps active, 85 us)
  for (; tile < tiles; tile += wstride) {
    const ProbeKey k0 = load_key(tile);
    DigRegs d0;
    tile_load_digests(a.t, k0, d0, lane);
    const int pos = tile_probe(a.t, k0, d0, slot_sm[wib], lane);
    const int64_t u = (tile << 5) + lane;
    int32_t first_bucket = -1;                                                    // >= 0: this lane pushed the FIRST key of that bucket
    if (u < n) {
      int64_t slot = -1, row = -1;
      if (k0.valid) {
        if (pos >= 0) {
          uint8_t* bk = a.t.bucket(k0.bucket);
          if (a.pol != kConst) policy_update(a.pol, a.t.scores(bk, pos), train_score(a, u, k0.tid), a.ts, false);
          slot = k0.slot_base + pos;
          row = (a.row_base ? a.row_base[k0.tid] : 0) + slot;
          atomicAdd(a.counter + k0.bucket * a.t.C + pos, 1);                      // pin (increment_counter, :607)
        } else {
          const int old = atomicExch(a.heads + k0.bucket, (int)u);
          a.next[u] = old;                                                        // >= -1: list link
          if (old == -1) first_bucket = (int32_t)k0.bucket;
        }
      }
      a.slots[u] = slot;
      a.rows[u] = row;
      if (a.hit_flags) a.hit_flags[u] = slot >= 0;
      if (slot >= 0 || !k0.valid) a.next[u] = -3;                                 // not on any list
    }
    const unsigned m = __ballot_sync(0xffffffffu, first_bucket >= 0);
    if (m) {
      const int leader = __ffs(m) - 1;
      unsigned long long base = 0;
      if (lane == leader) base = atomicAdd(a.n_touched, (unsigned long long)__popc(m));
      base = __shfl_sync(0xffffffffu, base, leader);
      if (first_bucket >= 0) a.touched[base + __popc(m & ((1u << lane) - 1u))] = first_bucket;
    }
  }
}

__device__ __forceinline__ bool key_less(uint64_t x, uint64_t y, int is_signed) { return is_signed ? ((int64_t)x < (int64_t)y) : (x < y); }

__device__ __forceinline__ void train_insert_key(const TrainArgs& a, int64_t b, int32_t u, uint64_t key, int lane) {
  const int64_t tid = a.utids ? a.utids[u] : 0;
  const int64_t bb = a.t.bkt_off[tid];
  const InsertOutcome o = warp_insert_one(a.t, b, key, train_score(a, u, tid), a.pol, a.ts, a.bucket_sizes, a.counter, lane);
  if (o.result <= kEvict) {
    const int64_t slot = (b - bb) * a.t.C + o.it;
    const int64_t row = (a.row_base ? a.row_base[tid] : 0) + slot;
    const int D4 = a.D >> 2, V4 = (int)(a.vdim >> 2);
    if (o.result != kAssignHit) {                                                   // new row: initializer + optimizer state (fused A10 + A11)
      const InitArgs ia = a.table_init ? a.table_init[tid] : a.init;
      for (int c = lane; c < D4; c += 32) st_f4(a.values + row * a.vdim + 4 * c, init4(ia, key, c));
      for (int c = D4 + lane; c < V4; c += 32) st_f4(a.values + row * a.vdim + 4 * c, make_float4(a.state_init, a.state_init, a.state_init, a.state_init));
    }
    if (lane == 0) { a.slots[u] = slot; a.rows[u] = row; atomicAdd(a.counter + b * a.t.C + o.it, 1); }
  }
  __threadfence_block();
  __syncwarp();
}

__global__ void __launch_bounds__(kBlock) train_insert_kernel(TrainArgs a) {
  const int lane = threadIdx.x & 31;
  const int64_t nt = (int64_t)*a.n_touched;
  const int64_t wstride = (int64_t)gridDim.x * (kBlock / 32);
  for (int64_t w = (int64_t)blockIdx.x * (kBlock / 32) + (threadIdx.x >> 5); w < nt; w += wstride) {
    const int64_t b = a.touched[w];
    const int head = *reinterpret_cast<volatile int*>(a.heads + b);
    // pop the list: lane l keeps the l-th element (lists are 1-3 long in practice; > 32 takes the selection path below)
    int cnt = 0, mine = -1;
    for (int cur = head; cur != -1; cur = a.next[cur]) { if (cnt == lane) mine = cur; ++cnt; }
    if (cnt <= 32) {
      const uint64_t mykey = mine >= 0 ? a.ukeys[mine] : 0;
      int rank = 0;
      for (int l = 0; l < cnt; ++l) {
        const uint64_t ok = __shfl_sync(0xffffffffu, mykey, l);
        if (mine >= 0 && l != lane && key_less(ok, mykey, a.key_is_signed)) ++rank;   // keys are unique => strict order
      }
      if (mine < 0) rank = -1;
      for (int r = 0; r < cnt; ++r) {
        const unsigned m = __ballot_sync(0xffffffffu, rank == r);
        const int src = __ffs(m) - 1;
        train_insert_key(a, b, __shfl_sync(0xffffffffu, mine, src), __shfl_sync(0xffffffffu, mykey, src), lane);
      }
    } else {
      // long list (tiny tables / adversarial batches): repeated selection of the smallest key greater than the last inserted one
      bool have_last = false; uint64_t last = 0;
      for (int r = 0; r < cnt; ++r) {
        uint64_t best = 0; int bu = -1; int idx = 0;
        for (int cur = head; cur != -1; cur = a.next[cur], ++idx) {
          if ((idx & 31) != lane) continue;
          const uint64_t k = a.ukeys[cur];
          if (have_last && !key_less(last, k, a.key_is_signed)) continue;
          if (bu < 0 || key_less(k, best, a.key_is_signed)) { best = k; bu = cur; }
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
          const uint64_t ok = __shfl_xor_sync(0xffffffffu, best, d);
          const int ou = __shfl_xor_sync(0xffffffffu, bu, d);
          if (ou >= 0 && (bu < 0 || key_less(ok, best, a.key_is_signed))) { best = ok; bu = ou; }
        }
        train_insert_key(a, b, bu, best, lane);
        last = best; have_last = true;
      }
    }
    if (lane == 0) a.heads[b] = -1;                                               // leave the list heads clean for the next step
  }
}

// Warp per touched bucket that may evict, 128-slot buckets: the bucket's keys, reduction scores and pin counters are loaded ONCE into
// registers (4 slots per lane, all loads independent), then every new key of the bucket (in key order) picks its slot from the registers —
// first Empty slot in probe order, else the minimum-score unpinned slot, first minimum in storage order (types.cuh:417-465,
// kernels.cuh:238-275) — and only the owning lane touches memory.  The per-key path above re-reads digests, scores, keys and counters of
// the whole bucket for every key (7 dependent round trips per insert; 0.25 ms of a 1.0 ms step at eviction steady state, where EVERY new
// key evicts).  A key on a bucket list is known to be absent (train_lookup), so there is no existence probe.
__global__ void __launch_bounds__(kBlock) train_evict_kernel(TrainArgs a) {
  const int lane = threadIdx.x & 31;
  const int64_t nt = (int64_t)*a.n_touched;
  const int64_t wstride = (int64_t)gridDim.x * (kBlock / 32);
  constexpr int C = kProbeC;
  int64_t w = (int64_t)blockIdx.x * (kBlock / 32) + (threadIdx.x >> 5);
  // software pipeline over this warp's buckets: the (bucket id -> list head) chain of the NEXT bucket is resolved, and its 2.6 KB of
  // keys / digests / scores / pin counters pulled into L2, while the current bucket is processed (the kernel is latency-bound: one warp
  // owns a bucket and walks a chain of dependent loads; ~24 warps per SM)
  int64_t b = w < nt ? a.touched[w] : -1;
  int head = b >= 0 ? *reinterpret_cast<volatile int*>(a.heads + b) : -1;
  bool prev_full = true;                                        // key lines of the next bucket are prefetched only while buckets are not full
  for (; w < nt; w += wstride) {
    const int64_t wn = w + wstride;
    const int64_t bn = wn < nt ? a.touched[wn] : -1;
    int head_n = -1;
    if (bn >= 0) {
      head_n = *reinterpret_cast<volatile int*>(a.heads + bn);
      const char* nb = reinterpret_cast<const char*>(a.t.bucket(bn));
      const int64_t bytes = a.t.bucket_bytes();
      if ((int64_t)lane * 128 < bytes && (!prev_full || lane * 128 >= 8 * C)) asm volatile("prefetch.global.L2 [%0];" ::"l"(nb + lane * 128));
      if (lane < 4) asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char*>(a.counter + bn * C) + lane * 128));
    }
    uint8_t* bk = a.t.bucket(b);
    // bucket state -> registers (issued before the list walk: independent of it)
    // A FULL bucket (bucket_sizes == C: the steady state) holds no Empty / Reclaim slot, so its keys are not needed to pick a victim:
    // 1 KB less to read per bucket and no first-empty search
    uint64_t kreg[4], sreg[4]; int32_t creg[4];
    const bool full = a.full_shortcut && a.bucket_sizes[b] >= C;
    prev_full = full;
    {
      kreg[0] = kreg[1] = kreg[2] = kreg[3] = 0;
      if (!full) {
        const ulonglong2 k01 = *reinterpret_cast<const ulonglong2*>(a.t.keys(bk) + lane * 4);
        const ulonglong2 k23 = *reinterpret_cast<const ulonglong2*>(a.t.keys(bk) + lane * 4 + 2);
        kreg[0] = k01.x; kreg[1] = k01.y; kreg[2] = k23.x; kreg[3] = k23.y;
      }
      const int4 c4 = *reinterpret_cast<const int4*>(a.counter + b * C + lane * 4);
      creg[0] = c4.x; creg[1] = c4.y; creg[2] = c4.z; creg[3] = c4.w;
#pragma unroll
      for (int q = 0; q < 4; ++q) sreg[q] = a.t.scores(bk, lane * 4 + q)[a.t.ns - 1];
    }
    int cnt = 0, mine = -1;
    for (int cur = head; cur != -1; cur = a.next[cur]) { if (cnt == lane) mine = cur; ++cnt; }
    if (cnt > 32) {                                                               // long list (tiny tables): per-key path, selection order
      bool have_last = false; uint64_t last = 0;
      for (int r = 0; r < cnt; ++r) {
        uint64_t best = 0; int bu = -1; int idx = 0;
        for (int cur = head; cur != -1; cur = a.next[cur], ++idx) {
          if ((idx & 31) != lane) continue;
          const uint64_t k = a.ukeys[cur];
          if (have_last && !key_less(last, k, a.key_is_signed)) continue;
          if (bu < 0 || key_less(k, best, a.key_is_signed)) { best = k; bu = cur; }
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
          const uint64_t ok = __shfl_xor_sync(0xffffffffu, best, d);
          const int ou = __shfl_xor_sync(0xffffffffu, bu, d);
          if (ou >= 0 && (bu < 0 || key_less(ok, best, a.key_is_signed))) { best = ok; bu = ou; }
        }
        train_insert_key(a, b, bu, best, lane);
        last = best; have_last = true;
      }
      if (lane == 0) a.heads[b] = -1;
      b = bn; head = head_n;
      continue;
    }
    const uint64_t mykey = mine >= 0 ? a.ukeys[mine] : 0;
    int rank = 0;
    for (int l = 0; l < cnt; ++l) {
      const uint64_t ok = __shfl_sync(0xffffffffu, mykey, l);
      if (mine >= 0 && l != lane && key_less(ok, mykey, a.key_is_signed)) ++rank;   // keys are unique => strict order
    }
    if (mine < 0) rank = -1;
    for (int r = 0; r < cnt; ++r) {
      const unsigned sel = __ballot_sync(0xffffffffu, rank == r);
      const int src = __ffs(sel) - 1;
      const int u = __shfl_sync(0xffffffffu, mine, src);
      const uint64_t key = __shfl_sync(0xffffffffu, mykey, src);
      const int64_t tid = a.utids ? a.utids[u] : 0;
      const int64_t h = hash63(key);
      const int start = (int)(h % C) & ~15;
      // first Empty slot in probe order: smallest (pos - start) mod C
      int dmin = 1 << 20;
      if (!full) {
#pragma unroll
        for (int q = 0; q < 4; ++q) if (kreg[q] == kEmptyKey) { const int d = ((lane * 4 + q) - start) & (C - 1); dmin = d < dmin ? d : dmin; }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) { const int o = __shfl_xor_sync(0xffffffffu, dmin, d); dmin = o < dmin ? o : dmin; }
      }
      int pos = -1; int result = kBusy;
      if (dmin < (1 << 20)) { pos = (start + dmin) & (C - 1); result = kInsert; }
      else {
        uint64_t best = 0xFFFFFFFFFFFFFFFFull; int bi = -1;
#pragma unroll
        for (int q = 0; q < 4; ++q)
          if (sreg[q] < best && kreg[q] != kLockedKey && kreg[q] != kEmptyKey && creg[q] <= 0) { best = sreg[q]; bi = lane * 4 + q; }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
          const uint64_t os = __shfl_xor_sync(0xffffffffu, best, d);
          const int oi = __shfl_xor_sync(0xffffffffu, bi, d);
          if (oi >= 0 && (bi < 0 || os < best || (os == best && oi < bi))) { best = os; bi = oi; }
        }
        if (bi >= 0) { pos = bi; result = kEvict; }                               // Reclaim vs Evict decided by the owning lane
      }
      if (pos < 0) continue;                                                     // every slot pinned: insert fails, slots[u] / rows[u] stay -1
      if (lane == (pos >> 2)) {
        const int q = pos & 3;
        uint64_t oldkey = 0;
#pragma unroll
        for (int qq = 0; qq < 4; ++qq) if (qq == q) { oldkey = kreg[qq]; kreg[qq] = key; creg[qq] += 1; }
        if (result == kEvict && oldkey == kReclaimKey) result = kReclaim;
        uint64_t* sc = a.t.scores(bk, pos);
        a.t.digests(bk)[pos] = digest_of(h);
        if (result != kEvict) a.bucket_sizes[b] += 1;
        else for (int s2 = 0; s2 < a.t.ns; ++s2) sc[s2] = 0;
        policy_update(a.pol, sc, train_score(a, u, tid), a.ts, false);
        a.t.keys(bk)[pos] = key;
        atomicAdd(a.counter + b * C + pos, 1);                                    // pin
      }
      const int64_t slot = (b - a.t.bkt_off[tid]) * C + pos;
      const int64_t row = (a.row_base ? a.row_base[tid] : 0) + slot;
      // the row itself is initialised by train_init_rows_kernel (a streaming kernel over the marks): doing it here put ~500 Philox /
      // store instructions per key on this latency-bound warp (ncu: 105 M warp instructions, 180 us)
      if (lane == 0) { a.slots[u] = slot; a.rows[u] = row; a.next[u] = -2; }
    }
    if (lane == 0) a.heads[b] = -1;                                               // leave the list heads clean for the next step
    b = bn; head = head_n;
  }
}

// Thread per touched bucket (see the file header).  Buckets that might overflow are handed to the warp kernel through touched2.
__global__ void __launch_bounds__(kBlock) train_insert_thread_kernel(TrainArgs a, int32_t* __restrict__ touched2, unsigned long long* n_touched2) {
  const int64_t nt = (int64_t)*a.n_touched;
  for (int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; w < nt; w += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = a.touched[w];
    const int head = a.heads[b];
    int cnt = 0;
    for (int cur = head; cur != -1; cur = a.next[cur]) ++cnt;
    if (a.bucket_sizes[b] + cnt > a.t.C) {                                        // may evict: cooperative scan in the warp kernel
      touched2[atomicAdd(n_touched2, 1ull)] = (int32_t)b;
      continue;
    }
    bool have_last = false; uint64_t last = 0;
    for (int r = 0; r < cnt; ++r) {                                               // lists are 1-3 long: repeated selection of the next key in order
      uint64_t best = 0; int bu = -1;
      for (int cur = head; cur != -1; cur = a.next[cur]) {
        const uint64_t k = a.ukeys[cur];
        if (have_last && !key_less(last, k, a.key_is_signed)) continue;
        if (bu < 0 || key_less(k, best, a.key_is_signed)) { best = k; bu = cur; }
      }
      const int64_t tid = a.utids ? a.utids[bu] : 0;
      const InsertOutcome o = thread_insert_one(a.t, b, best, train_score(a, bu, tid), a.pol, a.ts, a.bucket_sizes, a.counter);
      if (o.result <= kEvict) {
        const int64_t slot = (b - a.t.bkt_off[tid]) * a.t.C + o.it;
        a.slots[bu] = slot;
        a.rows[bu] = (a.row_base ? a.row_base[tid] : 0) + slot;
        atomicAdd(a.counter + b * a.t.C + o.it, 1);                               // pin
      }
      last = best; have_last = true;
    }
    for (int cur = head; cur != -1;) {                                            // the list is consumed: turn its links into "initialise me" marks
      const int nx = a.next[cur];
      a.next[cur] = a.slots[cur] >= 0 ? -2 : -4;
      cur = nx;
    }
    a.heads[b] = -1;                                                              // leave the list heads clean for the next step
  }
}

// Rows of the keys inserted in this step (next[u] == -2) get initializer + optimizer state (fused A10 + A11): a warp reads 32 marks at a
// time (coalesced), ballots, and initialises the marked rows one after the other with all 32 lanes — pure streaming writes, no dependent
// loads, every warp of the grid busy.
__global__ void __launch_bounds__(kBlock) train_init_rows_kernel(TrainArgs a) {
  const int lane = threadIdx.x & 31;
  const int64_t n = *a.n_u;
  const int64_t tiles = (n + 31) >> 5;
  const int64_t wstride = (int64_t)gridDim.x * (kBlock / 32);
  const int D4 = a.D >> 2, V4 = (int)(a.vdim >> 2);
  for (int64_t tile = (int64_t)blockIdx.x * (kBlock / 32) + (threadIdx.x >> 5); tile < tiles; tile += wstride) {
    const int64_t u0 = (tile << 5) + lane;
    const bool mine = u0 < n && a.next[u0] == -2;
    unsigned m = __ballot_sync(0xffffffffu, mine);
    const int64_t my_row = mine ? a.rows[u0] : -1;
    const uint64_t my_key = mine ? a.ukeys[u0] : 0;
    const int my_tid = (mine && a.utids) ? (int)a.utids[u0] : 0;
    while (m) {
      const int src = __ffs(m) - 1;
      m &= m - 1;
      const int64_t row = __shfl_sync(0xffffffffu, my_row, src);
      const uint64_t key = __shfl_sync(0xffffffffu, my_key, src);
      const int tid = __shfl_sync(0xffffffffu, my_tid, src);
      const InitArgs ia = a.table_init ? a.table_init[tid] : a.init;
      for (int c = lane; c < D4; c += 32) st_f4(a.values + row * a.vdim + 4 * c, init4(ia, key, c));
      for (int c = D4 + lane; c < V4; c += 32) st_f4(a.values + row * a.vdim + 4 * c, make_float4(a.state_init, a.state_init, a.state_init, a.state_init));
    }
  }
}

__global__ void counter_update_n_kernel(int32_t* counter, const int64_t* __restrict__ slots, const int64_t* __restrict__ tids,
                                        const int64_t* __restrict__ bkt_off, int64_t C, const int64_t* __restrict__ n_dev, int delta) {
  const int64_t n = *n_dev;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t s = slots[i];
    if (s >= 0) atomicAdd(counter + bkt_off[tids ? tids[i] : 0] * C + s, delta);
  }
}
__global__ void fill_i32_kernel(int32_t* p, int64_t n, int32_t v) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] = v;
}

__global__ void zero2_kernel(unsigned long long* p) { if (threadIdx.x < 3) p[threadIdx.x] = 0ull; }

size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }
int grid_for(int64_t n) { int64_t g = (n + kBlock - 1) / kBlock; const int64_t cap = (int64_t)sm_count() * 32; return (int)(g < 1 ? 1 : (g > cap ? cap : g)); }
}  // namespace

extern "C" {

int demb_fill_i32(int32_t* p, int64_t n, int32_t v, void* stream) {
  if (n <= 0) return 0;
  fill_i32_kernel<<<grid_for(n), kBlock, 0, (cudaStream_t)stream>>>(p, n, v);
  DEMB_CHECK_LAST();
  return 0;
}

int64_t demb_train_prefetch_workspace_bytes(int64_t n, int num_tables) {
  return demb_segmented_unique_workspace_bytes(n, num_tables) + (int64_t)(4 * align256(4 * (size_t)(n > 0 ? n : 1)) + 512);
}

// One-shot hook for the NEXT demb_train_prefetch issued by this host thread: `hit_flags[u]` (device, sized like the outputs) receives 1 for
// every unique key the lookup stage found, and `event` (a cudaEvent_t) is recorded on the prefetch's stream right after that stage —
// the rows of found keys are final and pinned from there on, so a consumer on another stream (the row-wise sharded wrapper's
// gather_to_peers) can start on them while insert / evict / row init are still running.  Either argument may be null.
namespace { thread_local void* g_hook_event = nullptr; thread_local int8_t* g_hook_hits = nullptr; }
int demb_train_prefetch_hook(void* event, int8_t* hit_flags) { g_hook_event = event; g_hook_hits = hit_flags; return 0; }

int demb_train_prefetch(void* storage, const int64_t* table_bucket_offsets, int64_t bucket_capacity, int num_scores, int32_t* bucket_sizes,
                        int32_t* ref_counter, int32_t* bucket_heads, float* values, int64_t value_dim, int emb_dim, const int64_t* row_base,
                        int64_t n, const int64_t* n_dev, const void* keys, const int64_t* table_range, int num_tables, const int64_t* freq_in, int policy,
                        const uint64_t* table_scores, uint64_t timestamp, int key_is_signed, int init_mode, float p0, float p1, float p2, float p3,
                        uint64_t seed, const demb_init_args_t* table_init, float state_init, void* unique_keys, int64_t* reverse_indices, int64_t* unique_table_ids,
                        int64_t* unique_freq, int64_t* slots, int64_t* rows, int64_t* num_unique, void* unique_scratch, void* workspace,
                        int64_t workspace_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  void* hook_event = g_hook_event; int8_t* hook_hits = g_hook_hits;
  g_hook_event = nullptr; g_hook_hits = nullptr;
  if (n <= 0) {
    if (hook_event) cudaEventRecord((cudaEvent_t)hook_event, stream);
    return demb_segmented_unique(0, nullptr, keys, table_range, num_tables, nullptr, unique_keys, reverse_indices, nullptr, nullptr, nullptr, num_unique, nullptr, workspace, workspace_bytes, stream_);
  }
  if (workspace_bytes < demb_train_prefetch_workspace_bytes(n, num_tables)) return DEMB_ERR_WORKSPACE;
  if ((emb_dim & 3) || (value_dim & 3) || value_dim < emb_dim) return DEMB_ERR_ARG;
  uint8_t* w = (uint8_t*)workspace;
  int32_t* next = (int32_t*)w; w += align256(4 * (size_t)n);
  int32_t* touched = (int32_t*)w; w += align256(4 * (size_t)n);
  int32_t* touched2 = (int32_t*)w; w += align256(4 * (size_t)n);
  int32_t* init_list = (int32_t*)w; w += align256(4 * (size_t)n);
  unsigned long long* n_touched = (unsigned long long*)w; w += 256;
  unsigned long long* n_touched2 = n_touched + 1;
  unsigned long long* n_init = n_touched + 2;
  const int64_t uws = (int64_t)((uint8_t*)workspace + workspace_bytes - w);
  zero2_kernel<<<1, 32, 0, stream>>>(n_touched);
  const bool need_freq = (policy == kAccumulate || policy == kLruLfu);
  int rc = demb_segmented_unique(n, n_dev, keys, table_range, num_tables, freq_in, unique_keys, reverse_indices, nullptr, need_freq ? unique_freq : nullptr,
                                 unique_table_ids, num_unique, unique_scratch, w, uws, stream_);
  if (rc) return rc;
  TrainArgs a;
  a.t = Table{(uint8_t*)storage, table_bucket_offsets, bucket_capacity, num_scores};
  a.bucket_sizes = bucket_sizes; a.counter = ref_counter; a.heads = bucket_heads;
  a.values = values; a.vdim = value_dim; a.D = emb_dim; a.row_base = row_base;
  a.ukeys = (const uint64_t*)unique_keys; a.utids = num_tables > 1 ? unique_table_ids : nullptr; a.n_u = num_unique;
  a.pol = policy; a.table_scores = table_scores; a.freq = need_freq ? unique_freq : nullptr; a.ts = timestamp; a.key_is_signed = key_is_signed;
  a.init = InitArgs{init_mode, p0, p1, p2, p3, seed}; a.table_init = reinterpret_cast<const InitArgs*>(table_init); a.state_init = state_init;
  a.full_shortcut = demb_get_option(3) != 0;
  a.slots = slots; a.rows = rows; a.next = next; a.touched = touched; a.n_touched = n_touched; a.init_list = init_list; a.n_init = n_init;
  a.hit_flags = hook_hits;
  if (bucket_capacity == kProbeC) {
    int64_t blocks = ((n + 31) / 32 + kLookupWarps - 1) / kLookupWarps;
    const int64_t cap = (int64_t)sm_count() * 3;
    train_lookup_tile_kernel<<<(int)(blocks > cap ? cap : blocks), kLookupWarps * 32, 0, stream>>>(a);
  } else {
    train_lookup_kernel<<<grid_for(n), kBlock, 0, stream>>>(a);
  }
  if (hook_event && cudaEventRecord((cudaEvent_t)hook_event, stream) != cudaSuccess) return DEMB_ERR_ARG;
  train_insert_thread_kernel<<<grid_for(n), kBlock, 0, stream>>>(a, touched2, n_touched2);
  TrainArgs a2 = a;
  a2.touched = touched2; a2.n_touched = n_touched2;
  // With a consumer hooked on the lookup stage (the sharded wrapper's copy of the found rows, on another stream) the two long kernels
  // leave a quarter of every SM's registers free — 3 resident CTAs instead of 4 at 64 registers x 256 threads — so that the consumer's CTAs
  // can be resident beside them; without that the kernels only alternate (measured at N=2: the overlapped pair took exactly the sum).
  const int hooked_bps = demb_get_option(6);
  const int bps = (hook_event && hooked_bps > 0) ? hooked_bps : 8;
  if (bucket_capacity == kProbeC) train_evict_kernel<<<sm_count() * bps, kBlock, 0, stream>>>(a2);   // buckets that may evict (all of them at steady state)
  else train_insert_kernel<<<sm_count() * 4, kBlock, 0, stream>>>(a2);
  train_init_rows_kernel<<<sm_count() * bps, kBlock, 0, stream>>>(a);
  DEMB_CHECK_LAST();
  return 0;
}

int demb_counter_update_n(int32_t* ref_counter, const int64_t* slot_indices, const int64_t* table_ids, const int64_t* table_bucket_offsets,
                          int64_t bucket_capacity, const int64_t* n_device, int64_t n_max, int delta, void* stream) {
  if (n_max <= 0) return 0;
  counter_update_n_kernel<<<grid_for(n_max), kBlock, 0, (cudaStream_t)stream>>>(ref_counter, slot_indices, table_ids, table_bucket_offsets, bucket_capacity,
                                                                                n_device, delta);
  DEMB_CHECK_LAST();
  return 0;
}

}  // extern "C"
