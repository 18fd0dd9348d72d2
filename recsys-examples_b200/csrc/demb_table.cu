// DynamicEmb scored hash table — sm_100a kernels + C-ABI.
// Replaces reference corelib/dynamicemb/src/table_operation/{lookup,insert,insert_and_evict,erase,
// bucketize}.cu + kernels.cuh.  The table image, slot numbering, InsertResult / policy enums and
// eviction rule are the reference's; the execution strategy is not:
//   * insert is deterministic by construction: new keys are ordered by (global bucket, key) and one
//     warp owns each touched bucket, inserting that bucket's keys sequentially with warp-wide
//     digest/score scans (the reference reaches the same table image only in DEMB_DETERMINISM_MODE,
//     through one kernel launch per "wave", scored_hashtable.py:1451-1557).
//   * no slot locks: lookups never run concurrently with inserts on one stream, score
//     read-modify-writes use atomics where multiplicity matters.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include "../../include/dynamicemb_b200.h"
#include "demb_common.cuh"
#include "demb_insert.cuh"

using namespace demb;

namespace {

constexpr int kBlock = 256;
inline int grid_for(int64_t n, int per_block = kBlock) { int64_t g = (n + per_block - 1) / per_block; return (int)(g < 1 ? 1 : g); }

__global__ void table_init_kernel(uint8_t* storage, int64_t num_buckets, int64_t C, int ns) {
  // one thread per 16 bytes of storage
  const int64_t bb = C * (9 + 8 * (int64_t)ns);
  const int64_t total16 = num_buckets * bb / 16;
  const uint32_t ed = (uint32_t)empty_digest() * 0x01010101u;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total16; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t off = (i * 16) % bb;
    uint4 v;
    if (off < 8 * C) v = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu);
    else if (off < 9 * C) v = make_uint4(ed, ed, ed, ed);
    else v = make_uint4(0, 0, 0, 0);
    reinterpret_cast<uint4*>(storage)[i] = v;
  }
}

// A4: thread per key.  kernels.cuh:83-187.
__global__ void table_lookup_kernel(Table t, int64_t n, const uint64_t* __restrict__ keys, const int64_t* __restrict__ tids,
                                    int pol, const uint64_t* __restrict__ score_in, uint64_t ts, uint8_t* __restrict__ founds,
                                    int64_t* __restrict__ indices, int64_t* __restrict__ score_out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    uint64_t key = keys[i];
    uint64_t score = policy_get(pol, score_in, i, ts);
    Locus L = locate(t, key, tids ? tids[i] : 0);
    int64_t idx = -1;
    if (L.cap > 0) {
      uint8_t* bk = t.bucket(L.bucket);
      int64_t it = probe_thread(t, bk, key, L.h, nullptr);
      if (it >= 0) {
        if (pol == kConst) score = t.scores(bk, it)[t.ns - 1];
        else score = policy_update(pol, t.scores(bk, it), score, ts, true);
        idx = (L.bucket - L.bkt_begin) * t.C + it;
      }
    }
    if (score_out) score_out[i] = (int64_t)score;
    if (founds) founds[i] = idx >= 0;
    indices[i] = idx;
  }
}

// ---- deterministic insert ------------------------------------------------------------------------
__global__ void insert_bucket_ids_kernel(Table t, int64_t n, const uint64_t* __restrict__ keys, const int64_t* __restrict__ tids,
                                         const int32_t* __restrict__ order, int32_t* __restrict__ bucket_of_sorted) {
  for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += (int64_t)gridDim.x * blockDim.x) {
    int32_t i = order[q];
    int64_t tid = tids ? tids[i] : 0;
    int64_t bb = t.bkt_off[tid];
    int64_t cap = (t.bkt_off[tid + 1] - bb) * t.C;
    int64_t b = bb;   // bucketize.cu:38-58: cap==0 keys sort under bkt_begin
    if (cap > 0) b = bb + (int64_t)((uint64_t)hash63(keys[i]) % (uint64_t)cap) / t.C;
    bucket_of_sorted[q] = (int32_t)b;
  }
}
__global__ void iota_kernel(int32_t* p, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] = (int32_t)i;
}

struct InsertArgs {
  const uint64_t* keys; const int64_t* tids; const uint64_t* score_in; uint64_t ts; int pol;
  int32_t* bucket_sizes; const int32_t* counter;
  uint8_t* results; int64_t* indices; int64_t* score_out;
  unsigned long long* ev_count; uint64_t* ev_keys; int64_t* ev_scores; int64_t* ev_indices; int64_t* ev_tids;
};

// One warp per touched bucket: warp q runs iff sorted position q starts a bucket segment.
// Sequentially inserts that bucket's keys (ascending key) with warp-wide scans.
__global__ void __launch_bounds__(kBlock) table_insert_segments_kernel(Table t, int64_t n, const int32_t* __restrict__ order,
                                                                       const int32_t* __restrict__ bucket_sorted, InsertArgs a) {
  const int lane = threadIdx.x & 31;
  const int64_t warps_per_grid = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t q0 = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); q0 < n; q0 += warps_per_grid) {
    const int32_t b = bucket_sorted[q0];
    if (q0 > 0 && bucket_sorted[q0 - 1] == b) continue;   // not a segment start (warp-uniform)
    for (int64_t q = q0; q < n && bucket_sorted[q] == b; ++q) {
      const int32_t i = order[q];
      const uint64_t key = a.keys[i];
      const int64_t tid = a.tids ? a.tids[i] : 0;
      const uint64_t score = policy_get(a.pol, a.score_in, i, a.ts);
      const int64_t bb = t.bkt_off[tid];
      const int64_t cap = (t.bkt_off[tid + 1] - bb) * t.C;
      if (!key_is_valid(key) || cap == 0) {               // kernels.cuh:338-348
        if (lane == 0) {
          if (a.results) a.results[i] = kIllegal;
          a.indices[i] = -1;
          if (a.score_out) a.score_out[i] = (int64_t)score;
        }
        continue;
      }
      const InsertOutcome o = warp_insert_one(t, b, key, score, a.pol, a.ts, a.bucket_sizes, a.counter, lane);
      if (lane == 0) {
        const int64_t index = o.result <= kEvict ? ((int64_t)b - bb) * t.C + o.it : -1;
        if (a.results) a.results[i] = (uint8_t)o.result;
        a.indices[i] = index;
        if (a.score_out) a.score_out[i] = (int64_t)o.score;
        if (a.ev_count && (o.result == kEvict || o.result == kBusy)) {
          unsigned long long e = atomicAdd(a.ev_count, 1ull);
          a.ev_keys[e] = o.ev_key; a.ev_scores[e] = (int64_t)o.ev_score;
          a.ev_indices[e] = (o.result == kEvict) ? index : -((int64_t)i + 1);   // kernels.cuh:548-552
          a.ev_tids[e] = tid;
        }
      }
    }
  }
}

// kernels.cuh:587-652.  Thread per key; keys must be unique per call (as in the reference).
__global__ void table_erase_kernel(Table t, int64_t n, const uint64_t* __restrict__ keys, const int64_t* __restrict__ tids,
                                   int32_t* __restrict__ bucket_sizes, int64_t* __restrict__ indices) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    uint64_t key = keys[i];
    Locus L = locate(t, key, tids ? tids[i] : 0);
    int64_t idx = -1;
    if (L.cap > 0) {
      uint8_t* bk = t.bucket(L.bucket);
      int64_t it = probe_thread(t, bk, key, L.h, nullptr);
      if (it >= 0) {
        // claim through CAS so duplicated keys in one call erase once
        unsigned long long old = atomicCAS(reinterpret_cast<unsigned long long*>(t.keys(bk) + it), (unsigned long long)key, (unsigned long long)kReclaimKey);
        if (old == key) {
          t.scores(bk, it)[0] = 0;
          t.digests(bk)[it] = empty_digest();
          atomicSub(bucket_sizes + L.bucket, 1);
          idx = (L.bucket - L.bkt_begin) * t.C + it;
        }
      }
    }
    if (indices) indices[i] = idx;
  }
}

// insert_and_evict.cu:27-60 update_counter_with_layout_kernel: pin / unpin rows.
__global__ void counter_update_kernel(int32_t* counter, const int64_t* __restrict__ slots, const int64_t* __restrict__ tids,
                                      const int64_t* __restrict__ bkt_off, int64_t C, int64_t n, int delta) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t s = slots[i];
    if (s < 0) continue;
    int64_t tid = tids ? tids[i] : 0;
    atomicAdd(counter + bkt_off[tid] * C + s, delta);
  }
}

// kernels.cuh:654-708 export: compact valid (key, score, slot) triples of slots [begin,end).
__global__ void table_export_kernel(Table t, int64_t begin, int64_t end, int64_t table_begin, uint64_t threshold, int use_threshold,
                                    int score_word, unsigned long long* counter, uint64_t* keys_out, uint64_t* scores_out, int64_t* idx_out) {
  for (int64_t i0 = begin + (int64_t)blockIdx.x * blockDim.x; i0 < end; i0 += (int64_t)gridDim.x * blockDim.x) {
    int64_t i = i0 + threadIdx.x;
    bool match = false; uint64_t key = 0, score = 0;
    if (i < end) {
      uint8_t* bk = t.bucket(i / t.C);
      int64_t it = i % t.C;
      key = t.keys(bk)[it]; score = t.scores(bk, it)[score_word];
      match = key_is_valid(key) && (!use_threshold || score >= threshold);
    }
    unsigned vote = __ballot_sync(0xffffffffu, match);
    unsigned long long base = 0;
    int lane = threadIdx.x & 31;
    if (lane == 0 && vote) base = atomicAdd(counter, (unsigned long long)__popc(vote));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (match) {
      unsigned long long o = base + __popc(vote & ((1u << lane) - 1));
      keys_out[o] = key; if (scores_out) scores_out[o] = score; if (idx_out) idx_out[o] = i - table_begin;
    }
  }
}

size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

// global bucket of every key (bucketize.cu:38-58 BucketizeFunctor)
__global__ void bucket_of_kernel(Table t, int64_t n, const uint64_t* __restrict__ keys, const int64_t* __restrict__ tids, int64_t* __restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t tid = tids ? tids[i] : 0;
    const int64_t bb = t.bkt_off[tid], cap = (t.bkt_off[tid + 1] - bb) * t.C;
    out[i] = cap > 0 ? bb + (int64_t)((uint64_t)hash63(keys[i]) % (uint64_t)cap) / t.C : bb;
  }
}

}  // namespace

extern "C" {

// bucket id of each key (the first half of the reference's bucketize_keys, table_operation/bucketize.cu:111)
int demb_bucket_of(const int64_t* table_bucket_offsets, int64_t bucket_capacity, int64_t n, const void* keys, const int64_t* table_ids, int64_t* buckets,
                   void* stream) {
  if (n <= 0) return 0;
  Table t{nullptr, table_bucket_offsets, bucket_capacity, 1};
  const int64_t gcap = (int64_t)sm_count() * 16; int64_t g = (n + kBlock - 1) / kBlock;
  bucket_of_kernel<<<(int)(g > gcap ? gcap : g), kBlock, 0, (cudaStream_t)stream>>>(t, n, (const uint64_t*)keys, table_ids, buckets);
  DEMB_CHECK_LAST();
  return 0;
}

int demb_table_init(void* storage, int64_t num_buckets, int64_t bucket_capacity, int num_scores, void* stream) {
  if (bucket_capacity % 16) return DEMB_ERR_ARG;
  if (num_buckets <= 0) return 0;
  int64_t total16 = num_buckets * bucket_capacity * (9 + 8 * (int64_t)num_scores) / 16;
  const int64_t gcap = (int64_t)sm_count() * 16; int grid = (int)((total16 + kBlock - 1) / kBlock < gcap ? (total16 + kBlock - 1) / kBlock : gcap);
  table_init_kernel<<<grid < 1 ? 1 : grid, kBlock, 0, (cudaStream_t)stream>>>((uint8_t*)storage, num_buckets, bucket_capacity, num_scores);
  DEMB_CHECK_LAST();
  return 0;
}

int demb_table_lookup(void* storage, const int64_t* table_bucket_offsets, int64_t bucket_capacity, int num_scores, int64_t n,
                      const void* keys, const int64_t* table_ids, int policy, const uint64_t* score_in, uint64_t timestamp,
                      uint8_t* founds, int64_t* indices, int64_t* score_out, void* stream) {
  if (n <= 0) return 0;
  Table t{(uint8_t*)storage, table_bucket_offsets, bucket_capacity, num_scores};
  table_lookup_kernel<<<grid_for(n), kBlock, 0, (cudaStream_t)stream>>>(t, n, (const uint64_t*)keys, table_ids, policy, score_in, timestamp,
                                                                          founds, indices, score_out);
  DEMB_CHECK_LAST();
  return 0;
}

// workspace: order[n] i32 x2, bucket[n] i32 x2, skeys[n] i64 x2, cub temp
int64_t demb_table_insert_workspace_bytes(int64_t n) {
  if (n <= 0) return 256;
  size_t t1 = 0, t2 = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, t1, (const int64_t*)nullptr, (int64_t*)nullptr, (const int32_t*)nullptr, (int32_t*)nullptr, (int)n);
  cub::DeviceRadixSort::SortPairs(nullptr, t2, (const int32_t*)nullptr, (int32_t*)nullptr, (const int32_t*)nullptr, (int32_t*)nullptr, (int)n);
  size_t tmp = t1 > t2 ? t1 : t2;
  return (int64_t)(4 * align256(4 * (size_t)n) + align256(8 * (size_t)n) + align256(tmp) + 256);
}

int demb_table_insert(void* storage, const int64_t* table_bucket_offsets, int64_t bucket_capacity, int num_scores, int64_t num_buckets_total,
                      int32_t* bucket_sizes, int64_t n, const void* keys, const int64_t* table_ids, int policy, const uint64_t* score_in,
                      uint64_t timestamp, const int32_t* ref_counter, int key_is_signed, uint8_t* results, int64_t* indices,
                      int64_t* score_out, uint64_t* evicted_count, void* evicted_keys, int64_t* evicted_scores, int64_t* evicted_indices,
                      int64_t* evicted_table_ids, void* workspace, int64_t workspace_bytes, void* stream_) {
  if (n <= 0) return 0;
  if (n >= (1ll << 31)) return DEMB_ERR_ARG;
  if (workspace_bytes < demb_table_insert_workspace_bytes(n)) return DEMB_ERR_WORKSPACE;
  cudaStream_t stream = (cudaStream_t)stream_;
  Table t{(uint8_t*)storage, table_bucket_offsets, bucket_capacity, num_scores};
  uint8_t* w = (uint8_t*)workspace;
  int32_t* order_a = (int32_t*)w; w += align256(4 * (size_t)n);
  int32_t* order_b = (int32_t*)w; w += align256(4 * (size_t)n);
  int32_t* bkt_a = (int32_t*)w; w += align256(4 * (size_t)n);
  int32_t* bkt_b = (int32_t*)w; w += align256(4 * (size_t)n);
  int64_t* skeys = (int64_t*)w; w += align256(8 * (size_t)n);
  size_t tmp_bytes = (size_t)((uint8_t*)workspace + workspace_bytes - w);
  // 1) order by key (stable)  2) stable order by global bucket  => (bucket, key) order, bucketize.cu:186-199
  iota_kernel<<<grid_for(n), kBlock, 0, stream>>>(order_a, n);
  cudaError_t e;
  if (key_is_signed) e = cub::DeviceRadixSort::SortPairs(w, tmp_bytes, (const int64_t*)keys, skeys, order_a, order_b, (int)n, 0, 64, stream);
  else e = cub::DeviceRadixSort::SortPairs(w, tmp_bytes, (const uint64_t*)keys, (uint64_t*)skeys, order_a, order_b, (int)n, 0, 64, stream);
  if (e != cudaSuccess) return -(int)e;
  insert_bucket_ids_kernel<<<grid_for(n), kBlock, 0, stream>>>(t, n, (const uint64_t*)keys, table_ids, order_b, bkt_a);
  int end_bit = 1; while (end_bit < 31 && (1ll << end_bit) < num_buckets_total) ++end_bit;
  e = cub::DeviceRadixSort::SortPairs(w, tmp_bytes, bkt_a, bkt_b, order_b, order_a, (int)n, 0, end_bit, stream);
  if (e != cudaSuccess) return -(int)e;
  InsertArgs a{(const uint64_t*)keys, table_ids, score_in, timestamp, policy, bucket_sizes, ref_counter, results, indices, score_out,
               (unsigned long long*)evicted_count, (uint64_t*)evicted_keys, evicted_scores, evicted_indices, evicted_table_ids};
  int64_t blocks = (n + (kBlock / 32) - 1) / (kBlock / 32);
  const int64_t gcap = (int64_t)sm_count() * 32; int grid = (int)(blocks < gcap ? blocks : gcap);
  table_insert_segments_kernel<<<grid, kBlock, 0, stream>>>(t, n, order_a, bkt_b, a);
  DEMB_CHECK_LAST();
  return 0;
}

int demb_table_erase(void* storage, const int64_t* table_bucket_offsets, int64_t bucket_capacity, int num_scores, int32_t* bucket_sizes,
                     int64_t n, const void* keys, const int64_t* table_ids, int64_t* indices, void* stream) {
  if (n <= 0) return 0;
  Table t{(uint8_t*)storage, table_bucket_offsets, bucket_capacity, num_scores};
  table_erase_kernel<<<grid_for(n), kBlock, 0, (cudaStream_t)stream>>>(t, n, (const uint64_t*)keys, table_ids, bucket_sizes, indices);
  DEMB_CHECK_LAST();
  return 0;
}

int demb_counter_update(int32_t* ref_counter, const int64_t* slot_indices, const int64_t* table_ids, const int64_t* table_bucket_offsets,
                        int64_t bucket_capacity, int64_t n, int delta, void* stream) {
  if (n <= 0) return 0;
  counter_update_kernel<<<grid_for(n), kBlock, 0, (cudaStream_t)stream>>>(ref_counter, slot_indices, table_ids, table_bucket_offsets, bucket_capacity, n, delta);
  DEMB_CHECK_LAST();
  return 0;
}

int demb_table_export(void* storage, const int64_t* table_bucket_offsets, int64_t bucket_capacity, int num_scores, int64_t slot_begin,
                      int64_t slot_end, int64_t table_slot_begin, uint64_t threshold, int use_threshold, int score_word,
                      uint64_t* d_counter, void* keys_out, uint64_t* scores_out, int64_t* indices_out, void* stream) {
  if (slot_end <= slot_begin) return 0;
  Table t{(uint8_t*)storage, table_bucket_offsets, bucket_capacity, num_scores};
  table_export_kernel<<<grid_for(slot_end - slot_begin), kBlock, 0, (cudaStream_t)stream>>>(
      t, slot_begin, slot_end, table_slot_begin, threshold, use_threshold, score_word, (unsigned long long*)d_counter, (uint64_t*)keys_out,
      scores_out, indices_out);
  DEMB_CHECK_LAST();
  return 0;
}

}  // extern "C"
