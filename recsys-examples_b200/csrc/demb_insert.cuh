// Warp-cooperative single-key insert shared by the table insert kernel (demb_table.cu) and the fused training path
// (demb_train.cu): the calling warp OWNS the bucket for the duration of the launch, so no locks are needed.
#pragma once
#include "demb_common.cuh"

namespace demb {

// ---- score policies (score.cuh:53-94), lock-free ------------------------------------------------
__device__ __forceinline__ uint64_t policy_get(int pol, const uint64_t* in, int64_t i, uint64_t ts) {
  if (pol == kConst) return 0;
  if (pol == kGlobalTimer) return ts ? ts : globaltimer();
  return in ? in[i] : 0;
}
// `atomic` = several threads may hit the same slot in this launch (lookup with duplicate keys).
__device__ __forceinline__ uint64_t policy_update(int pol, uint64_t* s, uint64_t score, uint64_t ts, bool atomic) {
  switch (pol) {
    case kConst: return s[0];
    case kAccumulate:
      if (atomic) return atomicAdd(reinterpret_cast<unsigned long long*>(s), (unsigned long long)score) + score;
      score += s[0]; s[0] = score; return score;
    case kLruLfu:
      s[0] = ts ? ts : globaltimer();
      if (atomic) return atomicAdd(reinterpret_cast<unsigned long long*>(s + 1), (unsigned long long)score) + score;
      score += s[1]; s[1] = score; return score;
    default: s[0] = score; return score;
  }
}

struct InsertOutcome { int result; int64_t it; uint64_t score; uint64_t ev_key; uint64_t ev_score; };

// Insert `key` (valid, cap > 0) into global bucket `b`.  All 32 lanes call with identical arguments; table state is
// read through volatile loads so earlier inserts of this warp into the same bucket are seen.  Probe order =
// types.cuh:325-396; eviction = min reduction score over unlocked, non-empty, unpinned slots, first minimum in storage
// order (types.cuh:417-465, kernels.cuh:238-275).  The pin counter is indexed by GLOBAL slot (bucket*C+j), as
// update_counter_with_layout_kernel writes it (insert_and_evict.cu:27-60); the reference's reduce() reads it
// table-locally, which only agrees for table 0.
__device__ __forceinline__ InsertOutcome warp_insert_one(const Table& t, int64_t b, uint64_t key, uint64_t score, int pol, uint64_t ts,
                                                         int32_t* bucket_sizes, const int32_t* counter, int lane) {
  uint8_t* bk = t.bucket(b);
  volatile uint64_t* vkeys = t.keys(bk);
  volatile uint8_t* vdig = t.digests(bk);
  const uint32_t emp4 = (uint32_t)empty_digest() * 0x01010101u;
  const int64_t h = hash63(key);
  const uint32_t want4 = (uint32_t)digest_of(h) * 0x01010101u;
  const int64_t start = (h % t.C) & ~(int64_t)15;
  int64_t hit = -1, empty = -1;
  for (int64_t base = 0; base < t.C && hit < 0 && empty < 0; base += 128) {     // 128 slots per warp step, 4 per lane
    int64_t s = base + lane * 4;
    int kind = 0; int64_t pos = -1;                   // 1 = existed, 2 = empty
    if (s < t.C) {
      int64_t p0 = start + s; if (p0 >= t.C) p0 -= t.C;
      uint32_t w = *reinterpret_cast<volatile const uint32_t*>(vdig + p0);
      uint32_t m = __vcmpeq4(w, want4) & 0x01010101u;
      while (m && !kind) { int o = (__ffs(m) - 1) >> 3; m &= m - 1; if (vkeys[p0 + o] == key) { kind = 1; pos = p0 + o; } }
      m = kind ? 0 : (__vcmpeq4(w, emp4) & 0x01010101u);
      while (m && !kind) { int o = (__ffs(m) - 1) >> 3; m &= m - 1; if (vkeys[p0 + o] == kEmptyKey) { kind = 2; pos = p0 + o; } }
    }
    unsigned any = __ballot_sync(0xffffffffu, kind != 0);
    if (any) {
      int src = __ffs(any) - 1;
      int k = __shfl_sync(0xffffffffu, kind, src);
      int64_t p = __shfl_sync(0xffffffffu, pos, src);
      if (k == 1) hit = p; else empty = p;
    }
  }
  InsertOutcome o{kInit, -1, score, 0, 0};
  if (hit >= 0) { o.result = kAssignHit; o.it = hit; }
  else if (empty >= 0) { o.result = kInsert; o.it = empty; }
  else {
    uint64_t best = 0xFFFFFFFFFFFFFFFFull; int64_t bi = -1; uint64_t bkey = 0;
    for (int64_t j = lane; j < t.C; j += 32) {
      uint64_t s = *reinterpret_cast<volatile const uint64_t*>(t.scores(bk, j) + (t.ns - 1));
      if (s < best) {
        uint64_t k = vkeys[j];
        if (k != kLockedKey && k != kEmptyKey && !(counter && *reinterpret_cast<volatile const int32_t*>(counter + b * t.C + j) > 0)) { best = s; bi = j; bkey = k; }
      }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
      uint64_t os = __shfl_xor_sync(0xffffffffu, best, d);
      int64_t oi = __shfl_xor_sync(0xffffffffu, bi, d);
      uint64_t ok = __shfl_xor_sync(0xffffffffu, bkey, d);
      bool take = (oi >= 0) && (bi < 0 || os < best || (os == best && oi < bi));
      if (take) { best = os; bi = oi; bkey = ok; }
    }
    if (bi >= 0) { o.it = bi; o.ev_key = bkey; o.ev_score = best; o.result = (bkey == kReclaimKey) ? kReclaim : kEvict; }
    else { o.result = kBusy; o.ev_key = key; o.ev_score = score; }
  }
  if (o.result <= kEvict) {
    if (lane == 0) {
      uint64_t* sc = t.scores(bk, o.it);
      if (o.result == kInsert || o.result == kReclaim || o.result == kEvict) vdig[o.it] = digest_of(h);
      if (o.result == kInsert || o.result == kReclaim) bucket_sizes[b] += 1;
      if (o.result == kEvict) for (int s = 0; s < t.ns; ++s) sc[s] = 0;
      o.score = policy_update(pol, sc, score, ts, false);
      vkeys[o.it] = key;
    }
    o.score = __shfl_sync(0xffffffffu, o.score, 0);
  }
  __threadfence_block();
  __syncwarp();
  return o;
}

// The same insert executed by ONE thread (its bucket must not be touched by any other thread of the grid): probe with probe_thread,
// serial eviction scan.  Identical outcome to warp_insert_one — same probe order, same "first minimum in storage order" victim.
__device__ __forceinline__ InsertOutcome thread_insert_one(const Table& t, int64_t b, uint64_t key, uint64_t score, int pol, uint64_t ts,
                                                           int32_t* bucket_sizes, const int32_t* counter) {
  uint8_t* bk = t.bucket(b);
  uint64_t* keys = t.keys(bk);
  uint8_t* dig = t.digests(bk);
  const int64_t h = hash63(key);
  int64_t empty = -1;
  const int64_t hit = probe_thread(t, bk, key, h, &empty);
  InsertOutcome o{kInit, -1, score, 0, 0};
  if (hit >= 0) { o.result = kAssignHit; o.it = hit; }
  else if (empty >= 0) { o.result = kInsert; o.it = empty; }
  else {
    uint64_t best = 0xFFFFFFFFFFFFFFFFull; int64_t bi = -1; uint64_t bkey = 0;
    for (int64_t j = 0; j < t.C; ++j) {
      const uint64_t s = __ldcv(t.scores(bk, j) + (t.ns - 1));
      if (s < best) {
        const uint64_t k = __ldcv(keys + j);
        if (k != kLockedKey && k != kEmptyKey && !(counter && __ldcv(counter + b * t.C + j) > 0)) { best = s; bi = j; bkey = k; }
      }
    }
    if (bi >= 0) { o.it = bi; o.ev_key = bkey; o.ev_score = best; o.result = (bkey == kReclaimKey) ? kReclaim : kEvict; }
    else { o.result = kBusy; o.ev_key = key; o.ev_score = score; }
  }
  if (o.result <= kEvict) {
    uint64_t* sc = t.scores(bk, o.it);
    if (o.result == kInsert || o.result == kReclaim || o.result == kEvict) dig[o.it] = digest_of(h);
    if (o.result == kInsert || o.result == kReclaim) bucket_sizes[b] += 1;
    if (o.result == kEvict) for (int s = 0; s < t.ns; ++s) sc[s] = 0;
    o.score = policy_update(pol, sc, score, ts, false);
    *reinterpret_cast<volatile uint64_t*>(keys + o.it) = key;
  }
  return o;
}

}  // namespace demb
