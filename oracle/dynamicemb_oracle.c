/*
 * ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product; only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference leg may load this file's library.
 *
 * CPU restatement (plain C, sequential) of the reference DynamicEmb scored hash table
 * ("LinearBucketTable") and of the row ops around it.  Every function cites the reference
 * file:line it follows (paths relative to /root/reference/corelib/dynamicemb/).
 *
 * Parity status: pinned against (a) the reference's own Python hash oracle
 * `murmur3_hash_64bits` (dynamicemb/scored_hashtable.py:279-291, executed unmodified by
 * tests/golden/gen_golden_cpu.py) and (b) golden table images / slot indices produced by the
 * reference's own CUDA kernels (`dynamicemb_extensions`, built unmodified by
 * baseline/build_ref_dynamicemb.py and run on a B200 by tests/golden/gen_golden_gpu.py).
 *
 * Storage layout (src/table_operation/types.cuh:242-284): per bucket of C slots, SoA
 *   keys[C] x 8 B | digests[C] x 1 B | scores[C x num_scores] x 8 B (AoS per key)
 * buckets contiguous.  Sentinels (types.cuh:117-121): Empty=~0, Locked=~0-2, Reclaim=~0-1.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define EMPTY_KEY   UINT64_C(0xFFFFFFFFFFFFFFFF)
#define LOCKED_KEY  UINT64_C(0xFFFFFFFFFFFFFFFD)
#define RECLAIM_KEY UINT64_C(0xFFFFFFFFFFFFFFFE)
#define RESERVE_MASK UINT64_C(0xFFFFFFFFFFFFFFFC)

enum { POL_CONST = 0, POL_ASSIGN = 1, POL_ACCUM = 2, POL_TIMER = 3, POL_LRULFU = 4 }; /* score.cuh:30-42 */
enum { R_INSERT = 0, R_RECLAIM = 1, R_ASSIGN = 2, R_EVICT = 3, R_DUP = 4, R_BUSY = 5, R_ILLEGAL = 6, R_INIT = 7 }; /* types.cuh:52-61 */

/* types.cuh:123-131  murmur3 fmix64, & INT64_MAX */
int64_t orc_hash(uint64_t k) {
  k ^= k >> 33; k *= UINT64_C(0xff51afd7ed558ccd);
  k ^= k >> 33; k *= UINT64_C(0xc4ceb9fe1a85ec53);
  k ^= k >> 33;
  return (int64_t)(k & (uint64_t)INT64_MAX);
}
/* un-masked fmix64 (sparse_block_bucketize_features.cu:30-37 uses the full 64-bit value) */
uint64_t orc_fmix64(uint64_t k) {
  k ^= k >> 33; k *= UINT64_C(0xff51afd7ed558ccd);
  k ^= k >> 33; k *= UINT64_C(0xc4ceb9fe1a85ec53);
  k ^= k >> 33;
  return k;
}
static inline uint8_t digest_of(int64_t h) { return (uint8_t)(h >> 32); }       /* types.cuh:207-210 */
static inline int is_valid(uint64_t key) { return (key & RESERVE_MASK) != RESERVE_MASK; } /* types.cuh:144-146 */
uint8_t orc_empty_digest(void) { return digest_of(orc_hash(EMPTY_KEY)); }     /* types.cuh:139-142 */

typedef struct {
  uint8_t *base; int64_t C; int64_t ns;
} bucket_t;
static inline int64_t bucket_bytes(int64_t C, int64_t ns) { return C * (8 + 1 + 8 * ns); }
static inline bucket_t get_bucket(uint8_t *storage, int64_t C, int64_t ns, int64_t idx) {
  bucket_t b = { storage + bucket_bytes(C, ns) * idx, C, ns }; return b;
}
static inline uint64_t *bkeys(bucket_t b) { return (uint64_t *)b.base; }
static inline uint8_t *bdig(bucket_t b) { return b.base + 8 * b.C; }
static inline uint64_t *bscores(bucket_t b, int64_t it) { return (uint64_t *)(b.base + 9 * b.C) + it * b.ns; }

/* scored_hashtable.py:476-496 _init_table */
void orc_table_init(uint8_t *storage, int64_t num_buckets, int64_t C, int64_t ns) {
  uint8_t ed = orc_empty_digest();
  for (int64_t i = 0; i < num_buckets; ++i) {
    bucket_t b = get_bucket(storage, C, ns, i);
    for (int64_t j = 0; j < C; ++j) { bkeys(b)[j] = EMPTY_KEY; bdig(b)[j] = ed; }
    memset(bscores(b, 0), 0, (size_t)(8 * C * ns));
  }
}

enum { P_EXISTED = 1, P_EMPTY = 2, P_EXHAUSTED = 3 };
/* types.cuh:309-396  probe: 16-aligned start, 16 digests per step, per 4-byte group: matches
 * (ascending byte) before empties (ascending byte); wrap (iter+16) % C; `step` persists so a
 * resumed probe does not rescan. */
static int probe(bucket_t b, uint64_t key, int64_t *iter, int64_t *step) {
  if (*step == b.C) return P_EXHAUSTED;
  int64_t h = orc_hash(key);
  uint8_t dg = digest_of(h), ed = orc_empty_digest();
  int64_t it = *iter;
  if (it < 0 || it > b.C) it = h % b.C;
  it &= ~(int64_t)15;
  for (; *step < b.C; *step += 16) {
    for (int g = 0; g < 4; ++g) {
      for (int o = 0; o < 4; ++o) {
        int64_t p = it + g * 4 + o;
        if (bdig(b)[p] == dg && bkeys(b)[p] == key) { *iter = p; return P_EXISTED; }
      }
      for (int o = 0; o < 4; ++o) {
        int64_t p = it + g * 4 + o;
        if (bdig(b)[p] == ed && bkeys(b)[p] == EMPTY_KEY) { *iter = p; return P_EMPTY; }
      }
    }
    it = (it + 16) % b.C;
  }
  *iter = it;
  return P_EXHAUSTED;
}

/* score.cuh:53-94 */
static inline uint64_t pol_get(int pol, const uint64_t *in, int64_t i, uint64_t timer) {
  if (pol == POL_CONST) return 0;
  if (pol == POL_TIMER) return timer;
  return in ? in[i] : 0;
}
static inline uint64_t pol_update(int pol, uint64_t *ts, uint64_t score, uint64_t timer) {
  if (pol == POL_CONST) return ts[0];
  if (pol == POL_ACCUM) { score += ts[0]; ts[0] = score; return score; }
  if (pol == POL_LRULFU) { ts[0] = timer; score += ts[1]; ts[1] = score; return score; }
  ts[0] = score; return score;
}

/* kernels.cuh:83-187 table_lookup_kernel (no overflow) */
void orc_lookup(uint8_t *storage, const int64_t *bkt_off, int64_t C, int64_t ns, int64_t n,
                const uint64_t *keys, const int64_t *table_ids, int pol, const uint64_t *score_in,
                uint64_t timer, uint8_t *founds, int64_t *indices, int64_t *score_out) {
  for (int64_t i = 0; i < n; ++i) {
    uint64_t key = keys[i];
    uint64_t score = pol_get(pol, score_in, i, timer);
    int64_t h = 0, bucket_id = 0, bb = 0, cap = 0;
    if (is_valid(key)) {
      h = orc_hash(key);
      int64_t t = table_ids ? table_ids[i] : 0;
      bb = bkt_off[t]; cap = (bkt_off[t + 1] - bb) * C;
      if (cap > 0) bucket_id = bb + (h % cap) / C;
    }
    if (cap == 0) { if (score_out) score_out[i] = (int64_t)score; founds[i] = 0; indices[i] = -1; continue; }
    bucket_t b = get_bucket(storage, C, ns, bucket_id);
    int64_t it = h % C, step = 0;
    int found = probe(b, key, &it, &step) == P_EXISTED;
    int64_t idx = -1;
    if (found) {
      if (pol == POL_CONST) score = bscores(b, it)[ns - 1];
      else score = pol_update(pol, bscores(b, it), score, timer);
      idx = (bucket_id - bb) * C + it;
    }
    if (score_out) score_out[i] = (int64_t)score;
    founds[i] = (uint8_t)found; indices[i] = idx;
  }
}

/* one insert (kernels.cuh:189-287 insert_probe + insert, :318-378 table_insert_kernel body,
 * :571-585 unlock) executed atomically — which is what the reference's deterministic
 * wave order guarantees per bucket (scored_hashtable.py:1451-1557). */
static void insert_one(uint8_t *storage, const int64_t *bkt_off, int64_t C, int64_t ns, int32_t *bucket_sizes,
                       uint64_t key, int64_t t, int pol, uint64_t score, uint64_t timer, const int32_t *counter,
                       uint8_t *result_out, int64_t *index_out, int64_t *score_out_v,
                       uint64_t *ev_key, uint64_t *ev_score, int64_t *ev_index, int *evicted_flag) {
  int64_t h = 0, bucket_id = 0, bb = 0, cap = 0;
  *evicted_flag = 0;
  if (is_valid(key)) {
    h = orc_hash(key);
    bb = bkt_off[t]; cap = (bkt_off[t + 1] - bb) * C;
    if (cap > 0) bucket_id = bb + (h % cap) / C;
  }
  if (cap == 0) { *result_out = R_ILLEGAL; *index_out = -1; *score_out_v = (int64_t)score; return; }
  bucket_t b = get_bucket(storage, C, ns, bucket_id);
  int64_t it = h % C, step = 0;
  int result = R_INIT;
  /* insert_probe */
  while (step != C) {
    int pr = probe(b, key, &it, &step);
    if (pr == P_EXISTED) { result = R_ASSIGN; break; }
    if (pr == P_EMPTY) {
      bdig(b)[it] = digest_of(h); bucket_sizes[bucket_id] += 1; result = R_INSERT; break;
    }
  }
  /* insert (evict path), types.cuh:398-465 reduce: strict '<' => first minimum wins */
  int64_t coff = (bucket_id - bb) * C;
  if (result == R_INIT) {
    uint64_t best = UINT64_MAX; int64_t bi = -1; uint64_t bk = 0;
    for (int64_t j = 0; j < C; ++j) {
      uint64_t s = bscores(b, j)[ns - 1];
      if (s < best) {
        uint64_t k = bkeys(b)[j];
        if (k != LOCKED_KEY && k != EMPTY_KEY) {
          if (counter && counter[bucket_id * C + j] > 0) continue;  /* global slot index, see demb_table.cu */
          best = s; bi = j; bk = k;
        }
      }
    }
    if (bi >= 0) {
      it = bi;
      bdig(b)[it] = digest_of(h);
      if (bk == RECLAIM_KEY) { bucket_sizes[bucket_id] += 1; result = R_RECLAIM; }
      else { for (int64_t s = 0; s < ns; ++s) bscores(b, it)[s] = 0; result = R_EVICT; }
      *ev_key = bk; *ev_score = best;
    } else {
      result = R_BUSY; *ev_key = key; *ev_score = score;
    }
  }
  int64_t index = -1;
  if (result <= R_EVICT) {
    score = pol_update(pol, bscores(b, it), score, timer);
    index = coff + it;
    bkeys(b)[it] = key; /* table_unlock_kernel */
  }
  if (result == R_EVICT) { *evicted_flag = 1; *ev_index = index; }
  else if (result == R_BUSY) { *evicted_flag = 1; *ev_index = 0; /* caller fills -(i+1) */ }
  *result_out = (uint8_t)result; *index_out = index; *score_out_v = (int64_t)score;
}

typedef struct { int64_t bucket; int64_t skey; int64_t pos; } sort_item_t;
static int cmp_item(const void *a, const void *b) {
  const sort_item_t *x = a, *y = b;
  if (x->bucket != y->bucket) return x->bucket < y->bucket ? -1 : 1;
  if (x->skey != y->skey) return x->skey < y->skey ? -1 : 1;   /* signed compare for torch.int64 keys */
  return x->pos < y->pos ? -1 : (x->pos > y->pos);
}

/* Deterministic insert[_and_evict]: order = (global bucket id, signed key) — bucketize.cu:38-58,
 * 186-199 — then "wave k = k-th key of every bucket" (scored_hashtable.py:1451-1557), which is
 * per-bucket sequential insertion in that order.  deterministic=0 inserts in the given order.
 * Evicted outputs are appended in processing order (the reference's order inside a wave is
 * racy; tests compare them as sets).  Returns the number of evicted/busy records. */
int64_t orc_insert(uint8_t *storage, const int64_t *bkt_off, int64_t C, int64_t ns, int32_t *bucket_sizes,
                   int64_t n, const uint64_t *keys, const int64_t *table_ids, int pol, const uint64_t *score_in,
                   uint64_t timer, const int32_t *counter, int deterministic,
                   uint8_t *results, int64_t *indices, int64_t *score_out,
                   uint64_t *ev_keys, int64_t *ev_scores, int64_t *ev_indices, int64_t *ev_tids) {
  sort_item_t *items = (sort_item_t *)malloc(sizeof(sort_item_t) * (size_t)(n > 0 ? n : 1));
  for (int64_t i = 0; i < n; ++i) {
    int64_t t = table_ids ? table_ids[i] : 0;
    int64_t bb = bkt_off[t], cap = (bkt_off[t + 1] - bb) * C;
    int64_t bucket = bb;
    if (cap > 0) bucket = bb + (int64_t)((uint64_t)orc_hash(keys[i]) % (uint64_t)cap) / C;
    items[i].bucket = bucket; items[i].skey = (int64_t)keys[i]; items[i].pos = i;
  }
  if (deterministic) qsort(items, (size_t)n, sizeof(sort_item_t), cmp_item);
  int64_t nev = 0;
  for (int64_t q = 0; q < n; ++q) {
    int64_t i = items[q].pos;
    int64_t t = table_ids ? table_ids[i] : 0;
    uint64_t score = pol_get(pol, score_in, i, timer);
    uint8_t res; int64_t idx, so; uint64_t ek = 0, es = 0; int64_t ei = 0; int evf = 0;
    insert_one(storage, bkt_off, C, ns, bucket_sizes, keys[i], t, pol, score, timer, counter,
               &res, &idx, &so, &ek, &es, &ei, &evf);
    if (results) results[i] = res;
    indices[i] = idx;
    if (score_out) score_out[i] = so;
    if (evf && ev_keys) {
      ev_keys[nev] = ek; ev_scores[nev] = (int64_t)es;
      ev_indices[nev] = (res == R_EVICT) ? ei : -(i + 1);   /* kernels.cuh:548-552 */
      ev_tids[nev] = t; nev++;
    }
  }
  free(items);
  return nev;
}

/* kernels.cuh:587-652 table_erase_kernel */
void orc_erase(uint8_t *storage, const int64_t *bkt_off, int64_t C, int64_t ns, int32_t *bucket_sizes,
               int64_t n, const uint64_t *keys, const int64_t *table_ids, int64_t *indices) {
  uint8_t ed = orc_empty_digest();
  for (int64_t i = 0; i < n; ++i) {
    uint64_t key = keys[i];
    int64_t h = 0, bucket_id = 0, bb = 0, cap = 0;
    if (is_valid(key)) {
      h = orc_hash(key);
      int64_t t = table_ids ? table_ids[i] : 0;
      bb = bkt_off[t]; cap = (bkt_off[t + 1] - bb) * C;
      if (cap > 0) bucket_id = bb + (h % cap) / C;
    }
    if (cap == 0) { if (indices) indices[i] = -1; continue; }
    bucket_t b = get_bucket(storage, C, ns, bucket_id);
    int64_t it = h % C, step = 0, idx = -1;
    if (probe(b, key, &it, &step) == P_EXISTED) {
      bscores(b, it)[0] = 0; bdig(b)[it] = ed; bkeys(b)[it] = RECLAIM_KEY; bucket_sizes[bucket_id] -= 1;
      idx = (bucket_id - bb) * C + it;
    }
    if (indices) indices[i] = idx;
  }
}

/* sparse_block_bucketize_features.cu:254-259 destination rank of an id.
 * mode 0 continuous (idx / blk), 1 roundrobin (idx % W), 2 hash_roundrobin (fmix64(idx) % W). */
void orc_dest_rank(int64_t n, const int64_t *ids, int mode, int64_t W, int64_t blk, int64_t *rank, int64_t *new_id) {
  for (int64_t i = 0; i < n; ++i) {
    uint64_t idx = (uint64_t)ids[i]; uint64_t p;
    if (mode == 0) { p = idx / (uint64_t)blk; if (p >= (uint64_t)W) p = (uint64_t)W - 1; }
    else if (mode == 1) p = idx % (uint64_t)W;
    else p = orc_fmix64(idx) % (uint64_t)W;
    rank[i] = (int64_t)p;
    if (new_id) new_id[i] = ids[i];
  }
}

/* ---------------------------------------------------------------------------------------------
 * Row ops (fp32).  lookup_kernel.cuh:828-857 gather; :901-962 pooled (SUM / MEAN = sum/len);
 * optimizer_kernel.cuh:41-404 optimizers.  All IEEE fp32 (no fast-math) in this restatement. */
void orc_gather_rows(const float *values, int64_t vdim, int64_t D, int64_t n, const int64_t *slots, float *out) {
  for (int64_t i = 0; i < n; ++i) {
    if (slots[i] < 0) { memset(out + i * D, 0, sizeof(float) * (size_t)D); continue; }
    memcpy(out + i * D, values + slots[i] * vdim, sizeof(float) * (size_t)D);
  }
}
/* bags: offsets[nb+1]; combiner 0 SUM, 1 MEAN; sequential fp32 accumulation in id order (the
 * product kernel accumulates in the same order, so results are bit-identical for SUM). */
void orc_pool_rows(const float *values, int64_t vdim, int64_t D, int64_t nb, const int64_t *offsets,
                   const int64_t *slots, int combiner, float *out) {
  for (int64_t b = 0; b < nb; ++b) {
    float *o = out + b * D;
    for (int64_t d = 0; d < D; ++d) o[d] = 0.f;
    int64_t s = offsets[b], e = offsets[b + 1];
    for (int64_t i = s; i < e; ++i) {
      if (slots[i] < 0) continue;
      const float *r = values + slots[i] * vdim;
      for (int64_t d = 0; d < D; ++d) o[d] = o[d] + r[d];
    }
    if (combiner == 1 && e > s) { float L = (float)(e - s); for (int64_t d = 0; d < D; ++d) o[d] = o[d] / L; }
  }
}
