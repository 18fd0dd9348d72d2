/*
 * dynamicemb_b200.h — C ABI of the B200-native DynamicEmb hot path (librecsys_b200.so).
 *
 * Drop-in boundary: these entry points are what the reference's pybind module
 * `dynamicemb_extensions` (corelib/dynamicemb/src/module_bind.cu:22-43) provides to the Python
 * package, re-cut as plain C: raw DEVICE pointers + sizes, a `cudaStream_t` passed as `void*`,
 * no torch types, no allocation inside (callers pass workspaces sized by the *_workspace_bytes
 * queries), no exceptions.  Every function returns 0 on success, a negative DEMB_ERR_* code, or
 * -(cudaError_t) for a CUDA failure.  All work is enqueued on `stream`; nothing synchronises.
 * Paths below are relative to /root/reference/corelib/dynamicemb/.
 *
 * Table image (byte-compatible with the reference, src/table_operation/types.cuh:242-284 and
 * dynamicemb/scored_hashtable.py:378-425): `storage` holds num_buckets buckets of
 * C=bucket_capacity slots (C % 16 == 0); bucket = keys[C] u64 | digests[C] u8 | scores[C][num_scores] u64.
 * `table_bucket_offsets[T+1]` (device, int64) gives each logical table's first global bucket.
 * Slot indices returned/accepted are table-local: (bucket - table_first_bucket) * C + position
 * (src/table_operation/kernels.cuh:149).
 */
#ifndef DYNAMICEMB_B200_H_
#define DYNAMICEMB_B200_H_
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DEMB_ERR_ARG (-1000)
#define DEMB_ERR_WORKSPACE (-1001)

/* ScorePolicyType, src/table_operation/score.cuh:30-42 */
#define DEMB_POLICY_CONST 0
#define DEMB_POLICY_ASSIGN 1
#define DEMB_POLICY_ACCUMULATE 2
#define DEMB_POLICY_GLOBAL_TIMER 3
#define DEMB_POLICY_LRU_LFU 4
/* InsertResult, src/table_operation/types.cuh:52-61 (success <=> value <= EVICT) */
#define DEMB_INSERT 0
#define DEMB_RECLAIM 1
#define DEMB_ASSIGN 2
#define DEMB_EVICT 3
#define DEMB_DUPLICATED 4
#define DEMB_BUSY 5
#define DEMB_ILLEGAL 6
#define DEMB_INIT 7
/* output dtypes */
#define DEMB_F32 0
#define DEMB_F16 1
#define DEMB_BF16 2
/* initializer modes, dynamicemb/types.py DynamicEmbInitializerMode / src/initializer.cuh:24-170 */
#define DEMB_INIT_NORMAL 0
#define DEMB_INIT_TRUNCATED_NORMAL 1
#define DEMB_INIT_UNIFORM 2
#define DEMB_INIT_DEBUG 3
#define DEMB_INIT_CONSTANT 4
/* optimizers, src/optimizer_kernel.cuh:41-404 */
#define DEMB_OPT_NONE 0
#define DEMB_OPT_SGD 1
#define DEMB_OPT_ADAM 2
#define DEMB_OPT_ADAGRAD 3
#define DEMB_OPT_ROWWISE_ADAGRAD 4

/* Per-table initializer arguments (device array indexed by table id; the reference builds one initializer per table,
 * dynamicemb/batched_dynamicemb_tables.py:789-796).  Same meaning as the scalar (mode, p0..p3, seed) arguments below. */
typedef struct { int32_t mode; float p0, p1, p2, p3; uint32_t reserved; uint64_t seed; } demb_init_args_t;

/* ---- hash table (replaces src/table_operation/{lookup,insert,insert_and_evict,erase,export_batch,bucketize}.cu) ---- */

/* scored_hashtable.py:476-496 _init_table: keys=~0, digests=digest(~0), scores=0 */
int demb_table_init(void* storage, int64_t num_buckets, int64_t bucket_capacity, int num_scores, void* stream);

/* table_lookup (src/table_operation/table.cuh:68, kernels.cuh:83-187).  keys[n] u64/i64, table_ids[n] (nullable => 0),
 * score_in[n] (ASSIGN/ACCUMULATE/LRU_LFU), timestamp: GLOBAL_TIMER score (0 => read %globaltimer per key as the
 * reference does).  Outputs founds[n] (u8, nullable), indices[n] (slot or -1), score_out[n] (nullable). */
int demb_table_lookup(void* storage, const int64_t* table_bucket_offsets, int64_t bucket_capacity, int num_scores, int64_t n,
                      const void* keys, const int64_t* table_ids, int policy, const uint64_t* score_in, uint64_t timestamp,
                      uint8_t* founds, int64_t* indices, int64_t* score_out, void* stream);

/* table_insert / table_insert_and_evict (table.cuh:81-130, kernels.cuh:189-567).  Keys must be unique per call.
 * Always deterministic: equivalent to the reference under DEMB_DETERMINISM_MODE (scored_hashtable.py:1451-1640):
 * keys ordered by (global bucket, key [signed if key_is_signed]) and inserted one per bucket at a time.
 * results[n] u8 InsertResult (nullable), indices[n] slot or -1, score_out[n] (nullable).
 * Evicted records (EVICT or BUSY, kernels.cuh:522-556) are appended at atomically claimed offsets of
 * evicted_*[>=n] when evicted_count (device u64, caller zeroes it) is non-null. */
int64_t demb_table_insert_workspace_bytes(int64_t n);
int demb_table_insert(void* storage, const int64_t* table_bucket_offsets, int64_t bucket_capacity, int num_scores, int64_t num_buckets_total,
                      int32_t* bucket_sizes, int64_t n, const void* keys, const int64_t* table_ids, int policy, const uint64_t* score_in,
                      uint64_t timestamp, const int32_t* ref_counter, int key_is_signed, uint8_t* results, int64_t* indices,
                      int64_t* score_out, uint64_t* evicted_count, void* evicted_keys, int64_t* evicted_scores, int64_t* evicted_indices,
                      int64_t* evicted_table_ids, void* workspace, int64_t workspace_bytes, void* stream);

/* table_erase (table.cuh:132, kernels.cuh:587-652): slot -> ReclaimKey, digest -> empty, score word 0 -> 0 */
int demb_table_erase(void* storage, const int64_t* table_bucket_offsets, int64_t bucket_capacity, int num_scores, int32_t* bucket_sizes,
                     int64_t n, const void* keys, const int64_t* table_ids, int64_t* indices, void* stream);

/* table_update_counter_with_layout (insert_and_evict.cu:27-60): ref_counter[bucket_off[tid]*C + slot] += delta; slot<0 skipped */
int demb_counter_update(int32_t* ref_counter, const int64_t* slot_indices, const int64_t* table_ids, const int64_t* table_bucket_offsets,
                        int64_t bucket_capacity, int64_t n, int delta, void* stream);

/* table_export_batch (export_batch.cu, kernels.cuh:654-708): compact valid (key, score[score_word], table-local slot) of global
 * slots [slot_begin, slot_end); d_counter (device u64) is the running output cursor. */
int demb_table_export(void* storage, const int64_t* table_bucket_offsets, int64_t bucket_capacity, int num_scores, int64_t slot_begin,
                      int64_t slot_end, int64_t table_slot_begin, uint64_t threshold, int use_threshold, int score_word,
                      uint64_t* d_counter, void* keys_out, uint64_t* scores_out, int64_t* indices_out, void* stream);

/* ---- dedup (replaces src/unique_op.cu, src/index_calculation.cu:237) ---- */
/* get_table_range: table_range[t] = offsets[feature_offsets[t] * batch_size], t in [0, T] */
int demb_get_table_range(const int64_t* offsets, const int64_t* feature_offsets, int num_tables, int64_t batch_size, int64_t* table_range,
                         void* stream);
/* segmented_unique_cuda (unique_op.cu:484): per-table dedup of keys[n] (grouped by table via table_range[T+1], nullable if T==1).
 * unique_keys in first-occurrence order (deterministic; the reference's order is racy), reverse_indices[n] id->unique idx,
 * table_offsets[T+1], freq_out (count or sum of freq_in per unique; nullable), unique_table_ids (nullable; = expand_table_ids),
 * num_unique device scalar (nullable).  n_dev (device, nullable): the real element count (<= n; n then bounds launches and buffers) —
 * what lets a caller chain this behind a device-side count without reading it back.  scratch (nullable): persistent dedup table of
 * demb_unique_scratch_bytes(n, T) bytes, initialised ONCE with demb_unique_scratch_init and used by nothing else (every call leaves it
 * clean); with NULL the scratch lives in `workspace` and is initialised per call.  Three launches, no library call. */
int64_t demb_unique_scratch_bytes(int64_t n_max, int num_tables);
int demb_unique_scratch_init(void* scratch, int64_t bytes, void* stream);
int64_t demb_segmented_unique_workspace_bytes(int64_t n, int num_tables);
int demb_segmented_unique(int64_t n, const int64_t* n_dev, const void* keys, const int64_t* table_range, int num_tables, const int64_t* freq_in,
                          void* unique_keys, int64_t* reverse_indices, int64_t* table_offsets, int64_t* freq_out, int64_t* unique_table_ids,
                          int64_t* num_unique, void* scratch, void* workspace, int64_t workspace_bytes, void* stream);
/* flagged_compact (index_calculation.cu:130): indices_out[0..*count_out) = positions with flags != 0, ascending; outputs[j][r] =
 * inputs[j][indices_out[r]] for up to 4 int64 arrays.  The count stays on the device (the reference reads it back: a host sync). */
int64_t demb_flagged_compact_workspace_bytes(int64_t n);
int demb_flagged_compact(int64_t n, const uint8_t* flags, int64_t* count_out, int64_t* indices_out, const int64_t* const* inputs,
                         int64_t* const* outputs, int num_inputs, void* workspace, int64_t workspace_bytes, void* stream);
/* expand_table_ids_cuda (unique_op.cu:471,719) */
int demb_expand_table_ids(const int64_t* table_offsets, int num_tables, int64_t n, int64_t* table_ids, void* stream);

/* ---- rows (replaces src/lookup_forward.cu, lookup_backward.cu, dynamic_emb_op.cu, optimizer.cu, initializer.cu) ---- */
/* Value table: fp32 rows `values[row * value_dim + 0..]` = [embedding(emb_dim) | optimizer state] (key_value_table.py:346-356);
 * global row of (table t, slot s) = row_base[t] + s.  emb_dim % 4 == 0, value_dim % 4 == 0, emb_dim <= 1024. */

/* FUSED forward for ids already in the table (eval path, batched_dynamicemb_function.py:836 dynamicemb_eval_forward;
 * fuses table_lookup + load_from_flat + gather_embedding[_pooled]).  combiner -1: sequence, out[n, D];
 * 0 SUM / 1 MEAN: pooled over bags given by offsets[F*B+1] (feature-major), out[B, F*D].  Absent ids contribute
 * `absent_value` (sequence) or nothing (pooled).  founds[n]/slots_out[n] optional. */
int demb_lookup_forward(void* storage, const int64_t* table_bucket_offsets, int64_t bucket_capacity, int num_scores, const float* values,
                        int64_t value_dim, int emb_dim, const int64_t* row_base, int64_t n, const void* keys, const int64_t* table_range,
                        int num_tables, const int64_t* offsets, int64_t batch_size, int num_features, int combiner, void* out, int out_dtype,
                        float absent_value, uint8_t* founds, int64_t* slots_out, void* stream);
/* training forward after prefetch (DynamicEmbeddingFunction.forward, batched_dynamicemb_function.py:1044: load_from_flat +
 * gather_embedding[_pooled] in one pass): row of id i = rows[inverse[i]] (inverse nullable => rows[i]); rows<0 => zeros */
int demb_gather_forward(const float* values, int64_t value_dim, int emb_dim, int64_t n, const int64_t* rows, const int64_t* inverse,
                        const int64_t* offsets, int64_t batch_size, int num_features, int combiner, void* out, int out_dtype,
                        const int64_t* n_dev /* device, nullable: real id count <= n (sequence mode) */, void* stream);
int demb_rows_from_slots(int64_t n, const int64_t* slots, const int64_t* table_ids, const int64_t* row_base, int64_t* rows, void* stream);
/* initializer + store_to_flat fused (initializer.cu, dynamic_emb_op.cu:400-490): values[rows[i]] = [init(keys[i]) | state_init];
 * params: UNIFORM(p0=lower,p1=upper) NORMAL(p0=mean,p1=std) TRUNCATED_NORMAL(+p2=lower,p3=upper) CONSTANT(p0) DEBUG(key%100000).
 * table_init (device, nullable) + table_ids[n] (nullable => 0): per-table arguments that override the scalars.
 * only_if[n] (nullable) masks rows; emb_out[n,D] (nullable) also receives the embedding. */
int demb_init_rows(float* values, int64_t value_dim, int emb_dim, int64_t n, const int64_t* rows, const void* keys, int mode, float p0, float p1,
                   float p2, float p3, uint64_t seed, const int64_t* table_ids, const demb_init_args_t* table_init, float state_init,
                   const uint8_t* only_if, float* emb_out, void* stream);
/* load_from_flat_table / store_to_flat_table (dynamic_emb_op.cu:295-490): copy `width` floats per row table<->dense */
int demb_copy_rows(float* values, int64_t value_dim, int width, int64_t n, const int64_t* rows, float* dense, int64_t dense_stride, int to_table,
                   void* stream);
/* FUSED backward (DynamicEmbeddingFunction.backward, :1194: reduce_grads + fused_update_for_flat_table).  See demb_rows.cu.
 * Optional (nullable) extras used by the row-wise sharded path: n_dev = device-side id count (<= n); grad_row_of[n] = gradient row id of
 * id i when gradient rows are not stored in id order (sequence mode); unique_grad_addr[num_unique] = destination ADDRESS of each
 * reduced gradient row (0 = drop) instead of unique_grads + u*D — may point into a peer GPU's memory (NVLink stores). */
int64_t demb_backward_workspace_bytes(int64_t n, int emb_dim);
int demb_backward(float* values, int64_t value_dim, int emb_dim, int64_t n, const int64_t* inverse, int64_t num_unique_bound, const int64_t* rows,
                  const float* grads, int64_t grad_stride, const int64_t* offsets, int64_t batch_size, int num_features, int combiner,
                  int opt_type, float lr, float eps, float beta1, float beta2, float weight_decay, float bias_correction1,
                  float bias_correction2, float* unique_grads, const int64_t* n_dev, const int64_t* grad_row_of, const int64_t* unique_grad_addr,
                  void* workspace, int64_t workspace_bytes, void* stream);
/* Split form: the gradient-independent half of demb_backward (pair list + radix sort by unique index) can be launched right after the
   prefetch, on a stream of the caller's, where it overlaps the forward gather; demb_backward_apply does the rest.  Same n / inverse /
   workspace in both calls; the caller orders apply behind sort (events), one outstanding sort per workspace. */
int demb_backward_sort(int emb_dim, int64_t n, const int64_t* inverse, int64_t num_unique_bound, const int64_t* offsets, int64_t batch_size,
                       int num_features, int combiner, const int64_t* n_dev, const int64_t* grad_row_of, void* workspace, int64_t workspace_bytes,
                       void* stream);
int demb_backward_apply(float* values, int64_t value_dim, int emb_dim, int64_t n, const int64_t* inverse, int64_t num_unique_bound,
                        const int64_t* rows, const float* grads, int64_t grad_stride, const int64_t* offsets, int64_t batch_size, int num_features,
                        int combiner, int opt_type, float lr, float eps, float beta1, float beta2, float weight_decay, float bias_correction1,
                        float bias_correction2, float* unique_grads, const int64_t* n_dev, const int64_t* unique_grad_addr, void* workspace,
                        int64_t workspace_bytes, void* stream);
/* load_from_flat_table_{contiguous,emb,value} / store_to_flat_table_{contiguous,value} (dynamic_emb_op.cu:295-490) in the reference's own
 * addressing: per-table base pointers (int64, device), per-table value / embedding widths (mixed dims allowed), per-row table id (or one
 * scalar table id), per-row index (< 0 skipped).  region 0 = contiguous prefix, 1 = embedding only, 2 = [emb | pad to max_emb_dim | state]. */
int demb_flat_table_copy(const int64_t* table_ptrs, const int64_t* table_ids, int64_t scalar_table_id, const int64_t* indices, int64_t n,
                         const int64_t* table_value_dims, const int64_t* table_emb_dims, int64_t max_emb_dim, float* dense, int64_t dense_stride,
                         int64_t dense_dim, int region, int to_table, void* stream);
/* {sgd,adam,adagrad,rowwise_adagrad}_update_for_flat_table (optimizer.cu:416-447), same addressing; grads[n, max_emb_dim] */
int demb_flat_table_update(const int64_t* table_ptrs, const int64_t* table_ids, const int64_t* indices, int64_t n, const int64_t* table_value_dims,
                           const int64_t* table_emb_dims, int64_t max_emb_dim, const float* grads, int64_t grad_stride, int opt_type, float lr,
                           float eps, float beta1, float beta2, float weight_decay, float bias_correction1, float bias_correction2, void* stream);
/* global bucket of each key (first half of bucketize_keys, table_operation/bucketize.cu:38-58,111) */
int demb_bucket_of(const int64_t* table_bucket_offsets, int64_t bucket_capacity, int64_t n, const void* keys, const int64_t* table_ids, int64_t* buckets,
                   void* stream);
/* {sgd,adam,adagrad,rowwise_adagrad}_update on one uniform value table: dense grads[n, D] -> rows */
int demb_update_rows(float* values, int64_t value_dim, int emb_dim, int64_t n, const int64_t* rows, const float* grads, int64_t grad_stride,
                     int opt_type, float lr, float eps, float beta1, float beta2, float weight_decay, float bias_correction1,
                     float bias_correction2, void* stream);

/* ---- fused training prefetch (replaces dynamicemb_prefetch / _prefetch_hbm_direct_path, batched_dynamicemb_function.py:559-830) ----
 * dedup -> probe (+score update, +pin) -> insert + row init of the missing keys (+pin), no host synchronisation, no sort.
 * bucket_heads[num_buckets] int32 must be -1 on entry and is left -1 (demb_fill_i32).  Outputs are sized n; only the first
 * *num_unique entries are meaningful.  slots: table-local slot or -1 (insert failed); rows: global value row or -1.
 * table_scores[T] (ASSIGN policies: one score per table); unique_freq[n] is filled for ACCUMULATE / LRU_LFU. */
int demb_fill_i32(int32_t* p, int64_t n, int32_t v, void* stream);
int64_t demb_train_prefetch_workspace_bytes(int64_t n, int num_tables);
int demb_train_prefetch(void* storage, const int64_t* table_bucket_offsets, int64_t bucket_capacity, int num_scores, int32_t* bucket_sizes,
                        int32_t* ref_counter, int32_t* bucket_heads, float* values, int64_t value_dim, int emb_dim, const int64_t* row_base,
                        int64_t n, const int64_t* n_dev /* device, nullable: real id count <= n */, const void* keys, const int64_t* table_range,
                        int num_tables, const int64_t* freq_in, int policy,
                        const uint64_t* table_scores, uint64_t timestamp, int key_is_signed, int init_mode, float p0, float p1, float p2, float p3,
                        uint64_t seed, const demb_init_args_t* table_init /* device [num_tables], nullable */, float state_init,
                        void* unique_keys, int64_t* reverse_indices, int64_t* unique_table_ids,
                        int64_t* unique_freq, int64_t* slots, int64_t* rows, int64_t* num_unique, void* unique_scratch /* nullable, see
                        demb_segmented_unique */, void* workspace, int64_t workspace_bytes, void* stream);
/* One-shot hook for the NEXT demb_train_prefetch issued by the calling host thread: hit_flags[u] (device int8, sized like the outputs,
 * nullable) receives 1 for every unique key the lookup stage found, and `event` (cudaEvent_t, nullable) is recorded on the prefetch's
 * stream right after that stage: rows[] / slots[] of found keys are final and pinned from there on.  (No reference counterpart: the
 * reference's prefetch synchronises with the host between its stages.) */
int demb_train_prefetch_hook(void* event, int8_t* hit_flags);
/* demb_counter_update with the element count read from device memory (*n_device <= n_max) */
int demb_counter_update_n(int32_t* ref_counter, const int64_t* slot_indices, const int64_t* table_ids, const int64_t* table_bucket_offsets,
                          int64_t bucket_capacity, const int64_t* n_device, int64_t n_max, int delta, void* stream);

/* measurement aid for bench.py: CUDA events around the stages of demb_backward; read returns ms of {pairs+sort, tiles, spans} */
int demb_set_option(int option, int value);   /* A/B switches for measurements; see csrc/demb_rows.cu */
int demb_get_option(int option);
int demb_profile_enable(int on);
int demb_profile_read(float* ms3);

/* ---- row-wise sharding input dist (replaces src/sparse_block_bucketize_features.cu:372) ---- */
int64_t demb_bucketize_workspace_bytes(int64_t num_slots, int world_size);
/* dist_type_per_feature[F]: 0 continuous, 1 roundrobin, 2 hash_roundrobin; block_sizes[F].  new_lengths[W*S] rank-major. */
int demb_block_bucketize_sparse_features(int64_t num_slots, int64_t batch_size, int world_size, const int64_t* offsets, const int64_t* ids,
                                         const int64_t* block_sizes, const int32_t* dist_type_per_feature, const float* weights,
                                         int64_t* new_lengths, int64_t* new_ids, int64_t* unbucketize_permute, float* new_weights,
                                         void* workspace, int64_t workspace_bytes, void* stream);
/* same, with the id count known on the host (ids.numel()): short slots (<= 4 ids on average, e.g. after index dedup) take a
   thread-per-slot kernel pair instead of the warp-per-slot one.  num_ids < 0 = unknown. */
int demb_block_bucketize_sparse_features_n(int64_t num_slots, int64_t batch_size, int world_size, int64_t num_ids, const int64_t* offsets,
                                           const int64_t* ids, const int64_t* block_sizes, const int32_t* dist_type_per_feature,
                                           const float* weights, int64_t* new_lengths, int64_t* new_ids, int64_t* unbucketize_permute,
                                           float* new_weights, void* workspace, int64_t workspace_bytes, void* stream);

/* ---- row-wise sharded exchange over NVLink peer memory (replaces RwSparseFeaturesDist + KJTAllToAll + SequenceEmbeddingsAllToAll:
 * dynamicemb/input_dist.py:225-285, shard/embedding.py:183-340, planner/rw_sharding.py:83,189).  csrc/demb_shard.cu explains the flow.
 * Every rank owns ONE symmetric buffer of layout[5] bytes mapped on all ranks; `peers[W]` (device) holds the W base addresses as mapped on
 * the calling rank.  pair_cap = most ids one rank may send to one owner per step, n_cap = most unique ids a rank may request per step,
 * recv_cap = most ids an owner accepts per step; exceeding one drops ids and sets *err = 1 (a barrier timeout sets 2).  All counts stay
 * on the device. */
int demb_shard_layout(int world, int64_t pair_cap, int64_t n_cap, int emb_dim, int64_t* out6 /* flags, meta, ids_in, rows_back, grads_in, total */);
int64_t demb_shard_route_workspace_bytes(int64_t n_max, int world, int num_tables);
int demb_shard_route(int world, int rank, int num_tables, int emb_dim, int64_t pair_cap, int64_t n_cap, const int64_t* peers, int32_t* err,
                     int64_t n_max, const int64_t* n_unique_dev, const void* unique_keys, const int64_t* unique_table_ids,
                     const int32_t* dist_type_per_table, const int64_t* block_size_per_table, int64_t* send_pos, int64_t* unique_grad_addr,
                     void* state /* persistent, zeroed once */, int64_t state_bytes, void* stream);
int64_t demb_shard_recv_workspace_bytes(int world, int num_tables);
int demb_shard_recv(int world, int rank, int num_tables, int emb_dim, int64_t pair_cap, int64_t n_cap, int64_t recv_cap, const int64_t* peers,
                    int32_t* err, void* ids_recv, int64_t* table_range, int64_t* n_recv, int64_t* src_pos, int64_t* dst_addr, void* workspace,
                    int64_t workspace_bytes, void* stream);
int demb_shard_gather_to_peers(const float* values, int64_t value_dim, int emb_dim, int64_t n_max, const int64_t* n_dev, const int64_t* rows,
                               const int64_t* inverse, const int64_t* dst_addr, void* stream);
/* the same copy restricted by per-UNIQUE-id flags (see demb_train_prefetch_hook): part 1 = ids whose key the prefetch's lookup stage
 * found (may run on another stream once the hook's event has fired, concurrently with insert / evict / row init), part 2 = the others
 * (after the prefetch), part 0 = all.  Lets the NVLink copy of the hit rows overlap the rest of the owner's prefetch. */
int demb_shard_gather_to_peers_part(const float* values, int64_t value_dim, int emb_dim, int64_t n_max, const int64_t* n_dev, const int64_t* rows,
                                    const int64_t* inverse, const int64_t* dst_addr, const int8_t* hit_flags, int part, void* stream);
int demb_peer_barrier(int world, int rank, int64_t pair_cap, int64_t n_cap, int emb_dim, const int64_t* peers, int32_t* err, int channel,
                      uint64_t* epochs /* 4 x u64 local device memory, zero at start */, void* stream);
int demb_zero_i64(int64_t* p, int64_t n, void* stream);
/* one-time setup of a symmetric buffer: cudaMalloc (zero-filled) + 64-byte CUDA IPC handle; peers open the handle (peer access enabled) */
int demb_ipc_alloc(int64_t bytes, void** ptr, void* handle64);
int demb_ipc_open(const void* handle64, void** ptr);
int demb_ipc_close(void* ptr);
int demb_ipc_free(void* ptr);

#ifdef __cplusplus
}
#endif
#endif
