// Row-wise sharding input_dist helper: route every id of a KJT to the rank that owns it.
// Replaces reference corelib/dynamicemb/src/sparse_block_bucketize_features.cu:197-372
// (block_bucketize_sparse_features: continuous / roundrobin / hash_roundrobin, new lengths,
// bucketized ids, unbucketize permute).
//
// The reference's scatter kernel is one THREAD per (feature, sample) slot walking its ids
// sequentially (kernel2, :290-360) — 64 threads for an HSTU batch of 32 x 4096-token sequences.
// Here one WARP owns a slot and stable-partitions 32 ids per step with match.any ballots, so long
// jagged sequences run at memory speed and the output order (stable inside each (rank, slot)
// bucket) is identical.
#include <cub/device/device_scan.cuh>

#include "../../include/dynamicemb_b200.h"
#include "demb_common.cuh"

using namespace demb;

namespace {
constexpr int kWarps = 8;
constexpr int kMaxRanks = 64;

__device__ __forceinline__ void route(uint64_t idx, int dist, int W, uint64_t blk, int& p, uint64_t& new_idx) {
  if (dist == 1) { p = (int)(idx % (uint64_t)W); new_idx = idx; }                       // roundrobin
  else if (dist == 2) { p = (int)(fmix64(idx) % (uint64_t)W); new_idx = idx; }          // hash_roundrobin (:30-37)
  else {                                                                                 // continuous (:254-259)
    bool in = idx < blk * (uint64_t)W;
    p = (int)(in ? idx / blk : idx % (uint64_t)W);
    new_idx = in ? idx % blk : idx / (uint64_t)W;
  }
}

// pass 0: count -> new_lengths[p*S + slot];  pass 1: scatter using new_offsets (exclusive scan of new_lengths)
template <int PASS>
__global__ void __launch_bounds__(kWarps * 32) bucketize_kernel(int64_t S, int64_t B, int W, const int64_t* __restrict__ offsets /*[S+1]*/,
                                                                  const int64_t* __restrict__ ids, const int64_t* __restrict__ block_sizes /*[F]*/,
                                                                  const int32_t* __restrict__ dist_type /*[F] nullable*/,
                                                                  int64_t* __restrict__ new_lengths, const int64_t* __restrict__ new_offsets,
                                                                  int64_t* __restrict__ new_ids, int64_t* __restrict__ unbucketize_permute,
                                                                  const float* __restrict__ weights, float* __restrict__ new_weights) {
  __shared__ int64_t cursor[kWarps][kMaxRanks];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  for (int64_t slot = (int64_t)blockIdx.x * kWarps + wib; slot < S; slot += (int64_t)gridDim.x * kWarps) {
    const int64_t f = slot / B;
    const int dist = dist_type ? dist_type[f] : 0;
    const uint64_t blk = (uint64_t)block_sizes[f];
    for (int p = lane; p < W; p += 32) cursor[wib][p] = PASS ? new_offsets[(int64_t)p * S + slot] : 0;
    __syncwarp();
    const int64_t beg = offsets[slot], end = offsets[slot + 1];
    for (int64_t base = beg; base < end; base += 32) {
      const int64_t i = base + lane;
      const bool act = i < end;
      int p = -1 - lane; uint64_t nid = 0;     // inactive lanes get unique negative ids so match groups stay singletons
      if (act) route((uint64_t)ids[i], dist, W, blk, p, nid);
      const unsigned grp = __match_any_sync(0xffffffffu, p);
      const int before = __popc(grp & ((1u << lane) - 1));
      if (act) {
        const int64_t at = cursor[wib][p] + before;
        if (PASS) {
          new_ids[at] = (int64_t)nid;
          if (unbucketize_permute) unbucketize_permute[i] = at;
          if (weights) new_weights[at] = weights[i];
        }
      }
      __syncwarp();
      if (act && before == 0) cursor[wib][p] += __popc(grp);
      __syncwarp();
    }
    if (!PASS) for (int p = lane; p < W; p += 32) new_lengths[(int64_t)p * S + slot] = cursor[wib][p];
    __syncwarp();
  }
}
// Short slots (index-dedup re-spreads the unique ids over the B samples, compute_dedup_lengths: most slots hold 0 or 1 id): a warp per
// slot would idle 31 lanes, so one THREAD owns a slot, counts straight into new_lengths (zero-filled, entries private to the thread)
// and in the scatter pass bumps its private new_offsets entries as cursors.  Same stable order as the warp kernel.
template <int PASS>
__global__ void __launch_bounds__(256) bucketize_short_kernel(int64_t S, int64_t B, int W, const int64_t* __restrict__ offsets, const int64_t* __restrict__ ids,
                                                              const int64_t* __restrict__ block_sizes, const int32_t* __restrict__ dist_type,
                                                              int64_t* __restrict__ new_lengths, int64_t* __restrict__ new_offsets,
                                                              int64_t* __restrict__ new_ids, int64_t* __restrict__ unbucketize_permute,
                                                              const float* __restrict__ weights, float* __restrict__ new_weights) {
  for (int64_t slot = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; slot < S; slot += (int64_t)gridDim.x * blockDim.x) {
    const int64_t f = slot / B;
    const int dist = dist_type ? dist_type[f] : 0;
    const uint64_t blk = (uint64_t)block_sizes[f];
    const int64_t beg = offsets[slot], end = offsets[slot + 1];
    for (int64_t i = beg; i < end; ++i) {
      int p; uint64_t nid;
      route((uint64_t)ids[i], dist, W, blk, p, nid);
      if (!PASS) {
        new_lengths[(int64_t)p * S + slot] += 1;
      } else {
        const int64_t at = new_offsets[(int64_t)p * S + slot]++;
        new_ids[at] = (int64_t)nid;
        if (unbucketize_permute) unbucketize_permute[i] = at;
        if (weights) new_weights[at] = weights[i];
      }
    }
  }
}
size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }
}  // namespace

extern "C" {

int64_t demb_bucketize_workspace_bytes(int64_t num_slots, int world_size) {
  size_t tmp = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, tmp, (const int64_t*)nullptr, (int64_t*)nullptr, (int)(num_slots * world_size));
  return (int64_t)(align256(8 * (size_t)(num_slots * world_size)) + align256(tmp) + 256);
}

// offsets[S+1] (S = F*B slots, feature-major), ids[n].  Outputs: new_lengths[W*S], new_ids[n], unbucketize_permute[n] (nullable).
int demb_block_bucketize_sparse_features_n(int64_t num_slots, int64_t batch_size, int world_size, int64_t num_ids, const int64_t* offsets,
                                           const int64_t* ids, const int64_t* block_sizes, const int32_t* dist_type_per_feature, const float* weights,
                                           int64_t* new_lengths, int64_t* new_ids, int64_t* unbucketize_permute, float* new_weights,
                                           void* workspace, int64_t workspace_bytes, void* stream_);

int demb_block_bucketize_sparse_features(int64_t num_slots, int64_t batch_size, int world_size, const int64_t* offsets, const int64_t* ids,
                                         const int64_t* block_sizes, const int32_t* dist_type_per_feature, const float* weights,
                                         int64_t* new_lengths, int64_t* new_ids, int64_t* unbucketize_permute, float* new_weights,
                                         void* workspace, int64_t workspace_bytes, void* stream_) {
  return demb_block_bucketize_sparse_features_n(num_slots, batch_size, world_size, -1, offsets, ids, block_sizes, dist_type_per_feature, weights, new_lengths,
                                                new_ids, unbucketize_permute, new_weights, workspace, workspace_bytes, stream_);
}

// num_ids >= 0 (known on the host, = ids.numel()) lets the launcher pick the thread-per-slot kernels when slots are short on average.
int demb_block_bucketize_sparse_features_n(int64_t num_slots, int64_t batch_size, int world_size, int64_t num_ids, const int64_t* offsets,
                                           const int64_t* ids, const int64_t* block_sizes, const int32_t* dist_type_per_feature, const float* weights,
                                           int64_t* new_lengths, int64_t* new_ids, int64_t* unbucketize_permute, float* new_weights,
                                           void* workspace, int64_t workspace_bytes, void* stream_) {
  if (world_size < 1 || world_size > kMaxRanks || batch_size <= 0) return DEMB_ERR_ARG;
  if (num_slots <= 0) return 0;
  if (workspace_bytes < demb_bucketize_workspace_bytes(num_slots, world_size)) return DEMB_ERR_WORKSPACE;
  cudaStream_t stream = (cudaStream_t)stream_;
  uint8_t* w = (uint8_t*)workspace;
  int64_t* new_offsets = (int64_t*)w; w += align256(8 * (size_t)(num_slots * world_size));
  size_t tmp_bytes = (size_t)((uint8_t*)workspace + workspace_bytes - w);
  if (num_ids >= 0 && num_ids <= 4 * num_slots) {
    const int64_t blocks = (num_slots + 255) / 256;
    const int grid = (int)(blocks > (int64_t)sm_count() * 8 ? (int64_t)sm_count() * 8 : blocks);
    cudaError_t e = cudaMemsetAsync(new_lengths, 0, 8 * (size_t)(num_slots * world_size), stream);
    if (e != cudaSuccess) return -(int)e;
    bucketize_short_kernel<0><<<grid, 256, 0, stream>>>(num_slots, batch_size, world_size, offsets, ids, block_sizes, dist_type_per_feature, new_lengths,
                                                        nullptr, nullptr, nullptr, nullptr, nullptr);
    e = cub::DeviceScan::ExclusiveSum(w, tmp_bytes, new_lengths, new_offsets, (int)(num_slots * world_size), stream);
    if (e != cudaSuccess) return -(int)e;
    bucketize_short_kernel<1><<<grid, 256, 0, stream>>>(num_slots, batch_size, world_size, offsets, ids, block_sizes, dist_type_per_feature, new_lengths,
                                                        new_offsets, new_ids, unbucketize_permute, weights, new_weights);
    DEMB_CHECK_LAST();
    return 0;
  }
  int64_t blocks = (num_slots + kWarps - 1) / kWarps;
  int grid = (int)(blocks > (int64_t)sm_count() * 16 ? (int64_t)sm_count() * 16 : blocks);
  bucketize_kernel<0><<<grid, kWarps * 32, 0, stream>>>(num_slots, batch_size, world_size, offsets, ids, block_sizes, dist_type_per_feature, new_lengths,
                                                        nullptr, nullptr, nullptr, nullptr, nullptr);
  cudaError_t e = cub::DeviceScan::ExclusiveSum(w, tmp_bytes, new_lengths, new_offsets, (int)(num_slots * world_size), stream);
  if (e != cudaSuccess) return -(int)e;
  bucketize_kernel<1><<<grid, kWarps * 32, 0, stream>>>(num_slots, batch_size, world_size, offsets, ids, block_sizes, dist_type_per_feature, new_lengths,
                                                        new_offsets, new_ids, unbucketize_permute, weights, new_weights);
  DEMB_CHECK_LAST();
  return 0;
}

}  // extern "C"
