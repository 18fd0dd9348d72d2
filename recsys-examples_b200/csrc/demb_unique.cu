// Segmented unique (per-table dedup of an id stream) — replaces reference
// corelib/dynamicemb/src/unique_op.cu:209-717 (segmented_unique_prepare/core/finalize) and
// :471,719 (expand_table_ids), :753 (compute_dedup_lengths), index_calculation.cu:237 (get_table_range).
//
// B200 design: an L2-resident open-addressing scratch (2 slots per id, 12 B per slot; 24 MiB for a
// 2^20-id step, well inside the 126 MB L2) is used only to elect, per distinct (table,key), the
// FIRST position it occurs at (atomicMin).  Unique ids are then numbered by an exclusive scan over
// "I am a first occurrence" flags, so — unlike the reference, whose order comes from an atomicAdd
// race (unique_op.cu:353-377) — unique_keys come out in first-occurrence order, deterministically.
#include <cub/device/device_scan.cuh>

#include "../../include/dynamicemb_b200.h"
#include "demb_common.cuh"

using namespace demb;

namespace {
constexpr int kBlock = 256;
inline int grid_for(int64_t n) { int64_t g = (n + kBlock - 1) / kBlock; const int64_t cap = (int64_t)sm_count() * 64; return (int)(g < 1 ? 1 : (g > cap ? cap : g)); }

// table of position i given table_range[T+1] (ids are grouped by table; T is small)
__device__ __forceinline__ int table_of(const int64_t* __restrict__ range, int T, int64_t i) {
  int lo = 0, hi = T;   // find t with range[t] <= i < range[t+1]
  while (hi - lo > 1) { int mid = (lo + hi) >> 1; if (range[mid] <= i) lo = mid; else hi = mid; }
  return lo;
}

// index_calculation.cu:237: table_range[t] = offsets[feature_offsets[t] * B]
__global__ void table_range_kernel(const int64_t* __restrict__ offsets, const int64_t* __restrict__ feature_offsets, int T, int64_t B,
                                   int64_t* __restrict__ range) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t <= T) range[t] = offsets[feature_offsets[t] * B];
}

struct Scratch { uint64_t* keys; int32_t* minpos; int32_t* cnt; };

__global__ void unique_claim_kernel(int64_t n, const uint64_t* __restrict__ keys, const int64_t* __restrict__ range, int T, Scratch s,
                                    int32_t* __restrict__ pslot, const int64_t* __restrict__ freq_in, int need_freq) {
  const int lane = threadIdx.x & 31;
  // grid-stride in whole warps so match.any sees a full warp; Zipf-hot ids mostly collide inside a warp, where one
  // leader lane does the CAS / atomicMin for the whole group (the hot key would otherwise serialise ~7% of all atomics).
  const int64_t n32 = (n + 31) & ~(int64_t)31;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n32; i += (int64_t)gridDim.x * blockDim.x) {
    const bool act = i < n;
    const uint64_t key = act ? keys[i] : 0;
    const int t = (act && T > 1) ? table_of(range, T, i) : 0;
    // group = same (table, key); inactive lanes form singleton groups
    const unsigned gk = __match_any_sync(0xffffffffu, act ? key : (0x8000000000000000ull | (uint64_t)lane));
    const unsigned gt = __match_any_sync(0xffffffffu, act ? t : -1 - lane);
    const unsigned grp = gk & gt;
    const int leader = __ffs(grp) - 1;                       // lowest lane = lowest position of the group
    int64_t p = -1;
    if (act && lane == leader) {
      const int64_t r0 = range ? range[t] : 0, r1 = range ? range[t + 1] : n;
      // region of table t: [2*r0 + t, 2*r1 + t + 1): 2*(r1-r0) hashed slots + 1 slot reserved for key == ~0
      const int64_t base = 2 * r0 + t, size = 2 * (r1 - r0);
      if (key == kEmptyKey) {
        p = base + size;
      } else {
        int64_t q = (int64_t)(fmix64(key) % (uint64_t)size);
        while (true) {
          // test before the atomic: a Zipf-hot key is claimed once and then only READ (same-address atomics serialise in L2)
          unsigned long long old = *reinterpret_cast<volatile unsigned long long*>(s.keys + base + q);
          if (old == kEmptyKey) old = atomicCAS(reinterpret_cast<unsigned long long*>(s.keys + base + q), (unsigned long long)kEmptyKey, (unsigned long long)key);
          if (old == kEmptyKey || old == key) break;
          if (++q == size) q = 0;
        }
        p = base + q;
      }
      if (*reinterpret_cast<volatile int32_t*>(s.minpos + p) > (int32_t)i) atomicMin(s.minpos + p, (int32_t)i);
      if (need_freq && !freq_in) atomicAdd(s.cnt + p, __popc(grp));
    }
    p = __shfl_sync(0xffffffffu, p, leader);
    if (act) {
      if (need_freq && freq_in) atomicAdd(s.cnt + p, (int32_t)freq_in[i]);
      pslot[i] = (int32_t)p;
    }
  }
}

__global__ void unique_flag_kernel(int64_t n, const int32_t* __restrict__ pslot, const int32_t* __restrict__ minpos, int32_t* __restrict__ flag) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    flag[i] = (minpos[pslot[i]] == (int32_t)i) ? 1 : 0;
}

__global__ void unique_emit_kernel(int64_t n, const uint64_t* __restrict__ keys, const int32_t* __restrict__ pslot, Scratch s,
                                   const int32_t* __restrict__ rank, const int64_t* __restrict__ range, int T,
                                   uint64_t* __restrict__ unique_keys, int64_t* __restrict__ reverse, int64_t* __restrict__ table_offsets,
                                   int64_t* __restrict__ freq_out, int64_t* __restrict__ unique_tids, int64_t* __restrict__ num_unique) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int32_t p = pslot[i];
    const int32_t m = s.minpos[p];
    const int32_t r = rank[m];
    reverse[i] = r;
    if (m == (int32_t)i) {
      unique_keys[r] = keys[i];
      if (freq_out) freq_out[r] = s.cnt[p];
      if (unique_tids) unique_tids[r] = T > 1 ? table_of(range, T, i) : 0;
    }
    if (i == n - 1) {
      int64_t total = rank[n - 1] + ((s.minpos[pslot[n - 1]] == (int32_t)(n - 1)) ? 1 : 0);
      if (num_unique) *num_unique = total;
      if (table_offsets) table_offsets[T] = total;
    }
  }
  // table_offsets[t] = number of uniques before table t's first id
  if (table_offsets && blockIdx.x == 0 && threadIdx.x < T) {
    int t = threadIdx.x;
    int64_t r0 = range ? range[t] : 0;
    int64_t tot_guard = n;
    table_offsets[t] = r0 < tot_guard ? rank[r0] : -1;   // -1 patched below (empty trailing tables)
  }
}
__global__ void unique_fix_offsets_kernel(int64_t* table_offsets, int T) {
  // tables whose range starts at n (empty tail) take the total
  for (int t = T - 1; t >= 0; --t) if (table_offsets[t] < 0) table_offsets[t] = table_offsets[T];
}
__global__ void unique_empty_kernel(int64_t* table_offsets, int T, int64_t* num_unique) {
  if (threadIdx.x <= T && table_offsets) table_offsets[threadIdx.x] = 0;
  if (threadIdx.x == 0 && num_unique) *num_unique = 0;
}

// unique_op.cu:471: table id of each unique key from table_offsets[T+1]
__global__ void expand_table_ids_kernel(const int64_t* __restrict__ table_offsets, int T, int64_t n, int64_t* __restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) out[i] = table_of(table_offsets, T, i);
}

size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }
}  // namespace

extern "C" {

int demb_get_table_range(const int64_t* offsets, const int64_t* feature_offsets, int num_tables, int64_t batch_size, int64_t* table_range,
                         void* stream) {
  table_range_kernel<<<1, 128, 0, (cudaStream_t)stream>>>(offsets, feature_offsets, num_tables, batch_size, table_range);
  DEMB_CHECK_LAST();
  return num_tables < 128 ? 0 : DEMB_ERR_ARG;
}

int64_t demb_segmented_unique_workspace_bytes(int64_t n, int num_tables) {
  if (n <= 0) return 256;
  size_t slots = 2 * (size_t)n + (size_t)num_tables + 1;
  size_t scan_tmp = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, scan_tmp, (const int32_t*)nullptr, (int32_t*)nullptr, (int)n);
  return (int64_t)(align256(8 * slots) + 2 * align256(4 * slots) + 3 * align256(4 * (size_t)n) + align256(scan_tmp) + 256);
}

// keys[n] grouped by table (table_range[T+1] device, nullable when T==1).
// Outputs: unique_keys[>=n], reverse_indices[n] (id -> unique idx), table_offsets[T+1] (nullable), freq_out[>=n] (nullable),
//          unique_table_ids[>=n] (nullable), num_unique (device scalar, nullable).
int demb_segmented_unique(int64_t n, const void* keys, const int64_t* table_range, int num_tables, const int64_t* freq_in, void* unique_keys,
                          int64_t* reverse_indices, int64_t* table_offsets, int64_t* freq_out, int64_t* unique_table_ids,
                          int64_t* num_unique, void* workspace, int64_t workspace_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (num_tables < 1 || num_tables > 256) return DEMB_ERR_ARG;
  if (n <= 0) { unique_empty_kernel<<<1, 288, 0, stream>>>(table_offsets, num_tables, num_unique); DEMB_CHECK_LAST(); return 0; }
  if (n >= (1ll << 30)) return DEMB_ERR_ARG;
  if (num_tables > 1 && !table_range) return DEMB_ERR_ARG;
  if (workspace_bytes < demb_segmented_unique_workspace_bytes(n, num_tables)) return DEMB_ERR_WORKSPACE;
  size_t slots = 2 * (size_t)n + (size_t)num_tables + 1;
  uint8_t* w = (uint8_t*)workspace;
  Scratch s;
  s.keys = (uint64_t*)w; w += align256(8 * slots);
  s.minpos = (int32_t*)w; w += align256(4 * slots);
  s.cnt = (int32_t*)w; w += align256(4 * slots);
  int32_t* pslot = (int32_t*)w; w += align256(4 * (size_t)n);
  int32_t* flag = (int32_t*)w; w += align256(4 * (size_t)n);
  int32_t* rank = (int32_t*)w; w += align256(4 * (size_t)n);
  size_t tmp_bytes = (size_t)((uint8_t*)workspace + workspace_bytes - w);
  cudaMemsetAsync(s.keys, 0xFF, 8 * slots, stream);
  cudaMemsetAsync(s.minpos, 0x7F, 4 * slots, stream);   // 0x7F7F7F7F > any position
  const int need_freq = freq_out != nullptr;
  if (need_freq) cudaMemsetAsync(s.cnt, 0, 4 * slots, stream);
  unique_claim_kernel<<<grid_for(n), kBlock, 0, stream>>>(n, (const uint64_t*)keys, table_range, num_tables, s, pslot, freq_in, need_freq);
  unique_flag_kernel<<<grid_for(n), kBlock, 0, stream>>>(n, pslot, s.minpos, flag);
  cudaError_t e = cub::DeviceScan::ExclusiveSum(w, tmp_bytes, flag, rank, (int)n, stream);
  if (e != cudaSuccess) return -(int)e;
  unique_emit_kernel<<<grid_for(n), kBlock, 0, stream>>>(n, (const uint64_t*)keys, pslot, s, rank, table_range, num_tables, (uint64_t*)unique_keys,
                                                          reverse_indices, table_offsets, freq_out, unique_table_ids, num_unique);
  if (table_offsets) unique_fix_offsets_kernel<<<1, 1, 0, stream>>>(table_offsets, num_tables);
  DEMB_CHECK_LAST();
  return 0;
}

int demb_expand_table_ids(const int64_t* table_offsets, int num_tables, int64_t n, int64_t* table_ids, void* stream) {
  if (n <= 0) return 0;
  expand_table_ids_kernel<<<grid_for(n), kBlock, 0, (cudaStream_t)stream>>>(table_offsets, num_tables, n, table_ids);
  DEMB_CHECK_LAST();
  return 0;
}

}  // extern "C"
