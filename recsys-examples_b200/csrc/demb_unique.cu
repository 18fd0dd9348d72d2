// Segmented unique (per-table dedup of an id stream) — replaces reference
// corelib/dynamicemb/src/unique_op.cu:209-717 (segmented_unique_prepare/core/finalize) and
// :471,719 (expand_table_ids), :753 (compute_dedup_lengths), index_calculation.cu:237 (get_table_range), :130 (flagged_compact).
//
// B200 design, three launches, no library calls, no host synchronisation, element count optionally read from device memory:
//   1. unique_claim_kernel      an L2-resident open-addressing scratch (2 slots of 16 B per id: 32 MiB for a 2^20-id step, inside the
//                               126 MB L2) elects, per distinct (table, key), the FIRST position it occurs at: warp-ballot dedup first
//                               (match.any groups equal keys of a warp; one leader lane does the CAS / atomicMin for its group —
//                               Zipf-hot ids mostly collide inside a warp), then one atomicMin per group.
//   2. unique_scan_emit_kernel  flag = "I am a first occurrence", numbered by a single-pass decoupled-look-back scan INSIDE the kernel
//                               (demb_scan.cuh); the first occurrence emits its key / table id / frequency at its rank, stores the rank
//                               in its scratch slot's side array and RESETS the slot — the scratch leaves every call clean, so a caller
//                               that owns a persistent scratch never pays a memset (round 1 cleared 25 MB per call).
//   3. unique_reverse_kernel    reverse_indices[i] = rank stored for i's slot.
// Unlike the reference, whose order comes from an atomicAdd race (unique_op.cu:353-377), unique_keys come out in first-occurrence order,
// deterministically.
#include "../../include/dynamicemb_b200.h"
#include "demb_common.cuh"
#include "demb_scan.cuh"

using namespace demb;

namespace {
constexpr int kBlock = 256;
constexpr int kItems = 4;                       // ids per thread in the scan kernel (1024-id tiles: 1024 CTAs for a 2^20-id step)
constexpr int kTile = kBlock * kItems;
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

// One scratch slot.  Clean state = {EmptyKey, INT32_MAX, 0}; every call leaves the scratch clean.
struct __align__(16) Slot { unsigned long long key; int32_t minpos; int32_t cnt; };
static_assert(sizeof(Slot) == 16, "Slot");

__global__ void scratch_init_kernel(Slot* s, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) s[i] = Slot{kEmptyKey, 0x7FFFFFFF, 0};
}

struct UArgs {
  int64_t n_max; const int64_t* n_dev; int mulshift;
  const uint64_t* keys; const int64_t* range; int T;
  Slot* slots; int32_t* pslot; int32_t* slot_rank;
  const int64_t* freq_in; int need_freq;
  ScanState scan; int64_t n_tiles;
  uint64_t* unique_keys; int64_t* reverse; int64_t* table_offsets; int64_t* freq_out; int64_t* unique_tids; int64_t* num_unique;
};
__device__ __forceinline__ int64_t actual_n(const UArgs& a) {
  if (!a.n_dev) return a.n_max;
  const int64_t v = *a.n_dev;
  return v < 0 ? 0 : (v > a.n_max ? a.n_max : v);
}

__global__ void __launch_bounds__(kBlock) unique_claim_kernel(UArgs a) {
  const int64_t n = actual_n(a);
  const int lane = threadIdx.x & 31;
  // this kernel runs before the scan kernel of the same call: it also zeroes that kernel's tile descriptors + ticket
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.n_tiles + 2; i += (int64_t)gridDim.x * blockDim.x) a.scan.desc[i] = 0ull;
  // grid-stride in whole warps so match.any sees a full warp
  const int64_t n32 = (n + 31) & ~(int64_t)31;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n32; i += (int64_t)gridDim.x * blockDim.x) {
    const bool act = i < n;
    const uint64_t key = act ? a.keys[i] : 0;
    const int t = (act && a.T > 1) ? table_of(a.range, a.T, i) : 0;
    // group = same (table, key); inactive lanes form singleton groups
    const unsigned gk = __match_any_sync(0xffffffffu, act ? key : (0x8000000000000000ull | (uint64_t)lane));
    const unsigned gt = __match_any_sync(0xffffffffu, act ? t : -1 - lane);
    const unsigned grp = gk & gt;
    const int leader = __ffs(grp) - 1;                       // lowest lane = lowest position of the group
    int64_t p = -1;
    if (act && lane == leader) {
      const int64_t r0 = a.range ? a.range[t] : 0, r1 = a.range ? a.range[t + 1] : n;
      // region of table t: [2*r0 + t, 2*r1 + t + 1): 2*(r1-r0) hashed slots + 1 slot reserved for key == ~0
      const int64_t base = 2 * r0 + t, size = 2 * (r1 - r0);
      if (key == kEmptyKey) {
        p = base + size;
      } else {
        int64_t q = a.mulshift ? (int64_t)__umul64hi(fmix64(key), (uint64_t)size)   // multiply-shift range reduction: no 64-bit modulo (~150 instructions)
                               : (int64_t)(fmix64(key) % (uint64_t)size);
        while (true) {
          // test before the atomic: a Zipf-hot key is claimed once and then only READ (same-address atomics serialise in L2)
          unsigned long long old = *reinterpret_cast<volatile unsigned long long*>(&a.slots[base + q].key);
          if (old == kEmptyKey) old = atomicCAS(&a.slots[base + q].key, (unsigned long long)kEmptyKey, (unsigned long long)key);
          if (old == kEmptyKey || old == key) break;
          if (++q == size) q = 0;
        }
        p = base + q;
      }
      if (*reinterpret_cast<volatile int32_t*>(&a.slots[p].minpos) > (int32_t)i) atomicMin(&a.slots[p].minpos, (int32_t)i);
      if (a.need_freq && !a.freq_in) atomicAdd(&a.slots[p].cnt, __popc(grp));
    }
    p = __shfl_sync(0xffffffffu, p, leader);
    if (act) {
      if (a.need_freq && a.freq_in) atomicAdd(&a.slots[p].cnt, (int32_t)a.freq_in[i]);
      a.pslot[i] = (int32_t)p;
    }
  }
}

__global__ void __launch_bounds__(kBlock) unique_scan_emit_kernel(UArgs a) {
  const int64_t n = actual_n(a);
  const int tile = scan_take_ticket(a.scan);
  const int64_t base = (int64_t)tile * kTile + (int64_t)threadIdx.x * kItems;
  // ids of one thread are consecutive: flags + local exclusive ranks stay in registers
  int32_t p[kItems]; int32_t mp[kItems]; bool first[kItems]; int mine = 0;
#pragma unroll
  for (int k = 0; k < kItems; ++k) p[k] = base + k < n ? a.pslot[base + k] : -1;              // 8 independent loads, then 8 more
#pragma unroll
  for (int k = 0; k < kItems; ++k) mp[k] = p[k] >= 0 ? __ldcv(&a.slots[p[k]].minpos) : -1;
#pragma unroll
  for (int k = 0; k < kItems; ++k) {
    first[k] = p[k] >= 0 && mp[k] == (int32_t)(base + k);
    mine += first[k] ? 1 : 0;
  }
  int block_total = 0;
  const int incl = block_inclusive_scan(mine, block_total);
  const unsigned int tile_prefix = chained_tile_prefix(a.scan, tile, (unsigned int)block_total);
  int64_t r = (int64_t)tile_prefix + (incl - mine);          // rank of this thread's first flagged id
#pragma unroll
  for (int k = 0; k < kItems; ++k) {
    const int64_t i = base + k;
    if (i >= n) break;
    int t = -1;
    if (a.table_offsets && a.T > 1) {                        // table_offsets[t] = number of uniques before table t's first id
      t = table_of(a.range, a.T, i);
      if (a.range[t] == i) { for (int tt = t; tt >= 0 && a.range[tt] == i; --tt) a.table_offsets[tt] = r; }
    }
    if (first[k]) {
      Slot* s = &a.slots[p[k]];
      a.unique_keys[r] = a.keys[i];
      if (a.freq_out) a.freq_out[r] = s->cnt;
      if (a.unique_tids) a.unique_tids[r] = a.T > 1 ? (t >= 0 ? t : table_of(a.range, a.T, i)) : 0;
      a.slot_rank[p[k]] = (int32_t)r;
      *s = Slot{kEmptyKey, 0x7FFFFFFF, 0};                   // leave the scratch clean (only the owner of a slot writes it here)
      ++r;
    }
    if (i == n - 1) {                                        // the last id knows the total
      if (a.num_unique) *a.num_unique = r;
      if (a.table_offsets) {
        a.table_offsets[a.T] = r;
        if (a.T == 1) a.table_offsets[0] = 0;
        else for (int tt = a.T - 1; tt >= 0 && a.range[tt] >= n; --tt) a.table_offsets[tt] = r;   // empty trailing tables
      }
    }
  }
  if (n == 0 && tile == 0 && threadIdx.x == 0) {
    if (a.num_unique) *a.num_unique = 0;
    if (a.table_offsets) for (int tt = 0; tt <= a.T; ++tt) a.table_offsets[tt] = 0;
  }
}

__global__ void __launch_bounds__(kBlock) unique_reverse_kernel(UArgs a) {
  const int64_t n = actual_n(a);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) a.reverse[i] = a.slot_rank[a.pslot[i]];
}

// unique_op.cu:471: table id of each unique key from table_offsets[T+1]
__global__ void expand_table_ids_kernel(const int64_t* __restrict__ table_offsets, int T, int64_t n, int64_t* __restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) out[i] = table_of(table_offsets, T, i);
}

// index_calculation.cu:130 flagged_compact: positions (and up to 4 gathered int64 arrays) of the flagged items, in order; the count stays
// on the device.  One launch: flags -> in-kernel chained scan -> emit.
struct CompactArgs { int64_t n; const uint8_t* flags; ScanState scan; int64_t* count; int64_t* indices; const int64_t* in[4]; int64_t* out[4]; };
__global__ void __launch_bounds__(kBlock) flagged_compact_kernel(CompactArgs a) {
  const int tile = scan_take_ticket(a.scan);
  const int64_t base = (int64_t)tile * kTile + (int64_t)threadIdx.x * kItems;
  bool f[kItems]; int mine = 0;
#pragma unroll
  for (int k = 0; k < kItems; ++k) { f[k] = (base + k < a.n) && a.flags[base + k] != 0; mine += f[k] ? 1 : 0; }
  int block_total = 0;
  const int incl = block_inclusive_scan(mine, block_total);
  int64_t r = (int64_t)chained_tile_prefix(a.scan, tile, (unsigned int)block_total) + (incl - mine);
#pragma unroll
  for (int k = 0; k < kItems; ++k) {
    const int64_t i = base + k;
    if (i >= a.n) break;
    if (f[k]) {
      a.indices[r] = i;
#pragma unroll
      for (int j = 0; j < 4; ++j) if (a.in[j]) a.out[j][r] = a.in[j][i];
      ++r;
    }
    if (i == a.n - 1) *a.count = r;
  }
}
__global__ void zero_u64_kernel(unsigned long long* p, int64_t n, int64_t* also) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] = 0ull;
  if (also && blockIdx.x == 0 && threadIdx.x == 0) *also = 0;
}

size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }
inline size_t num_slots(int64_t n, int T) { return 2 * (size_t)(n > 0 ? n : 0) + (size_t)T + 1; }
}  // namespace

extern "C" {

int demb_get_table_range(const int64_t* offsets, const int64_t* feature_offsets, int num_tables, int64_t batch_size, int64_t* table_range,
                         void* stream) {
  table_range_kernel<<<1, 128, 0, (cudaStream_t)stream>>>(offsets, feature_offsets, num_tables, batch_size, table_range);
  DEMB_CHECK_LAST();
  return num_tables < 128 ? 0 : DEMB_ERR_ARG;
}

int64_t demb_unique_scratch_bytes(int64_t n_max, int num_tables) { return (int64_t)align256(sizeof(Slot) * num_slots(n_max, num_tables)); }

int demb_unique_scratch_init(void* scratch, int64_t bytes, void* stream) {
  if (bytes <= 0) return 0;
  if (bytes % (int64_t)sizeof(Slot)) return DEMB_ERR_ARG;
  scratch_init_kernel<<<grid_for(bytes / 16), kBlock, 0, (cudaStream_t)stream>>>((Slot*)scratch, bytes / 16);
  DEMB_CHECK_LAST();
  return 0;
}

// temp workspace: [own scratch when the caller passes none] + pslot + slot_rank + scan descriptors
int64_t demb_segmented_unique_workspace_bytes(int64_t n, int num_tables) {
  if (n <= 0) return 256;
  const size_t slots = num_slots(n, num_tables);
  const int64_t n_tiles = (n + kTile - 1) / kTile;
  return (int64_t)(align256(sizeof(Slot) * slots) + align256(4 * (size_t)n) + align256(4 * slots) + align256(scan_state_bytes(n_tiles)) + 256);
}

// keys[n] grouped by table (table_range[T+1] device, nullable when T==1).  n_dev (device, nullable): the real element count
// (<= n, which then only bounds the launch and the buffers).  scratch (nullable): a persistent buffer of demb_unique_scratch_bytes(n, T)
// prepared ONCE by demb_unique_scratch_init and touched by nothing else — saves re-initialising it on every call.
// Outputs: unique_keys[>=n], reverse_indices[n] (id -> unique idx), table_offsets[T+1] (nullable), freq_out[>=n] (nullable),
//          unique_table_ids[>=n] (nullable), num_unique (device scalar, nullable).
int demb_segmented_unique(int64_t n, const int64_t* n_dev, const void* keys, const int64_t* table_range, int num_tables, const int64_t* freq_in,
                          void* unique_keys, int64_t* reverse_indices, int64_t* table_offsets, int64_t* freq_out, int64_t* unique_table_ids,
                          int64_t* num_unique, void* scratch, void* workspace, int64_t workspace_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (num_tables < 1 || num_tables > 256) return DEMB_ERR_ARG;
  if (n >= (1ll << 30)) return DEMB_ERR_ARG;
  if (num_tables > 1 && !table_range) return DEMB_ERR_ARG;
  if (n <= 0) {
    zero_u64_kernel<<<1, 32, 0, stream>>>((unsigned long long*)table_offsets, table_offsets ? num_tables + 1 : 0, num_unique);
    DEMB_CHECK_LAST();
    return 0;
  }
  if (workspace_bytes < demb_segmented_unique_workspace_bytes(n, num_tables)) return DEMB_ERR_WORKSPACE;
  const size_t slots = num_slots(n, num_tables);
  const int64_t n_tiles = (n + kTile - 1) / kTile;
  uint8_t* w = (uint8_t*)workspace;
  UArgs a;
  a.n_max = n; a.n_dev = n_dev; a.mulshift = demb_get_option(4) != 0; a.keys = (const uint64_t*)keys; a.range = table_range; a.T = num_tables;
  Slot* own = (Slot*)w; w += align256(sizeof(Slot) * slots);
  a.slots = scratch ? (Slot*)scratch : own;
  a.pslot = (int32_t*)w; w += align256(4 * (size_t)n);
  a.slot_rank = (int32_t*)w; w += align256(4 * slots);
  a.scan = scan_state_at(w, n_tiles); a.n_tiles = n_tiles;
  a.freq_in = freq_in; a.need_freq = freq_out != nullptr;
  a.unique_keys = (uint64_t*)unique_keys; a.reverse = reverse_indices; a.table_offsets = table_offsets; a.freq_out = freq_out;
  a.unique_tids = unique_table_ids; a.num_unique = num_unique;
  if (!scratch) scratch_init_kernel<<<grid_for((int64_t)slots), kBlock, 0, stream>>>(own, (int64_t)slots);
  unique_claim_kernel<<<grid_for(n), kBlock, 0, stream>>>(a);
  unique_scan_emit_kernel<<<(int)n_tiles, kBlock, 0, stream>>>(a);
  unique_reverse_kernel<<<grid_for(n), kBlock, 0, stream>>>(a);
  DEMB_CHECK_LAST();
  return 0;
}

int demb_expand_table_ids(const int64_t* table_offsets, int num_tables, int64_t n, int64_t* table_ids, void* stream) {
  if (n <= 0) return 0;
  expand_table_ids_kernel<<<grid_for(n), kBlock, 0, (cudaStream_t)stream>>>(table_offsets, num_tables, n, table_ids);
  DEMB_CHECK_LAST();
  return 0;
}

int64_t demb_flagged_compact_workspace_bytes(int64_t n) { return (int64_t)align256(scan_state_bytes((n + kTile - 1) / kTile + 1)) + 256; }

// flagged_compact (index_calculation.cu:130): indices_out[0..count) = positions i with flags[i] != 0, ascending; for each non-null
// inputs[j] (int64 arrays, up to 4) outputs[j][r] = inputs[j][indices_out[r]].  count_out: device scalar (no host sync).
int demb_flagged_compact(int64_t n, const uint8_t* flags, int64_t* count_out, int64_t* indices_out, const int64_t* const* inputs, int64_t* const* outputs,
                         int num_inputs, void* workspace, int64_t workspace_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (num_inputs < 0 || num_inputs > 4 || !count_out) return DEMB_ERR_ARG;
  const int64_t n_tiles = (n + kTile - 1) / kTile;
  if (n <= 0) { zero_u64_kernel<<<1, 32, 0, stream>>>(nullptr, 0, count_out); DEMB_CHECK_LAST(); return 0; }
  if (workspace_bytes < demb_flagged_compact_workspace_bytes(n)) return DEMB_ERR_WORKSPACE;
  CompactArgs a;
  a.n = n; a.flags = flags; a.scan = scan_state_at(workspace, n_tiles); a.count = count_out; a.indices = indices_out;
  for (int j = 0; j < 4; ++j) { a.in[j] = j < num_inputs ? inputs[j] : nullptr; a.out[j] = j < num_inputs ? outputs[j] : nullptr; }
  zero_u64_kernel<<<1, 256, 0, stream>>>(a.scan.desc, n_tiles + 2, count_out);
  flagged_compact_kernel<<<(int)n_tiles, kBlock, 0, stream>>>(a);
  DEMB_CHECK_LAST();
  return 0;
}

}  // extern "C"
