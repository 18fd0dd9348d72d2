// Stable LSD radix sort of (int32 key, int32 value) pairs — the backward's "order the gradient rows by unique index" step, hand-written
// so that no library kernel is left on the training step (round 1 called cub::DeviceRadixSort: histogram + scan + 3 onesweep passes).
//
// Keys use `end_bit` (<= 31) low bits; digits of up to 11 bits => two passes for the <= 2^22 unique indices of a step.  Per pass:
//   radix_hist_kernel     per-tile (4096 keys) digit histogram in shared memory -> counts[digit][tile]
//   radix_scan_kernel     exclusive scan over counts in (digit, tile) order: single pass, decoupled look-back (demb_scan.cuh)
//   radix_scatter_kernel  every warp ranks its 512 keys round by round with match.any (equal digits of a round are numbered by lane,
//                         rounds by a per-warp digit counter in shared memory), warps are offset by a per-digit prefix over the CTA:
//                         final position = scanned base of (digit, tile) + keys of earlier warps + rank inside the warp.  Stable, hence
//                         deterministic: equal keys keep their input order, which is what fixes the summation order of the backward.
#pragma once
#include "demb_common.cuh"
#include "demb_scan.cuh"

namespace demb {
namespace rsort {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kPerThread = 16;
constexpr int kTile = kThreads * kPerThread;         // 4096 keys per CTA
constexpr int kMaxBits = 11;
constexpr int kMaxBins = 1 << kMaxBits;

__global__ void __launch_bounds__(kThreads) radix_hist_kernel(const int32_t* __restrict__ keys, int64_t n, int shift, int bins, int64_t n_tiles,
                                                              uint32_t* __restrict__ counts /*[bins][n_tiles]*/, unsigned long long* scan_desc, int64_t scan_words) {
  __shared__ uint32_t h[kMaxBins];
  // the scan kernel of this pass runs next: zero its tile descriptors + ticket here
  for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < scan_words; i += (int64_t)gridDim.x * kThreads) scan_desc[i] = 0ull;
  for (int i = threadIdx.x; i < bins; i += kThreads) h[i] = 0;
  __syncthreads();
  const int64_t base = (int64_t)blockIdx.x * kTile;
  const uint32_t mask = (uint32_t)bins - 1u;
#pragma unroll
  for (int j = 0; j < kPerThread; ++j) {
    const int64_t i = base + j * kThreads + threadIdx.x;
    if (i < n) atomicAdd(&h[((uint32_t)keys[i] >> shift) & mask], 1u);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < bins; i += kThreads) counts[(int64_t)i * n_tiles + blockIdx.x] = h[i];
}

// in-place exclusive scan of `m` uint32 counters (single pass, chained)
constexpr int kScanItems = 8;
__global__ void __launch_bounds__(kThreads) radix_scan_kernel(uint32_t* __restrict__ counts, int64_t m, ScanState st) {
  const int tile = scan_take_ticket(st);
  const int64_t base = (int64_t)tile * kThreads * kScanItems + (int64_t)threadIdx.x * kScanItems;
  uint32_t v[kScanItems]; int mine = 0;
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) { v[k] = base + k < m ? counts[base + k] : 0u; mine += (int)v[k]; }
  int total = 0;
  const int incl = block_inclusive_scan(mine, total);
  uint32_t run = chained_tile_prefix(st, tile, (unsigned int)total) + (uint32_t)(incl - mine);
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) { if (base + k < m) counts[base + k] = run; run += v[k]; }
}

__global__ void __launch_bounds__(kThreads) radix_scatter_kernel(const int32_t* __restrict__ kin, const int32_t* __restrict__ vin, int32_t* __restrict__ kout,
                                                                 int32_t* __restrict__ vout, int64_t n, int shift, int bins, int64_t n_tiles,
                                                                 const uint32_t* __restrict__ base_of /*[bins][n_tiles], scanned*/) {
  extern __shared__ uint16_t wcnt[];                 // [kWarps][bins]: keys of each digit seen so far by each warp (<= 512)
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < kWarps * bins; i += kThreads) wcnt[i] = 0;
  __syncthreads();
  const uint32_t mask = (uint32_t)bins - 1u;
  const int64_t wbase = (int64_t)blockIdx.x * kTile + (int64_t)w * (32 * kPerThread);
  int32_t key[kPerThread], val[kPerThread]; uint16_t rank[kPerThread];
  uint16_t* mine = wcnt + w * bins;
#pragma unroll
  for (int j = 0; j < kPerThread; ++j) {               // round j: keys wbase + 32 j .. + 31, one per lane, in input order
    const int64_t i = wbase + j * 32 + lane;
    const bool act = i < n;
    key[j] = act ? kin[i] : 0; val[j] = act ? vin[i] : 0;
    const uint32_t d = ((uint32_t)key[j] >> shift) & mask;
    const unsigned peers = __match_any_sync(0xffffffffu, act ? d : 0xFFFFFFFFu);
    const int leader = __ffs(peers) - 1;
    uint16_t before = 0;
    if (act && lane == leader) { before = mine[d]; mine[d] = (uint16_t)(before + __popc(peers)); }
    before = (uint16_t)__shfl_sync(0xffffffffu, (int)before, leader);
    rank[j] = (uint16_t)(before + __popc(peers & ((1u << lane) - 1u)));
    __syncwarp();
  }
  __syncthreads();
  // per digit: exclusive prefix of the warp counts over the CTA (in place)
  for (int d = threadIdx.x; d < bins; d += kThreads) {
    uint16_t run = 0;
#pragma unroll
    for (int ww = 0; ww < kWarps; ++ww) { const uint16_t c = wcnt[ww * bins + d]; wcnt[ww * bins + d] = run; run = (uint16_t)(run + c); }
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < kPerThread; ++j) {
    const int64_t i = wbase + j * 32 + lane;
    if (i < n) {
      const uint32_t d = ((uint32_t)key[j] >> shift) & mask;
      const int64_t pos = (int64_t)base_of[(int64_t)d * n_tiles + blockIdx.x] + mine[d] + rank[j];
      kout[pos] = key[j]; vout[pos] = val[j];
    }
  }
}

inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }
inline size_t workspace_bytes(int64_t n) {
  const int64_t n_tiles = (n + kTile - 1) / kTile;
  const int64_t m = (int64_t)kMaxBins * n_tiles;
  const int64_t scan_tiles = (m + kThreads * kScanItems - 1) / (kThreads * kScanItems);
  return align256(4 * (size_t)m) + align256(scan_state_bytes(scan_tiles)) + 256;
}

inline int num_passes(int end_bit) { if (end_bit < 1) end_bit = 1; if (end_bit > 31) end_bit = 31; return (end_bit + kMaxBits - 1) / kMaxBits; }

// Sorts (k0, v0)[n] by the low `end_bit` bits of the key, stable.  (k1, v1) is the ping-pong buffer; *ok / *ov say where the result is.
inline int sort_pairs(int32_t* k0, int32_t* v0, int32_t* k1, int32_t* v1, int64_t n, int end_bit, void* ws, size_t ws_bytes, cudaStream_t stream,
                      int32_t** ok, int32_t** ov) {
  *ok = k0; *ov = v0;
  if (n <= 1) return 0;
  if (ws_bytes < workspace_bytes(n)) return DEMB_ERR_WORKSPACE;
  if (end_bit < 1) end_bit = 1;
  if (end_bit > 31) end_bit = 31;
  const int passes = (end_bit + kMaxBits - 1) / kMaxBits;
  const int bits = (end_bit + passes - 1) / passes;
  const int bins = 1 << bits;
  const int64_t n_tiles = (n + kTile - 1) / kTile;
  const int64_t m = (int64_t)bins * n_tiles;
  const int64_t scan_tiles = (m + kThreads * kScanItems - 1) / (kThreads * kScanItems);
  uint32_t* counts = (uint32_t*)ws;
  ScanState st = scan_state_at((uint8_t*)ws + align256(4 * (size_t)kMaxBins * n_tiles), scan_tiles);
  static std::atomic<int> configured[kMaxDevices];
  cudaError_t ce = once_per_device(configured, [] { return cudaFuncSetAttribute(radix_scatter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kWarps * kMaxBins * 2); });
  if (ce != cudaSuccess) return -(int)ce;
  int32_t *ki = k0, *vi = v0, *ko = k1, *vo = v1;
  for (int p = 0; p < passes; ++p) {
    const int shift = p * bits;
    radix_hist_kernel<<<(int)n_tiles, kThreads, 0, stream>>>(ki, n, shift, bins, n_tiles, counts, st.desc, scan_tiles + 2);
    radix_scan_kernel<<<(int)scan_tiles, kThreads, 0, stream>>>(counts, m, st);
    radix_scatter_kernel<<<(int)n_tiles, kThreads, kWarps * bins * 2, stream>>>(ki, vi, ko, vo, n, shift, bins, n_tiles, counts);
    int32_t* t = ki; ki = ko; ko = t; t = vi; vi = vo; vo = t;
  }
  *ok = ki; *ov = vi;
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : -(int)e;
}

}  // namespace rsort
}  // namespace demb
