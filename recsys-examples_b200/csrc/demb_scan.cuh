// Single-pass device-wide exclusive scan (decoupled look-back) as a building block INSIDE other kernels: a kernel that needs "how many
// flagged items precede my tile" calls chained_tile_prefix() once per block, so flag -> scan -> emit is one launch instead of three
// (and no library scan on the step: the reference and our round-1 code called cub::DeviceScan here).
//
// Protocol (Merrill & Garland): tiles take a ticket (so every tile a block waits for is already running or done), publish their own
// aggregate, then walk back over the published descriptors of the preceding tiles until one carries an inclusive prefix.
// Descriptor = (status << 32) | value, status 0 = not yet, 1 = aggregate of that tile only, 2 = inclusive prefix up to that tile.
// The descriptor array (`n_tiles` x 8 B) and the ticket must be zero at kernel start: scan_state_bytes() / scan_state_reset().
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace demb {

struct ScanState { unsigned long long* desc; unsigned int* ticket; };

__host__ __device__ inline size_t scan_state_bytes(int64_t n_tiles) { return (size_t)(n_tiles + 2) * 8; }
// carve a ScanState out of a zeroed (or to-be-zeroed) buffer of scan_state_bytes(n_tiles)
__host__ __device__ inline ScanState scan_state_at(void* buf, int64_t n_tiles) {
  ScanState s;
  s.desc = reinterpret_cast<unsigned long long*>(buf);
  s.ticket = reinterpret_cast<unsigned int*>(s.desc + n_tiles);
  return s;
}

// block-wide: returns this block's tile index (ticket order)
__device__ __forceinline__ int scan_take_ticket(const ScanState& s) {
  __shared__ int tile_sh;
  if (threadIdx.x == 0) tile_sh = (int)atomicAdd(s.ticket, 1u);
  __syncthreads();
  const int t = tile_sh;
  __syncthreads();                                               // the slot may be rewritten by a later call
  return t;
}

// block-wide inclusive scan of one int per thread (blockDim.x multiple of 32, <= 1024); returns the thread's inclusive value, block total in `total`
__device__ __forceinline__ int block_inclusive_scan(int v, int& total) {
  __shared__ int warp_tot[32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) { int o = __shfl_up_sync(0xffffffffu, v, d); if (lane >= d) v += o; }
  if (lane == 31) warp_tot[w] = v;
  __syncthreads();
  if (w == 0) {
    int t = lane < nw ? warp_tot[lane] : 0;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { int o = __shfl_up_sync(0xffffffffu, t, d); if (lane >= d) t += o; }
    warp_tot[lane] = t;
  }
  __syncthreads();
  total = warp_tot[nw - 1];
  const int before = w > 0 ? warp_tot[w - 1] : 0;
  __syncthreads();
  return v + before;
}

// block-wide: publish `block_agg` for `tile`, return the sum of the aggregates of all tiles < tile.  All threads call it.
__device__ __forceinline__ unsigned int chained_tile_prefix(const ScanState& s, int tile, unsigned int block_agg) {
  __shared__ unsigned int prefix_sh;
  volatile unsigned long long* desc = s.desc;
  if (threadIdx.x == 0) {
    __threadfence();
    desc[tile] = ((unsigned long long)(tile == 0 ? 2u : 1u) << 32) | block_agg;
    if (tile == 0) prefix_sh = 0;
  }
  if (tile > 0 && threadIdx.x < 32) {
    const int lane = threadIdx.x;
    unsigned int excl = 0;
    int idx = tile - 1;
    while (true) {
      const int j = idx - lane;
      unsigned long long d = 0x200000000ull;                     // tiles before 0: inclusive prefix 0
      if (j >= 0) { do { d = desc[j]; } while ((d >> 32) == 0); }
      const unsigned incl = __ballot_sync(0xffffffffu, (d >> 32) == 2);
      const int stop = incl ? (__ffs(incl) - 1) : 31;            // nearest tile that already carries an inclusive prefix
      unsigned int v = lane <= stop ? (unsigned int)d : 0u;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      excl += v;
      if (incl) break;
      idx -= 32;
    }
    if (lane == 0) {
      prefix_sh = excl;
      __threadfence();
      desc[tile] = (2ull << 32) | (unsigned long long)(excl + block_agg);
    }
  }
  __syncthreads();
  const unsigned int r = prefix_sh;
  __syncthreads();                                               // the slot may be rewritten by a later call
  return r;
}

}  // namespace demb
