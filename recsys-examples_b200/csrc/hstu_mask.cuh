// HSTU attention mask, restated from the reference (third_party/FBGEMM/fbgemm_gpu/experimental/hstu/src/hstu_blackwell/mask.py:61-127,
// test/hstu_test.py:86-171) in a form that is branch-free per element: for a FIXED query row the valid keys — and for a FIXED key the
// valid queries — are the union of at most two half-open intervals, so each score needs two unsigned compares instead of a chain of
// data-dependent branches (the branchy form blew the instruction cache: ncu showed stall_no_inst on every reconvergence point).
//   valid(row, col) = col < L  and  row < L
//                     and (wr < 0 or col <= row + wr) and (wl < 0 or col >= row - wl)                       causal / local window
//                     and not (targets: row >= h and col >= h and col < h + floor((row-h)/G)*G)              other target groups hidden
//                     or  (contexts: row < c and col < h)                                                     context rows see all history
#pragma once
#include <cuda_runtime.h>

namespace hstu {

struct SeqMask {
  int L, seqlen_c, seqlen_h, G, wl, wr;
  bool has_t, has_c;
  // every (row, col) of rows [r0, r1] x cols [c0, c1] valid?  (conservative: false => the interval mask is applied)
  __device__ __forceinline__ bool tile_full(int r0, int r1, int c0, int c1) const {
    if (c1 >= L || r1 >= L) return false;
    if (wr >= 0 && c1 > r0 + wr) return false;
    if (wl >= 0 && c0 < r1 - wl) return false;
    if (has_t && r1 >= seqlen_h && c1 >= seqlen_h) return false;
    return true;
  }
};

// x is valid iff (unsigned)(x - a0) < a_len  ||  (unsigned)(x - b0) < b_len
struct Intervals {
  int a0; unsigned a_len; int b0; unsigned b_len;
  __device__ __forceinline__ bool has(int x) const { return (unsigned)(x - a0) < a_len || (unsigned)(x - b0) < b_len; }
};
__device__ __forceinline__ unsigned span(int lo, int hi) { return hi > lo ? (unsigned)(hi - lo) : 0u; }

// keys a fixed query `row` may attend
__device__ __forceinline__ Intervals cols_of_row(const SeqMask& m, int row) {
  Intervals iv{0, 0u, 0, 0u};
  if (row >= m.L) return iv;
  int lo = m.wl >= 0 ? max(0, row - m.wl) : 0;
  int hi = m.wr >= 0 ? min(m.L, row + m.wr + 1) : m.L;
  if (m.has_c && row < m.seqlen_c) { lo = 0; hi = max(hi, m.seqlen_h); }
  if (m.has_t && row >= m.seqlen_h) {
    const int ex_lo = m.seqlen_h, ex_hi = m.seqlen_h + ((row - m.seqlen_h) / m.G) * m.G;
    iv.a0 = lo; iv.a_len = span(lo, min(hi, ex_lo));
    iv.b0 = max(lo, ex_hi); iv.b_len = span(iv.b0, hi);
  } else {
    iv.a0 = lo; iv.a_len = span(lo, hi);
  }
  return iv;
}
// queries that may attend a fixed key `col`
__device__ __forceinline__ Intervals rows_of_col(const SeqMask& m, int col) {
  Intervals iv{0, 0u, 0, 0u};
  if (col >= m.L) return iv;
  int lo = m.wr >= 0 ? max(0, col - m.wr) : 0;
  int hi = m.wl >= 0 ? min(m.L, col + m.wl + 1) : m.L;
  if (m.has_t && col >= m.seqlen_h) hi = min(hi, m.seqlen_h + ((col - m.seqlen_h) / m.G + 1) * m.G);   // only its own target group
  iv.a0 = lo; iv.a_len = span(lo, hi);
  if (m.has_c && col < m.seqlen_h) { iv.b0 = 0; iv.b_len = (unsigned)m.seqlen_c; }
  return iv;
}

}  // namespace hstu
