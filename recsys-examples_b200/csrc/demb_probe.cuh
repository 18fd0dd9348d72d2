// Warp-tile hash probe with shared-memory staging — the lookup path of the table for buckets of 128 slots.
//
// The reference (and our round-1 kernels) probe one key per THREAD: 16 digests per dependent step from the key's start position, a
// dependent 8-byte key load per digest match, up to 8 steps when the bucket is full (kernels.cuh:83-187, types.cuh:309-396) — a chain of
// up to ~10 dependent memory round trips executed in lock-step by the 32 lanes of a warp (ncu: 89 of 100 issue slots stalled on the
// long scoreboard).  Here a warp probes a TILE of 32 keys:
//   1. the whole 128-byte digest line of every key's bucket is fetched with coalesced 128-bit loads — eight lanes per line, four lines per
//      warp-wide load, eight loads per tile — a full pipeline stage before it is needed (a first version staged the lines in shared
//      memory with one cp.async.bulk per line: correct, but one more TMA operation per id, and the per-SM rate of small bulk copies, not
//      bandwidth, was the limit: 0.24 ms against 0.14 ms for the gather alone);
//   2. the line of a key is scanned by the EIGHT lanes that loaded it (16 digests each), four keys per step; every lane that sees a digest
//      match loads that slot's key straight away — all candidate key loads of the tile are independent and in flight together — and the
//      lane whose candidate equals the key publishes the slot through the warp's shared memory (the per-bucket result staging).
// Result parity: a key is unique in its bucket and (because erase leaves Reclaim, not Empty, behind and insert takes the first Empty in
// probe order) never lies behind an Empty slot of its probe sequence, so "the slot whose digest and key match, anywhere in the bucket"
// is exactly the slot the reference's ordered probe returns, and "no such slot" is exactly its not-found.
#pragma once
#include "demb_common.cuh"
#include "sm100_ptx.cuh"

namespace demb {

constexpr int kProbeC = 128;                       // bucket capacity this path is specialised for

// per-lane description of one key of a tile
struct ProbeKey {
  uint64_t key;
  int64_t bucket;        // global bucket
  int64_t slot_base;     // (bucket - first bucket of its table) * C: table-local slot of position 0
  int32_t tid;
  bool valid;            // key legal and table non-empty
  uint32_t want4;        // the key's digest in all four bytes (kept so that a later pipeline stage does not hash again)
};

// bucket = (h % (nb * 128)) / 128 = (h >> 7) % nb.  A 64-bit modulo by a run-time value is ~150 dependent instructions on this GPU (no
// hardware divider) — more than the rest of a probe put together, and the fused lookup kernel is ALU-latency-bound on its 12 warps
// (ncu: 53 M warp instructions against 13 M for the plain gather).  The table's bucket count is invariant over long runs of ids (ids are
// grouped by table), so each thread caches (table id, first bucket, bucket count, magic = floor((2^64 - 1) / nb)) and reduces with one
// 64-bit multiply-high + at most two corrections; the cache is refilled (one real division) only when the table id changes.
struct TableCache { int32_t tid; int64_t bb; uint64_t nb; uint64_t magic; };
__device__ __forceinline__ TableCache empty_table_cache() { return TableCache{-1, 0, 0, 0}; }

__device__ __forceinline__ ProbeKey make_probe_key(const Table& t, uint64_t key, int tid, TableCache& tc) {
  ProbeKey p{key, 0, 0, tid, false, 0u};
  if (key_is_valid(key)) {
    if (tc.tid != tid) {
      tc.tid = tid;
      tc.bb = t.bkt_off[tid];
      tc.nb = (uint64_t)(t.bkt_off[tid + 1] - tc.bb);
      tc.magic = tc.nb ? 0xFFFFFFFFFFFFFFFFull / tc.nb : 0ull;
    }
    if (tc.nb > 0) {
      const int64_t h = hash63(key);
      p.want4 = (uint32_t)digest_of(h) * 0x01010101u;
      const uint64_t x = (uint64_t)h >> 7;                            // C == 128
      uint64_t r = x - __umul64hi(x, tc.magic) * tc.nb;              // quotient estimate is low by at most 2
      if (r >= tc.nb) r -= tc.nb;
      if (r >= tc.nb) r -= tc.nb;
      p.bucket = tc.bb + (int64_t)r; p.slot_base = (int64_t)r * kProbeC; p.valid = true;
    }
  }
  return p;
}

// The digest lines of a tile, in registers: lane l holds, for step j, the 16-byte chunk (l & 7) of the line of key (l >> 3) + 4 j — the
// chunk it will scan.  One warp-wide 128-bit load covers the four 128-byte lines of four keys, fully coalesced (8 loads per tile);
// issued one pipeline stage ahead of the scan.
struct DigRegs { uint4 d[8]; };

__device__ __forceinline__ uint4 ld_nc_u4(const void* p) {
  uint4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}

// all 32 lanes; lanes without a key pass valid=false
__device__ __forceinline__ void tile_load_digests(const Table& t, const ProbeKey& p, DigRegs& r, int lane) {
  const uint64_t my_line = p.valid ? reinterpret_cast<uint64_t>(t.digests(t.bucket(p.bucket))) : 0ull;
  const int c = lane & 7;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const uint64_t line = __shfl_sync(0xffffffffu, my_line, (lane >> 3) + 4 * j);
    r.d[j] = line ? ld_nc_u4(reinterpret_cast<const uint8_t*>(line) + c * 16) : make_uint4(0u, 0u, 0u, 0u);
  }
}

// 16-bit mask of the bytes of a 16-byte chunk that equal the byte replicated in w4 (bit b = byte b)
__device__ __forceinline__ uint32_t match16(const uint4& d, uint32_t w4) {
  const uint32_t m0 = __vcmpeq4(d.x, w4) & 0x01010101u, m1 = __vcmpeq4(d.y, w4) & 0x01010101u;
  const uint32_t m2 = __vcmpeq4(d.z, w4) & 0x01010101u, m3 = __vcmpeq4(d.w, w4) & 0x01010101u;
  // (m * 0x00204081) >> 21 gathers bits 0, 8, 16, 24 into bits 0..3 (the partial products do not collide)
  return ((m0 * 0x00204081u) >> 21 & 0xFu) | (((m1 * 0x00204081u) >> 21 & 0xFu) << 4) | (((m2 * 0x00204081u) >> 21 & 0xFu) << 8) |
         (((m3 * 0x00204081u) >> 21 & 0xFu) << 12);
}

// The probe in two halves, so that a kernel can put other work (and its one long wait) between them:
//   tile_probe_issue   digest scan -> per-lane candidate masks; ISSUES the first candidate key load of every (lane, step) — nothing waits
//   tile_probe_finish  compares the keys that have arrived meanwhile, slow path for second candidates, publishes the slots
struct ProbeCand { uint64_t key[8]; uint32_t mask[8]; };

__device__ __forceinline__ uint64_t ld_u64_volatile_nc(const uint64_t* p) {   // issued where written, not sunk to the use
  uint64_t v;
  asm volatile("ld.global.nc.L1::no_allocate.u64 %0, [%1];" : "=l"(v) : "l"(p));
  return v;
}

__device__ __forceinline__ void tile_probe_issue(const Table& t, const ProbeKey& p, const DigRegs& dig, ProbeCand& cand, int lane) {
  const uint32_t want = (uint32_t)digest_of(hash63(p.key)) * 0x01010101u;
  const uint64_t my_keys = p.valid ? reinterpret_cast<uint64_t>(t.keys(t.bucket(p.bucket))) : 0ull;
  const int c = lane & 7;                          // 16-byte chunk of the line this lane scans
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int k = (lane >> 3) + 4 * j;             // key of the tile this lane works for in step j
    const uint32_t w4 = __shfl_sync(0xffffffffu, want, k);
    const uint64_t* keys_k = reinterpret_cast<const uint64_t*>(__shfl_sync(0xffffffffu, my_keys, k));
    cand.mask[j] = 0; cand.key[j] = 0;
    if (keys_k) {
      cand.mask[j] = match16(dig.d[j], w4);
      if (cand.mask[j]) cand.key[j] = ld_u64_volatile_nc(keys_k + c * 16 + __ffs(cand.mask[j]) - 1);
    }
  }
}

__device__ __forceinline__ int tile_probe_finish(const Table& t, const ProbeKey& p, const ProbeCand& cand, int* slot_sm, int lane) {
  slot_sm[lane] = -1;
  __syncwarp();
  const uint64_t my_keys = p.valid ? reinterpret_cast<uint64_t>(t.keys(t.bucket(p.bucket))) : 0ull;
  const int c = lane & 7;
  bool more = false;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int k = (lane >> 3) + 4 * j;
    const uint64_t key_k = __shfl_sync(0xffffffffu, p.key, k);
    if (cand.mask[j]) {
      if (cand.key[j] == key_k) slot_sm[k] = c * 16 + __ffs(cand.mask[j]) - 1;
      else if ((cand.mask[j] & (cand.mask[j] - 1)) && cand.key[j] != kEmptyKey) more = true;
    }
  }
  if (__any_sync(0xffffffffu, more)) {               // slow path: remaining candidates of a chunk, one dependent load each
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int k = (lane >> 3) + 4 * j;
      const uint64_t key_k = __shfl_sync(0xffffffffu, p.key, k);
      const uint64_t* keys_k = reinterpret_cast<const uint64_t*>(__shfl_sync(0xffffffffu, my_keys, k));
      uint32_t m = (cand.key[j] == key_k || cand.key[j] == kEmptyKey) ? 0u : (cand.mask[j] & (cand.mask[j] - 1));
      while (m) {
        const int pos = c * 16 + __ffs(m) - 1;
        m &= m - 1;
        const uint64_t kk = keys_k[pos];
        if (kk == key_k) { slot_sm[k] = pos; break; }
        if (kk == kEmptyKey) break;
      }
    }
  }
  __syncwarp();
  return slot_sm[lane];
}

// ---- second-generation tile probe (same results, fewer instructions and no exposed dependent load on the common path).  Measured on the
// first generation (ncu source view, per tile of 32 ids): ~300 of ~1300 warp instructions were match16 (__vcmpeq4 is emulated), ~350
// the per-step shuffles / compares / publishes of its EIGHT steps, and three tiles in four took the slow path (a second digest match
// inside a 16-slot chunk, first candidate not the key) — a dependent key load with its full latency exposed.  Here:
//   * FOUR lanes per line (two 16-byte pieces each: slots [16c, 16c+16) and [64+16c, 64+16c+16)), eight keys per step, FOUR steps;
//     every warp-wide load still covers whole 32-byte sectors;
//   * matches as a sparse 32-bit word straight from the SWAR zero-byte test ((y - 0x01..) & ~y & 0x80..: never misses a match; may flag
//     the byte above one, which only costs a candidate whose key then does not compare equal) — no per-word gather multiply;
//   * the first TWO candidates of every (lane, step) are loaded at once, so the slow path needs three digest matches in 32 slots with
//     the key not among the first two (~1 tile in 10, against 3 in 4).
// Candidates are visited in mask-bit order, not probe order; that is immaterial except for the stop-at-Empty rule, which only ever
// applies to keys whose digest equals the empty digest (1 key in 256): those take the ordered walk in the slow path.
struct DigRegs4 { uint4 d[4][2]; };
struct ProbeCand4 { uint64_t key[4][2]; uint32_t mask[4]; uint32_t pos[4]; };   // pos: the two candidate positions, bytes 0 and 1

__device__ __forceinline__ uint32_t zero_bytes(uint32_t y) { return (y - 0x01010101u) & ~y & 0x80808080u; }
// bit 8 b + w  <=>  byte b of word w of the 16-byte piece equals the byte replicated in w4  (slot 4 w + b of the piece)
__device__ __forceinline__ uint32_t sparse16(const uint4& d, uint32_t w4) {
  return (zero_bytes(d.x ^ w4) >> 7) | (zero_bytes(d.y ^ w4) >> 6) | (zero_bytes(d.z ^ w4) >> 5) | (zero_bytes(d.w ^ w4) >> 4);
}
// mask bit i of lane-quarter c -> position in the bucket: piece (i >> 2) & 1, word i & 3, byte i >> 3
__device__ __forceinline__ int sparse_pos(int i, int c) { return ((i >> 2) & 1) * 64 + c * 16 + (i & 3) * 4 + (i >> 3); }

__device__ __forceinline__ void tile_load_digests4(const Table& t, const ProbeKey& p, DigRegs4& r, int lane) {
  const uint64_t my_line = p.valid ? reinterpret_cast<uint64_t>(t.digests(t.bucket(p.bucket))) : 0ull;
  const int c = lane & 3;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const uint64_t line = __shfl_sync(0xffffffffu, my_line, (lane >> 2) + 8 * j);
    r.d[j][0] = line ? ld_nc_u4(reinterpret_cast<const uint8_t*>(line) + c * 16) : make_uint4(0u, 0u, 0u, 0u);
    r.d[j][1] = line ? ld_nc_u4(reinterpret_cast<const uint8_t*>(line) + 64 + c * 16) : make_uint4(0u, 0u, 0u, 0u);
  }
}

__device__ __forceinline__ void tile_probe_issue4(const Table& t, const ProbeKey& p, const DigRegs4& dig, ProbeCand4& cand, int lane) {
  const uint32_t want = p.want4;
  const uint64_t my_keys = p.valid ? reinterpret_cast<uint64_t>(t.keys(t.bucket(p.bucket))) : 0ull;
  const int c = lane & 3;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int k = (lane >> 2) + 8 * j;             // key of the tile this lane works for in step j
    const uint32_t w4 = __shfl_sync(0xffffffffu, want, k);
    const uint64_t* keys_k = reinterpret_cast<const uint64_t*>(__shfl_sync(0xffffffffu, my_keys, k));
    uint32_t u = 0;
    cand.key[j][0] = 0; cand.key[j][1] = 0; cand.pos[j] = 0;
    if (keys_k) {
      u = sparse16(dig.d[j][0], w4) | (sparse16(dig.d[j][1], w4) << 4);
      if (u) {
        const int p0 = sparse_pos(__ffs(u) - 1, c);
        cand.key[j][0] = ld_u64_volatile_nc(keys_k + p0);
        cand.pos[j] = (uint32_t)p0;
        const uint32_t u2 = u & (u - 1);
        if (u2) {
          const int p1 = sparse_pos(__ffs(u2) - 1, c);
          cand.key[j][1] = ld_u64_volatile_nc(keys_k + p1);
          cand.pos[j] |= (uint32_t)p1 << 8;
        }
      }
    }
    cand.mask[j] = u;
  }
}

__device__ __forceinline__ int tile_probe_finish4(const Table& t, const ProbeKey& p, const ProbeCand4& cand, int* slot_sm, int lane) {
  slot_sm[lane] = -1;
  __syncwarp();
  const int c = lane & 3;
  bool more = false;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int k = (lane >> 2) + 8 * j;
    const uint64_t key_k = __shfl_sync(0xffffffffu, p.key, k);
    const uint32_t u = cand.mask[j];
    if (u) {
      if (cand.key[j][0] == key_k) slot_sm[k] = (int)(cand.pos[j] & 0xFFu);
      else {
        const uint32_t u2 = u & (u - 1);
        if (u2) {
          if (cand.key[j][1] == key_k) slot_sm[k] = (int)(cand.pos[j] >> 8);
          else if (u2 & (u2 - 1)) more = true;
        }
      }
    }
  }
  if (__any_sync(0xffffffffu, more)) {               // slow path: third and later candidates of a (lane, step), one dependent load each
    const uint32_t want = p.want4;
    const uint32_t emp4 = (uint32_t)empty_digest() * 0x01010101u;
    const uint64_t my_keys = p.valid ? reinterpret_cast<uint64_t>(t.keys(t.bucket(p.bucket))) : 0ull;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k = (lane >> 2) + 8 * j;
      const uint64_t key_k = __shfl_sync(0xffffffffu, p.key, k);
      const uint32_t w4 = __shfl_sync(0xffffffffu, want, k);
      const uint64_t* keys_k = reinterpret_cast<const uint64_t*>(__shfl_sync(0xffffffffu, my_keys, k));
      const uint32_t u = cand.mask[j], u2 = u & (u - 1), u3 = u2 & (u2 - 1);
      if (u3 == 0 || cand.key[j][0] == key_k || cand.key[j][1] == key_k) continue;
      if (w4 != emp4) {                              // no candidate can be an Empty slot: any order, until the key turns up
        uint32_t m = u3;
        while (m) {
          const int pos = sparse_pos(__ffs(m) - 1, c);
          m &= m - 1;
          if (keys_k[pos] == key_k) { slot_sm[k] = pos; break; }
        }
      } else {                                       // digest == empty digest: ascending position inside each 16-slot piece, stop at Empty
        bool done = false;
        for (int piece = 0; piece < 2 && !done; ++piece)
          for (int q = 0; q < 16; ++q) {             // q = 4 w + b
            const int i = (q & 3) * 8 + piece * 4 + (q >> 2);
            if (!((u >> i) & 1u)) continue;
            const int pos = piece * 64 + c * 16 + q;
            const uint64_t kk = keys_k[pos];
            if (kk == key_k) { slot_sm[k] = pos; done = true; break; }
            if (kk == kEmptyKey) break;
          }
      }
    }
  }
  __syncwarp();
  return slot_sm[lane];
}

// the two generations behind one name, for kernels templated on the probe
template <int GEN> struct TileProbe;
template <> struct TileProbe<1> {
  using Dig = DigRegs; using Cand = ProbeCand;
  static __device__ __forceinline__ void load(const Table& t, const ProbeKey& p, Dig& r, int lane) { tile_load_digests(t, p, r, lane); }
  static __device__ __forceinline__ void issue(const Table& t, const ProbeKey& p, const Dig& d, Cand& c, int lane) { tile_probe_issue(t, p, d, c, lane); }
  static __device__ __forceinline__ int finish(const Table& t, const ProbeKey& p, const Cand& c, int* sm, int lane) { return tile_probe_finish(t, p, c, sm, lane); }
};
template <> struct TileProbe<2> {
  using Dig = DigRegs4; using Cand = ProbeCand4;
  static __device__ __forceinline__ void load(const Table& t, const ProbeKey& p, Dig& r, int lane) { tile_load_digests4(t, p, r, lane); }
  static __device__ __forceinline__ void issue(const Table& t, const ProbeKey& p, const Dig& d, Cand& c, int lane) { tile_probe_issue4(t, p, d, c, lane); }
  static __device__ __forceinline__ int finish(const Table& t, const ProbeKey& p, const Cand& c, int* sm, int lane) { return tile_probe_finish4(t, p, c, sm, lane); }
};

// The same probe for warps that must stay SMALL (64 registers: 32 warps per SM): nothing is carried between tiles; the eight steps
// run as two halves of four (digest chunks -> masks -> first candidate keys -> compare), and the pointers are re-derived by shuffles
// instead of being kept.  Two dependent round trips per half; meant for many resident warps, each behind an L2 prefetch of its lines.
__device__ __forceinline__ int tile_probe_small(const Table& t, const ProbeKey& p, int* slot_sm, int lane) {
  slot_sm[lane] = -1;
  __syncwarp();
  const uint32_t want = (uint32_t)digest_of(hash63(p.key)) * 0x01010101u;
  const uint64_t my_bucket = p.valid ? reinterpret_cast<uint64_t>(t.bucket(p.bucket)) : 0ull;    // keys at +0, digests at +8 C
  const int c = lane & 7;
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    uint4 d[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint64_t bk = __shfl_sync(0xffffffffu, my_bucket, (lane >> 3) + 4 * (4 * h + j));
      d[j] = bk ? ld_nc_u4(reinterpret_cast<const uint8_t*>(bk) + 8 * kProbeC + c * 16) : make_uint4(0u, 0u, 0u, 0u);
    }
    uint32_t m[4]; uint64_t ck[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k = (lane >> 3) + 4 * (4 * h + j);
      const uint32_t w4 = __shfl_sync(0xffffffffu, want, k);
      const uint64_t bk = __shfl_sync(0xffffffffu, my_bucket, k);
      m[j] = bk ? match16(d[j], w4) : 0u;
      ck[j] = 0;
      if (m[j]) ck[j] = ld_u64_volatile_nc(reinterpret_cast<const uint64_t*>(bk) + c * 16 + __ffs(m[j]) - 1);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k = (lane >> 3) + 4 * (4 * h + j);
      const uint64_t key_k = __shfl_sync(0xffffffffu, p.key, k);
      const uint64_t bk = __shfl_sync(0xffffffffu, my_bucket, k);
      if (m[j]) {
        if (ck[j] == key_k) slot_sm[k] = c * 16 + __ffs(m[j]) - 1;
        else if (ck[j] != kEmptyKey) {               // further digest matches inside the same 16 slots (rare): one dependent load each
          uint32_t mm = m[j] & (m[j] - 1);
          while (mm) {
            const int pos = c * 16 + __ffs(mm) - 1;
            mm &= mm - 1;
            const uint64_t kk = reinterpret_cast<const uint64_t*>(bk)[pos];
            if (kk == key_k) { slot_sm[k] = pos; break; }
            if (kk == kEmptyKey) break;
          }
        }
      }
    }
  }
  __syncwarp();
  return slot_sm[lane];
}

// all 32 lanes; returns the position (0..127) of this lane's key in its bucket, or -1.  slot_sm: 32 ints of the warp's shared memory.
__device__ __forceinline__ int tile_probe(const Table& t, const ProbeKey& p, const DigRegs& dig, int* slot_sm, int lane) {
  slot_sm[lane] = -1;
  __syncwarp();
  const uint32_t want = (uint32_t)digest_of(hash63(p.key)) * 0x01010101u;
  const uint64_t my_keys = p.valid ? reinterpret_cast<uint64_t>(t.keys(t.bucket(p.bucket))) : 0ull;
  const int c = lane & 7;                          // 16-byte chunk of the line this lane scans
  uint64_t cand_key[8]; uint32_t mask[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int k = (lane >> 3) + 4 * j;             // key of the tile this lane works for in step j
    const uint32_t w4 = __shfl_sync(0xffffffffu, want, k);
    const uint64_t* keys_k = reinterpret_cast<const uint64_t*>(__shfl_sync(0xffffffffu, my_keys, k));
    mask[j] = 0; cand_key[j] = 0;
    if (keys_k) {
      mask[j] = match16(dig.d[j], w4);
      if (mask[j]) cand_key[j] = keys_k[c * 16 + __ffs(mask[j]) - 1];     // independent loads: every candidate of the tile is in flight together
    }
  }
  bool more = false;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int k = (lane >> 3) + 4 * j;
    const uint64_t key_k = __shfl_sync(0xffffffffu, p.key, k);
    if (mask[j]) {
      if (cand_key[j] == key_k) slot_sm[k] = c * 16 + __ffs(mask[j]) - 1;
      else if (mask[j] & (mask[j] - 1)) more = true;                       // a second digest match inside the same 16 bytes (rare)
    }
  }
  if (__any_sync(0xffffffffu, more)) {               // slow path: remaining candidates of a chunk, one dependent load each
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int k = (lane >> 3) + 4 * j;
      const uint64_t key_k = __shfl_sync(0xffffffffu, p.key, k);
      const uint64_t* keys_k = reinterpret_cast<const uint64_t*>(__shfl_sync(0xffffffffu, my_keys, k));
      // drop the candidate already checked; nothing left to do if it was the key, or an Empty slot: inside a 16-slot group ascending
      // position is probe order, and a key never lies behind an Empty slot of its probe sequence (keys whose digest equals the empty
      // digest would otherwise walk every empty slot of the chunk)
      uint32_t m = (cand_key[j] == key_k || cand_key[j] == kEmptyKey) ? 0u : (mask[j] & (mask[j] - 1));
      while (m) {
        const int pos = c * 16 + __ffs(m) - 1;
        m &= m - 1;
        const uint64_t kk = keys_k[pos];
        if (kk == key_k) { slot_sm[k] = pos; break; }
        if (kk == kEmptyKey) break;
      }
    }
  }
  __syncwarp();
  return slot_sm[lane];
}

}  // namespace demb
