"""ORACLE — TEST INFRASTRUCTURE ONLY (see oracle/dynamicemb_oracle.c header).

numpy/ctypes front-end of the CPU restatement of the reference DynamicEmb path.  Nothing in the
product imports this; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference leg do.  Reference paths are relative to /root/reference/corelib/dynamicemb/.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "_build", "liboracle.so")


def build():
    srcs = [os.path.join(_HERE, f) for f in sorted(os.listdir(_HERE)) if f.endswith(".c")]
    os.makedirs(os.path.dirname(_LIB), exist_ok=True)
    if not os.path.exists(_LIB) or any(os.path.getmtime(s) > os.path.getmtime(_LIB) for s in srcs):
        subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-fno-fast-math", "-ffp-contract=off", "-o", _LIB] + srcs + ["-lm"])
    return _LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        _lib.orc_hash.restype = ctypes.c_int64
        _lib.orc_hash.argtypes = [ctypes.c_uint64]
        _lib.orc_fmix64.restype = ctypes.c_uint64
        _lib.orc_fmix64.argtypes = [ctypes.c_uint64]
        _lib.orc_empty_digest.restype = ctypes.c_uint8
        _lib.orc_insert.restype = ctypes.c_int64
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def hash63(key: int) -> int:
    return lib().orc_hash(ctypes.c_uint64(key & 0xFFFFFFFFFFFFFFFF))


def fmix64(key: int) -> int:
    return lib().orc_fmix64(ctypes.c_uint64(key & 0xFFFFFFFFFFFFFFFF))


EMPTY_KEY = 0xFFFFFFFFFFFFFFFF
RECLAIM_KEY = 0xFFFFFFFFFFFFFFFE
LOCKED_KEY = 0xFFFFFFFFFFFFFFFD


class OracleTable:
    """Mirror of scored_hashtable.py LinearBucketTable on host memory (single score spec)."""

    def __init__(self, capacities, bucket_capacity=128, num_scores=1):
        C = ((bucket_capacity + 15) // 16) * 16          # scored_hashtable.py:362-376
        self.C, self.ns = C, num_scores
        nbs = [(c + C - 1) // C for c in capacities]
        self.bkt_off = np.zeros(len(nbs) + 1, dtype=np.int64)
        self.bkt_off[1:] = np.cumsum(nbs)
        self.num_buckets = int(self.bkt_off[-1])
        self.storage = np.zeros(self.num_buckets * C * (9 + 8 * num_scores), dtype=np.uint8)
        self.bucket_sizes = np.zeros(self.num_buckets, dtype=np.int32)
        self.counter = np.zeros(self.num_buckets * C, dtype=np.int32)
        lib().orc_table_init(_p(self.storage), ctypes.c_int64(self.num_buckets), ctypes.c_int64(C), ctypes.c_int64(num_scores))

    # strided field views (table.cu:22-65 table_partition)
    def keys_view(self):
        bb = self.C * (9 + 8 * self.ns)
        return np.lib.stride_tricks.as_strided(self.storage.view(np.uint64), (self.num_buckets, self.C), (bb, 8))

    def digests_view(self):
        bb = self.C * (9 + 8 * self.ns)
        return np.lib.stride_tricks.as_strided(self.storage[8 * self.C:], (self.num_buckets, self.C), (bb, 1))

    def scores_view(self):
        bb = self.C * (9 + 8 * self.ns)
        base = self.storage[9 * self.C:]
        return np.lib.stride_tricks.as_strided(base.view(np.uint8), (self.num_buckets, self.C * self.ns * 8), (bb, 1)).copy().view(np.uint64).reshape(
            self.num_buckets, self.C, self.ns)

    def lookup(self, keys, table_ids=None, policy=0, score_in=None, timer=0):
        keys = np.ascontiguousarray(keys).view(np.uint64)
        n = keys.size
        founds = np.zeros(n, dtype=np.uint8)
        idx = np.zeros(n, dtype=np.int64)
        so = np.zeros(n, dtype=np.int64)
        tids = None if table_ids is None else np.ascontiguousarray(table_ids, dtype=np.int64)
        si = None if score_in is None else np.ascontiguousarray(score_in).view(np.uint64)
        lib().orc_lookup(_p(self.storage), _p(self.bkt_off), ctypes.c_int64(self.C), ctypes.c_int64(self.ns), ctypes.c_int64(n), _p(keys), _p(tids),
                         ctypes.c_int(policy), _p(si), ctypes.c_uint64(timer), _p(founds), _p(idx), _p(so))
        return so, founds.astype(bool), idx

    def insert(self, keys, table_ids=None, policy=1, score_in=None, timer=0, deterministic=True, use_counter=True):
        keys = np.ascontiguousarray(keys).view(np.uint64)
        n = keys.size
        res = np.zeros(n, dtype=np.uint8)
        idx = np.zeros(n, dtype=np.int64)
        so = np.zeros(n, dtype=np.int64)
        evk = np.zeros(max(n, 1), dtype=np.uint64)
        evs = np.zeros(max(n, 1), dtype=np.int64)
        evi = np.zeros(max(n, 1), dtype=np.int64)
        evt = np.zeros(max(n, 1), dtype=np.int64)
        tids = None if table_ids is None else np.ascontiguousarray(table_ids, dtype=np.int64)
        si = None if score_in is None else np.ascontiguousarray(score_in).view(np.uint64)
        nev = lib().orc_insert(_p(self.storage), _p(self.bkt_off), ctypes.c_int64(self.C), ctypes.c_int64(self.ns), _p(self.bucket_sizes),
                               ctypes.c_int64(n), _p(keys), _p(tids), ctypes.c_int(policy), _p(si), ctypes.c_uint64(timer),
                               _p(self.counter) if use_counter else None, ctypes.c_int(1 if deterministic else 0), _p(res), _p(idx), _p(so),
                               _p(evk), _p(evs), _p(evi), _p(evt))
        return idx, res, so, (evk[:nev], evi[:nev], evs[:nev], evt[:nev])

    def erase(self, keys, table_ids=None):
        keys = np.ascontiguousarray(keys).view(np.uint64)
        idx = np.zeros(keys.size, dtype=np.int64)
        tids = None if table_ids is None else np.ascontiguousarray(table_ids, dtype=np.int64)
        lib().orc_erase(_p(self.storage), _p(self.bkt_off), ctypes.c_int64(self.C), ctypes.c_int64(self.ns), _p(self.bucket_sizes),
                        ctypes.c_int64(keys.size), _p(keys), _p(tids), _p(idx))
        return idx


# ---- dedup: unique_op.cu:484 semantics (per-table distinct keys + inverse); our order = first occurrence ----
def segmented_unique(keys, table_range):
    keys = np.asarray(keys)
    T = len(table_range) - 1
    uk, inv, offs = [], np.zeros(keys.size, dtype=np.int64), [0]
    for t in range(T):
        seen = {}
        for i in range(int(table_range[t]), int(table_range[t + 1])):
            k = int(keys[i])
            if k not in seen:
                seen[k] = len(uk)
                uk.append(k)
            inv[i] = seen[k]
        offs.append(len(uk))
    return np.array(uk, dtype=keys.dtype), inv, np.array(offs, dtype=np.int64)


# ---- rows ------------------------------------------------------------------------------------------
def gather_rows(values, D, slots):
    values = np.ascontiguousarray(values, dtype=np.float32)
    slots = np.ascontiguousarray(slots, dtype=np.int64)
    out = np.zeros((slots.size, D), dtype=np.float32)
    lib().orc_gather_rows(_p(values), ctypes.c_int64(values.shape[1]), ctypes.c_int64(D), ctypes.c_int64(slots.size), _p(slots), _p(out))
    return out


def pool_rows(values, D, offsets, slots, combiner, B, F):
    """ids feature-major (bag f*B+b); out[b, f*D:(f+1)*D]  (lookup_forward.cu:53-59)."""
    values = np.ascontiguousarray(values, dtype=np.float32)
    slots = np.ascontiguousarray(slots, dtype=np.int64)
    offsets = np.ascontiguousarray(offsets, dtype=np.int64)
    tmp = np.zeros((F * B, D), dtype=np.float32)
    lib().orc_pool_rows(_p(values), ctypes.c_int64(values.shape[1]), ctypes.c_int64(D), ctypes.c_int64(F * B), _p(offsets), _p(slots),
                        ctypes.c_int(combiner), _p(tmp))
    return np.ascontiguousarray(tmp.reshape(F, B, D).transpose(1, 0, 2).reshape(B, F * D))


def reduce_grads(inverse, grad_rows, scales, num_unique, D, tile=32, window=32):
    """Per-unique gradient sums in the product's documented order (demb_rows.cu backward_tiles / _windows / _spans kernels):
    stable sort by unique idx.  Level 1: inside each 32-row TILE of the sorted list rows are added in order.  A segment that
    spans tiles sums its per-tile partials in order up to the end of the 32-tile WINDOW (1024 rows) its first tile lies in;
    every later window contributes the ordered sum of the segment's partials inside that window; the window sums are added in
    order.  fp32, no FMA contraction.  grad_rows[i]: gradient row of id i ([n, D]); scales[i]: MEAN scale or 1."""
    inverse = np.asarray(inverse, dtype=np.int64)
    order = np.argsort(inverse, kind="stable")
    g = (np.asarray(grad_rows, dtype=np.float32) * np.asarray(scales, dtype=np.float32)[:, None]).astype(np.float32)
    out = np.zeros((num_unique, D), dtype=np.float32)
    n = inverse.size
    sk = inverse[order]
    q = 0
    f32 = np.float32
    while q < n:
        u = sk[q]
        e = q
        while e < n and sk[e] == u:
            e += 1
        # level-1 partials, keyed by tile index
        parts = []
        p = q
        while p < e:
            t = p // tile
            tile_end = min(e, (t + 1) * tile)
            part = np.zeros(D, dtype=f32)
            for r in range(p, tile_end):
                part = (part + g[order[r]]).astype(f32)
            parts.append((t, part))
            p = tile_end
        w0 = parts[0][0] // window
        total = parts[0][1]
        i = 1
        while i < len(parts) and parts[i][0] // window == w0:
            total = (total + parts[i][1]).astype(f32)
            i += 1
        while i < len(parts):
            w = parts[i][0] // window
            wsum = np.zeros(D, dtype=f32)
            while i < len(parts) and parts[i][0] // window == w:
                wsum = (wsum + parts[i][1]).astype(f32)
                i += 1
            total = (total + wsum).astype(f32)
        out[u] = total
        q = e
    return out


def optimizer_update(values, D, rows, grads, opt, lr, eps=1e-8, beta1=0.9, beta2=0.999, weight_decay=0.0, step=1):
    """optimizer_kernel.cuh:41-404 in IEEE fp32.  values[row] = [w(D) | state]. opt: 'sgd','adagrad','adam','rowwise_adagrad'."""
    f = np.float32
    lr, eps, beta1, beta2, weight_decay = f(lr), f(eps), f(beta1), f(beta2), f(weight_decay)
    for i, r in enumerate(rows):
        if r < 0:
            continue
        g = grads[i].astype(f)
        w = values[r, :D]
        if opt == "sgd":
            w -= lr * g
        elif opt == "adagrad":
            s = values[r, D:2 * D]
            s += g * g
            w -= lr * g / (np.sqrt(s) + eps)
        elif opt == "adam":
            m, v = values[r, D:2 * D], values[r, 2 * D:3 * D]
            m[:] = beta1 * m + (f(1) - beta1) * g
            v[:] = beta2 * v + (f(1) - beta2) * g * g
            bc1 = f(1) - f(np.power(np.float32(beta1), np.float32(step)))
            bc2 = f(1) - f(np.power(np.float32(beta2), np.float32(step)))
            w -= lr * ((m / bc1) / (np.sqrt(v / bc2) + eps) + weight_decay * w)
        elif opt == "rowwise_adagrad":
            acc = values[r, D] + f(np.sum((g * g).astype(f), dtype=f) / f(D))
            values[r, D] = acc
            w -= (lr / (np.sqrt(acc) + eps)) * g
        else:
            raise ValueError(opt)
    return values


def dest_rank(ids, mode, W, blk=1):
    """sparse_block_bucketize_features.cu:254-259"""
    ids = np.ascontiguousarray(ids, dtype=np.int64)
    r = np.zeros(ids.size, dtype=np.int64)
    lib().orc_dest_rank(ctypes.c_int64(ids.size), _p(ids), ctypes.c_int({"continuous": 0, "roundrobin": 1, "hash_roundrobin": 2}[mode]),
                        ctypes.c_int64(W), ctypes.c_int64(blk), _p(r), None)
    return r


def block_bucketize(lengths, ids, B, W, block_sizes, dist_types):
    """Stable partition of every (feature,sample) slot's ids by destination rank (kernel1/kernel2, :218-360)."""
    lengths = np.asarray(lengths, dtype=np.int64)
    ids = np.asarray(ids, dtype=np.int64)
    S = lengths.size
    offs = np.concatenate([[0], np.cumsum(lengths)])
    new_len = np.zeros(W * S, dtype=np.int64)
    dest, nid = np.zeros(ids.size, dtype=np.int64), np.zeros(ids.size, dtype=np.int64)
    for s in range(S):
        f = s // B
        blk, dt = int(block_sizes[f]), int(dist_types[f])
        for i in range(int(offs[s]), int(offs[s + 1])):
            u = int(ids[i]) & 0xFFFFFFFFFFFFFFFF
            if dt == 1:
                p, ni = u % W, u
            elif dt == 2:
                p, ni = fmix64(u) % W, u
            else:
                p, ni = (u // blk, u % blk) if u < blk * W else (u % W, u // W)
            dest[i], nid[i] = p, ni if ni < (1 << 63) else ni - (1 << 64)
            new_len[p * S + s] += 1
    new_off = np.concatenate([[0], np.cumsum(new_len)])
    cursor = new_off[:-1].copy()
    new_ids = np.zeros(ids.size, dtype=np.int64)
    perm = np.zeros(ids.size, dtype=np.int64)
    for s in range(S):
        for i in range(int(offs[s]), int(offs[s + 1])):
            c = dest[i] * S + s
            new_ids[cursor[c]] = nid[i]
            perm[i] = cursor[c]
            cursor[c] += 1
    return new_len, new_ids, perm
