"""Op layer: the names the reference's pybind module `dynamicemb_extensions` exports
(/root/reference/corelib/dynamicemb/src/module_bind.cu:22-43, table_operation/table.cu:94-204),
implemented over the C ABI of librecsys_b200.so.  Tensors in, current CUDA stream, Python
exceptions on error — the same calling convention the reference Python package relies on.
"""
import enum
from typing import List, Optional, Tuple

import ctypes

import torch

from . import _native as N


class ScorePolicy(enum.IntEnum):
    """score.cuh:30-42"""
    CONST = 0
    ASSIGN = 1
    ACCUMULATE = 2
    GLOBAL_TIMER = 3
    LRU_LFU = 4


class InsertResult(enum.IntEnum):
    """types.cuh:52-61"""
    INSERT = 0
    RECLAIM = 1
    ASSIGN = 2
    EVICT = 3
    DUPLICATED = 4
    BUSY = 5
    ILLEGAL = 6
    INIT = 7


class EvictStrategy(enum.IntEnum):
    KLru = 0
    KLfu = 1
    KEpochLru = 2
    KEpochLfu = 3
    KCustomized = 4


class OptimizerType(enum.IntEnum):
    NONE = 0
    SGD = 1
    ADAM = 2
    ADAGRAD = 3
    ROWWISE_ADAGRAD = 4


class InitializerMode(enum.IntEnum):
    NORMAL = 0
    TRUNCATED_NORMAL = 1
    UNIFORM = 2
    DEBUG = 3
    CONSTANT = 4


_OUT_DTYPE = {torch.float32: 0, torch.float16: 1, torch.bfloat16: 2}


def _i64(t: Optional[torch.Tensor]) -> Optional[torch.Tensor]:
    if t is None:
        return None
    if t.dtype != torch.int64:
        t = t.to(torch.int64)
    return t.contiguous()


def _u64_view(t: Optional[torch.Tensor]) -> Optional[torch.Tensor]:
    if t is None:
        return None
    return t.contiguous()


def _num_buckets(storage: torch.Tensor, bucket_capacity: int, num_scores: int) -> int:
    return storage.numel() * storage.element_size() // (bucket_capacity * (9 + 8 * num_scores))


def device_timestamp() -> int:
    import time
    return time.time_ns()


# ---------------------------------------------------------------------------------------------------
# table ops
# ---------------------------------------------------------------------------------------------------
def table_partition(storage: torch.Tensor, dtypes: List[torch.dtype], bucket_capacity: int, num_buckets: int) -> List[torch.Tensor]:
    """table.cu:22-65: per-field strided [num_buckets, bucket_capacity] views of the bucket-SoA storage."""
    sizes = [torch.empty((), dtype=d).element_size() for d in dtypes]
    bucket_bytes = sum(sizes) * bucket_capacity
    if bucket_bytes * num_buckets != storage.numel() * storage.element_size():
        raise RuntimeError("Storage size mismatched with bucket_bytes * num_buckets")
    out, byte_off = [], 0
    flat = storage.view(torch.uint8).reshape(-1)
    for d, sz in zip(dtypes, sizes):
        # bucket_bytes is a multiple of 16*17; every field offset is a multiple of its element size
        v = flat[byte_off * bucket_capacity:].view(d) if (byte_off * bucket_capacity) % sz == 0 else None
        assert v is not None
        out.append(torch.as_strided(v, (num_buckets, bucket_capacity), (bucket_bytes // sz, 1)))
        byte_off += sz
    return out


def table_init(table_storage: torch.Tensor, bucket_capacity: int, num_scores: int = 1) -> None:
    nb = _num_buckets(table_storage, bucket_capacity, num_scores)
    N.check(N.lib.demb_table_init(N.ptr(table_storage), nb, bucket_capacity, num_scores, N.stream()), "table_init")


def table_lookup(table_storage, table_bucket_offsets, bucket_capacity, keys, table_ids, score_input, policy_type,
                 num_scores: int = 1, timestamp: int = 0) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """table.cuh:68 / lookup.cu.  Returns (score_out int64, founds bool, indices int64)."""
    n = keys.numel()
    dev = keys.device
    score_out = torch.empty(n, dtype=torch.int64, device=dev)
    founds = torch.empty(n, dtype=torch.bool, device=dev)
    indices = torch.empty(n, dtype=torch.int64, device=dev)
    if n == 0:
        return score_out, founds, indices
    keys = keys.contiguous()
    table_ids = _i64(table_ids)
    score_input = _u64_view(score_input)
    N.check(N.launch("table_lookup", 1, N.lib.demb_table_lookup, N.ptr(table_storage), N.ptr(table_bucket_offsets), bucket_capacity, num_scores, n, N.ptr(keys),
                                    N.ptr(table_ids), int(policy_type), N.ptr(score_input), int(timestamp), N.ptr(founds), N.ptr(indices),
                                    N.ptr(score_out), N.stream()), "table_lookup")
    return score_out, founds, indices


def _insert_impl(table_storage, table_bucket_offsets, bucket_capacity, bucket_sizes, keys, table_ids, score_input, policy_type, counter,
                 insert_results, score_output, num_scores, want_evicted, timestamp):
    n = keys.numel()
    dev = keys.device
    indices = torch.empty(n, dtype=torch.int64, device=dev)
    ev = None
    if want_evicted:
        ev = (torch.zeros(1, dtype=torch.int64, device=dev), torch.empty(n, dtype=keys.dtype, device=dev),
              torch.empty(n, dtype=torch.int64, device=dev), torch.empty(n, dtype=torch.int64, device=dev),
              torch.empty(n, dtype=torch.int64, device=dev))
    if n == 0:
        return indices, ev
    keys = keys.contiguous()
    table_ids = _i64(table_ids)
    score_input = _u64_view(score_input)
    ws_bytes = N.lib.demb_table_insert_workspace_bytes(n)
    ws = N.workspace(ws_bytes, dev)
    nb_total = bucket_sizes.numel()
    N.check(N.launch("table_insert", 4, N.lib.demb_table_insert,
        N.ptr(table_storage), N.ptr(table_bucket_offsets), bucket_capacity, num_scores, nb_total, N.ptr(bucket_sizes), n, N.ptr(keys),
        N.ptr(table_ids), int(policy_type), N.ptr(score_input), int(timestamp), N.ptr(counter), 1 if keys.dtype == torch.int64 else 0,
        N.ptr(insert_results), N.ptr(indices), N.ptr(score_output),
        N.ptr(ev[0]) if ev else None, N.ptr(ev[1]) if ev else None, N.ptr(ev[2]) if ev else None, N.ptr(ev[3]) if ev else None,
        N.ptr(ev[4]) if ev else None, N.ptr(ws), ws.numel(), N.stream()), "table_insert")
    return indices, ev


def table_insert(table_storage, table_bucket_offsets, bucket_capacity, bucket_sizes, keys, table_ids, score_input, policy_type, counter,
                 insert_results=None, score_output=None, num_scores: int = 1, timestamp: int = 0) -> torch.Tensor:
    """table.cuh:81 / insert.cu.  Keys must be unique; always deterministic (see include/dynamicemb_b200.h)."""
    indices, _ = _insert_impl(table_storage, table_bucket_offsets, bucket_capacity, bucket_sizes, keys, table_ids, score_input, policy_type,
                              counter, insert_results, score_output, num_scores, False, timestamp)
    return indices


def table_insert_and_evict(table_storage, table_bucket_offsets, bucket_capacity, bucket_sizes, keys, table_ids, score_input, policy_type,
                           counter, insert_results=None, score_output=None, num_scores: int = 1, timestamp: int = 0):
    """table.cuh:100 / insert_and_evict.cu:330.  Returns (indices, num_evicted[1] device tensor, evicted_keys,
    evicted_indices, evicted_scores, evicted_table_ids) — buffers sized n, first num_evicted valid."""
    indices, ev = _insert_impl(table_storage, table_bucket_offsets, bucket_capacity, bucket_sizes, keys, table_ids, score_input, policy_type,
                               counter, insert_results, score_output, num_scores, True, timestamp)
    return indices, ev[0], ev[1], ev[3], ev[2], ev[4]


def table_erase(table_storage, table_bucket_offsets, bucket_capacity, bucket_sizes, keys, table_ids, indices=None, num_scores: int = 1) -> None:
    n = keys.numel()
    if n == 0:
        return
    N.check(N.lib.demb_table_erase(N.ptr(table_storage), N.ptr(table_bucket_offsets), bucket_capacity, num_scores, N.ptr(bucket_sizes), n,
                                   N.ptr(keys.contiguous()), N.ptr(_i64(table_ids)), N.ptr(indices), N.stream()), "table_erase")


def table_update_counter_with_layout(counter, slot_indices, delta, table_bucket_offsets, bucket_capacity, total_capacity=None, num_tables=None,
                                     table_ids=None, overflow_output_offsets=None, overflow_bucket_capacity=0) -> None:
    n = slot_indices.numel()
    if n == 0:
        return
    N.check(N.launch("counter_update", 1, N.lib.demb_counter_update, N.ptr(counter), N.ptr(_i64(slot_indices)), N.ptr(_i64(table_ids)), N.ptr(table_bucket_offsets),
                                      bucket_capacity, n, int(delta), N.stream()), "table_update_counter_with_layout")


def table_export_batch(table_storage, bucket_capacity, batch, offset, key_dtype, threshold=None, table_begin=0, num_scores: int = 1,
                       score_index: int = 0):
    """export_batch.cu: returns (counter[1], keys[batch], scores[batch], indices[batch])."""
    dev = table_storage.device
    counter = torch.zeros(1, dtype=torch.int64, device=dev)
    keys = torch.empty(batch, dtype=key_dtype, device=dev)
    scores = torch.empty(batch, dtype=torch.int64, device=dev)
    indices = torch.empty(batch, dtype=torch.int64, device=dev)
    N.check(N.lib.demb_table_export(N.ptr(table_storage), None, bucket_capacity, num_scores, offset, offset + batch, table_begin,
                                    int(threshold) if threshold is not None else 0, 0 if threshold is None else 1, score_index,
                                    N.ptr(counter), N.ptr(keys), N.ptr(scores), N.ptr(indices), N.stream()), "table_export_batch")
    return counter, keys, scores.view(torch.uint64), indices


def fill_i32(t: torch.Tensor, v: int) -> None:
    N.check(N.lib.demb_fill_i32(N.ptr(t), t.numel(), int(v), N.stream()), "fill_i32")


def table_update_counter_n(counter, slot_indices, delta, table_bucket_offsets, bucket_capacity, n_device, table_ids=None) -> None:
    """counter update whose element count lives on the device (fused training path: no host sync)."""
    N.check(N.launch("counter_update", 1, N.lib.demb_counter_update_n, N.ptr(counter), N.ptr(slot_indices), N.ptr(table_ids),
                     N.ptr(table_bucket_offsets), bucket_capacity, N.ptr(n_device), slot_indices.numel(), int(delta), N.stream()),
            "counter_update_n")


def train_prefetch(table_storage, table_bucket_offsets, bucket_capacity, bucket_sizes, ref_counter, bucket_heads, values, emb_dim, row_base,
                   keys, table_range, num_tables, policy, table_scores, timestamp, init_mode, init_params, seed, state_init,
                   freq_in=None, num_scores=1, table_init=None, n_dev=None, unique_scratch=None):
    """Fused dedup + probe + insert/init + pin (demb_train.cu).  Returns (unique_keys[n], reverse[n], unique_table_ids[n], slots[n],
    rows[n], num_unique[1] device) — only the first num_unique entries of the per-unique outputs are meaningful.
    table_init: device tensor from `make_table_init` (one initializer per table), overrides (init_mode, init_params, seed).
    n_dev: device int64 scalar with the real id count (<= keys.numel()).  unique_scratch: persistent dedup scratch (`unique_scratch`)."""
    n = keys.numel()
    dev = keys.device
    uk = torch.empty(n, dtype=keys.dtype, device=dev)
    rev = torch.empty(n, dtype=torch.int64, device=dev)
    utids = torch.empty(n, dtype=torch.int64, device=dev)
    ufreq = torch.empty(n, dtype=torch.int64, device=dev) if int(policy) in (2, 4) else None
    slots = torch.empty(n, dtype=torch.int64, device=dev)
    rows = torch.empty(n, dtype=torch.int64, device=dev)
    nu = torch.zeros(1, dtype=torch.int64, device=dev)
    ws = N.workspace(N.lib.demb_train_prefetch_workspace_bytes(n, num_tables), dev)
    p0, p1, p2, p3 = init_params
    N.check(N.launch("train_prefetch", 8, N.lib.demb_train_prefetch, N.ptr(table_storage), N.ptr(table_bucket_offsets), bucket_capacity, num_scores,
                     N.ptr(bucket_sizes), N.ptr(ref_counter), N.ptr(bucket_heads), N.ptr(values), values.stride(0), emb_dim, N.ptr(row_base), n,
                     N.ptr(n_dev), N.ptr(keys.contiguous()), N.ptr(table_range), num_tables, N.ptr(_i64(freq_in)), int(policy), N.ptr(table_scores), int(timestamp),
                     1 if keys.dtype == torch.int64 else 0, int(init_mode), float(p0), float(p1), float(p2), float(p3), int(seed), N.ptr(table_init),
                     float(state_init), N.ptr(uk), N.ptr(rev), N.ptr(utids), N.ptr(ufreq), N.ptr(slots), N.ptr(rows), N.ptr(nu), N.ptr(unique_scratch),
                     N.ptr(ws), ws.numel(), N.stream()), "train_prefetch")
    return uk, rev, utids, slots, rows, nu


# ---------------------------------------------------------------------------------------------------
# dedup
# ---------------------------------------------------------------------------------------------------
def get_table_range(offsets: torch.Tensor, feature_offsets: torch.Tensor, num_features: Optional[int] = None) -> torch.Tensor:
    """index_calculation.cu:237: table_range[t] = offsets[feature_offsets[t] * B].  `num_features` avoids a device read."""
    T = feature_offsets.numel() - 1
    if num_features is None:
        num_features = int(feature_offsets[-1].item())
    B = (offsets.numel() - 1) // num_features if num_features > 0 else 0
    out = torch.empty(T + 1, dtype=torch.int64, device=offsets.device)
    N.check(N.lib.demb_get_table_range(N.ptr(_i64(offsets)), N.ptr(_i64(feature_offsets)), T, B, N.ptr(out), N.stream()), "get_table_range")
    return out


def unique_scratch(n_max: int, num_tables: int, device) -> torch.Tensor:
    """Persistent dedup scratch for segmented_unique_cuda / train_prefetch calls of up to n_max ids: initialised once here, left clean by
    every call, so the per-call 32 MB initialisation disappears.  Must not be used for anything else."""
    t = torch.empty(N.lib.demb_unique_scratch_bytes(n_max, num_tables), dtype=torch.uint8, device=device)
    N.check(N.lib.demb_unique_scratch_init(N.ptr(t), t.numel(), N.stream()), "unique_scratch_init")
    return t


def segmented_unique_cuda(keys: torch.Tensor, segment_range: Optional[torch.Tensor], num_tables: int,
                          input_frequencies: Optional[torch.Tensor] = None, want_table_ids: bool = False, n_dev=None, scratch=None):
    """unique_op.cu:484.  Returns (num_uniques[1] device, unique_keys[n], reverse_indices[n], table_offsets[T+1], freq[n] or None
    [, unique_table_ids[n]]).  Unique order = first occurrence (deterministic)."""
    n = keys.numel()
    dev = keys.device
    num_unique = torch.zeros(1, dtype=torch.int64, device=dev)
    unique_keys = torch.empty(n, dtype=keys.dtype, device=dev)
    reverse = torch.empty(n, dtype=torch.int64, device=dev)
    table_offsets = torch.zeros(num_tables + 1, dtype=torch.int64, device=dev)
    need_freq = input_frequencies is not None
    freq_in = input_frequencies if (need_freq and input_frequencies.numel() == n and n > 0) else None
    freq_out = torch.empty(n, dtype=torch.int64, device=dev) if need_freq else None
    utids = torch.empty(n, dtype=torch.int64, device=dev) if want_table_ids else None
    ws_bytes = N.lib.demb_segmented_unique_workspace_bytes(n, num_tables)
    ws = N.workspace(ws_bytes, dev)
    N.check(N.launch("segmented_unique", 3, N.lib.demb_segmented_unique, n, N.ptr(n_dev), N.ptr(keys.contiguous()),
                     N.ptr(_i64(segment_range)) if num_tables > 1 else None, num_tables, N.ptr(_i64(freq_in)), N.ptr(unique_keys), N.ptr(reverse),
                     N.ptr(table_offsets), N.ptr(freq_out), N.ptr(utids), N.ptr(num_unique), N.ptr(scratch), N.ptr(ws), ws.numel(), N.stream()),
            "segmented_unique")
    if want_table_ids:
        return num_unique, unique_keys, reverse, table_offsets, freq_out, utids
    return num_unique, unique_keys, reverse, table_offsets, freq_out


def expand_table_ids_cuda(table_offsets: torch.Tensor, num_unique: int) -> torch.Tensor:
    out = torch.empty(num_unique, dtype=torch.int64, device=table_offsets.device)
    N.check(N.lib.demb_expand_table_ids(N.ptr(table_offsets), table_offsets.numel() - 1, num_unique, N.ptr(out), N.stream()), "expand_table_ids")
    return out


# ---------------------------------------------------------------------------------------------------
# rows
# ---------------------------------------------------------------------------------------------------
def lookup_forward(table_storage, table_bucket_offsets, bucket_capacity, values, emb_dim, keys, *, row_base=None, table_range=None,
                   num_tables=1, offsets=None, batch_size=0, num_features=0, combiner=-1, out_dtype=torch.float32, absent_value=0.0,
                   want_founds=False, num_scores=1):
    """Fused probe + gather (+pool): the eval-path forward."""
    n = keys.numel()
    dev = keys.device
    D = emb_dim
    if combiner < 0:
        out = torch.empty(n, D, dtype=out_dtype, device=dev)
    else:
        out = torch.empty(batch_size, num_features * D, dtype=out_dtype, device=dev)
    founds = torch.empty(n, dtype=torch.bool, device=dev) if want_founds else None
    slots = torch.empty(n, dtype=torch.int64, device=dev) if want_founds else None
    N.check(N.launch("lookup_forward", 1, N.lib.demb_lookup_forward, N.ptr(table_storage), N.ptr(table_bucket_offsets), bucket_capacity, num_scores, N.ptr(values),
                                      values.stride(0), D, N.ptr(row_base), n, N.ptr(keys.contiguous()), N.ptr(table_range), num_tables,
                                      N.ptr(_i64(offsets)), batch_size, num_features, combiner, N.ptr(out), _OUT_DTYPE[out_dtype],
                                      float(absent_value), N.ptr(founds), N.ptr(slots), N.stream()), "lookup_forward")
    return (out, founds, slots) if want_founds else out


def gather_forward(values, emb_dim, rows, inverse, n, *, offsets=None, batch_size=0, num_features=0, combiner=-1, out_dtype=torch.float32, n_dev=None,
                   out=None):
    dev = values.device
    if out is None:
        if combiner < 0:
            out = torch.empty(n, emb_dim, dtype=out_dtype, device=dev)
        else:
            out = torch.empty(batch_size, num_features * emb_dim, dtype=out_dtype, device=dev)
    N.check(N.launch("gather_forward", 1, N.lib.demb_gather_forward, N.ptr(values), values.stride(0), emb_dim, n, N.ptr(rows), N.ptr(inverse), N.ptr(_i64(offsets)),
                                      batch_size, num_features, combiner, N.ptr(out), _OUT_DTYPE[out.dtype], N.ptr(n_dev), N.stream()), "gather_forward")
    return out


def rows_from_slots(slots, table_ids, row_base):
    rows = torch.empty_like(slots)
    N.check(N.launch("rows_from_slots", 1, N.lib.demb_rows_from_slots, slots.numel(), N.ptr(slots), N.ptr(_i64(table_ids)), N.ptr(row_base), N.ptr(rows), N.stream()),
            "rows_from_slots")
    return rows


def make_table_init(per_table, device) -> torch.Tensor:
    """Device image of demb_init_args_t[T] (include/dynamicemb_b200.h) from [(mode, (p0, p1, p2, p3), seed), ...]."""
    import numpy as np
    dt = np.dtype([("mode", "<i4"), ("p", "<f4", (4,)), ("reserved", "<u4"), ("seed", "<u8")])
    assert dt.itemsize == 32
    a = np.zeros(len(per_table), dtype=dt)
    for i, (mode, p, seed) in enumerate(per_table):
        a[i]["mode"], a[i]["p"], a[i]["seed"] = int(mode), tuple(float(x) for x in p), int(seed) & 0xFFFFFFFFFFFFFFFF
    return torch.from_numpy(a.view(np.uint8).copy()).to(device)


def init_rows(values, emb_dim, rows, keys, mode, p0=0.0, p1=1.0, p2=0.0, p3=0.0, seed=0, state_init=0.0, only_if=None, emb_out=None,
              table_ids=None, table_init=None):
    n = keys.numel()
    N.check(N.launch("init_rows", 1, N.lib.demb_init_rows, N.ptr(values), values.stride(0) if values is not None else emb_dim, emb_dim, n, N.ptr(rows),
                                 N.ptr(keys.contiguous()), int(mode), float(p0), float(p1), float(p2), float(p3), int(seed),
                                 N.ptr(_i64(table_ids)), N.ptr(table_init), float(state_init), N.ptr(only_if), N.ptr(emb_out), N.stream()), "init_rows")


def host_values(rows: int, width: int) -> torch.Tensor:
    """Zeroed fp32 value rows of a backing table in PINNED HOST memory (cache tier, key_value_table.py host-resident `DynamicEmbStorage`).
    Page-locked through the CUDA host allocator, so with unified addressing the row-copy kernels dereference `data_ptr()` directly over
    the host link; no CPU code touches the rows."""
    return torch.zeros(rows, width, dtype=torch.float32, pin_memory=True)


def copy_rows(values, width, rows, dense, to_table: bool):
    N.check(N.lib.demb_copy_rows(N.ptr(values), values.stride(0), width, rows.numel(), N.ptr(rows), N.ptr(dense), dense.stride(0),
                                 1 if to_table else 0, N.stream()), "copy_rows")


class BackwardPrep:
    """Side stream for `backward_prepare` (one per module).  A torch stream, so the caching allocator can be told
    (`record_stream`) that the sort workspace and `inverse` are in use there."""

    def __init__(self, device=None):
        self.stream = torch.cuda.Stream(device)


class PreparedBackward:
    """Workspace holding the pre-sorted (unique idx, gradient row) pairs + the event the gradient-dependent half waits on."""

    def __init__(self, ws: torch.Tensor, done: torch.cuda.Event):
        self.ws, self.done = ws, done


def backward_prepare(prep: BackwardPrep, emb_dim: int, inverse: torch.Tensor, num_unique_bound: int, n_dev=None, grad_row_of=None) -> Optional[PreparedBackward]:
    """Sequence mode: launch the gradient-independent half of `backward` (pair list + radix sort by unique index) NOW, on the prep
    object's stream, behind everything already enqueued on the current stream.  Returns what backward(..., prepared=) needs.
    A PreparedBackward that is dropped without a backward is safe: its tensors were recorded on the side stream, so the allocator
    does not hand the memory out again before the sort has finished."""
    n = inverse.numel()
    if n == 0:
        return None
    cur = torch.cuda.current_stream(inverse.device)
    ws = torch.empty(N.lib.demb_backward_workspace_bytes(n, emb_dim), dtype=torch.uint8, device=inverse.device)
    prep.stream.wait_stream(cur)
    with torch.cuda.stream(prep.stream):
        N.check(N.launch("backward_prepare", 7, N.lib.demb_backward_sort, emb_dim, n, N.ptr(inverse), int(num_unique_bound), None, 0, 0, -1,
                         N.ptr(n_dev), N.ptr(grad_row_of), N.ptr(ws), ws.numel(), N.stream()), "backward_sort")
        done = torch.cuda.Event()
        done.record(prep.stream)
    ws.record_stream(prep.stream)
    inverse.record_stream(prep.stream)
    return PreparedBackward(ws, done)


def backward(values, emb_dim, inverse, num_unique_bound, rows, grads, *, offsets=None, batch_size=0, num_features=0, combiner=-1,
             opt_type=0, lr=0.0, eps=1e-8, beta1=0.9, beta2=0.999, weight_decay=0.0, bc1=1.0, bc2=1.0, want_unique_grads=False, prepared=None,
             n_dev=None, grad_row_of=None, unique_grad_addr=None, grad_stride=None):
    """Fused reduce_grads + optimizer row update.  grads: [n, D] (sequence) or [B, F*D] (pooled).
    prepared = PreparedBackward from backward_prepare(same inverse / bound): only the gradient-dependent half runs here."""
    n = inverse.numel()
    dev = inverse.device
    ug = torch.zeros(num_unique_bound, emb_dim, dtype=torch.float32, device=dev) if want_unique_grads else None
    if n == 0:
        return ug
    grads = grads.contiguous()
    vstride = values.stride(0) if values is not None else emb_dim
    gstride = emb_dim if grad_stride is None else int(grad_stride)
    if prepared is not None:
        ws = prepared.ws
        torch.cuda.current_stream(dev).wait_event(prepared.done)
        N.check(N.launch("backward", 3, N.lib.demb_backward_apply, N.ptr(values), vstride, emb_dim, n, N.ptr(inverse), int(num_unique_bound),
                         N.ptr(rows), N.ptr(grads), gstride, N.ptr(_i64(offsets)), batch_size, num_features, combiner, int(opt_type), lr, eps, beta1, beta2,
                         weight_decay, bc1, bc2, N.ptr(ug), N.ptr(n_dev), N.ptr(unique_grad_addr), N.ptr(ws), ws.numel(), N.stream()), "backward")
        return ug
    ws_bytes = N.lib.demb_backward_workspace_bytes(n, emb_dim)
    ws = N.workspace(ws_bytes, dev)
    N.check(N.launch("backward", 10, N.lib.demb_backward, N.ptr(values), vstride, emb_dim, n, N.ptr(inverse),
                                int(num_unique_bound), N.ptr(rows), N.ptr(grads), gstride, N.ptr(_i64(offsets)), batch_size, num_features,
                                combiner, int(opt_type), lr, eps, beta1, beta2, weight_decay, bc1, bc2, N.ptr(ug), N.ptr(n_dev), N.ptr(grad_row_of),
                                N.ptr(unique_grad_addr), N.ptr(ws), ws.numel(), N.stream()), "backward")
    return ug


def update_rows(values, emb_dim, rows, grads, opt_type, lr, eps=1e-8, beta1=0.9, beta2=0.999, weight_decay=0.0, bc1=1.0, bc2=1.0):
    grads = grads.contiguous()
    N.check(N.lib.demb_update_rows(N.ptr(values), values.stride(0), emb_dim, rows.numel(), N.ptr(rows), N.ptr(grads), grads.stride(0),
                                   int(opt_type), lr, eps, beta1, beta2, weight_decay, bc1, bc2, N.stream()), "update_rows")


# reference-named thin aliases (dynamic_emb_op.cu:79,106,160) -----------------------------------------
def gather_embedding(unique_embs: torch.Tensor, output_embs: torch.Tensor, reverse_indices: torch.Tensor) -> None:
    n = reverse_indices.numel()
    rows = torch.arange(unique_embs.shape[0], dtype=torch.int64, device=unique_embs.device)
    N.check(N.lib.demb_gather_forward(N.ptr(unique_embs), unique_embs.stride(0), unique_embs.shape[1], n, N.ptr(rows), N.ptr(reverse_indices),
                                      None, 0, 0, -1, N.ptr(output_embs), _OUT_DTYPE[output_embs.dtype], None, N.stream()), "gather_embedding")


def gather_embedding_pooled(unique_embs, output_embs, reverse_indices, offsets, combiner, total_D, batch_size, D_offsets=None, max_D=0) -> None:
    D = unique_embs.shape[1]
    F = total_D // D
    rows = torch.arange(unique_embs.shape[0], dtype=torch.int64, device=unique_embs.device)
    N.check(N.lib.demb_gather_forward(N.ptr(unique_embs), unique_embs.stride(0), D, reverse_indices.numel(), N.ptr(rows), N.ptr(reverse_indices),
                                      N.ptr(_i64(offsets)), batch_size, F, combiner, N.ptr(output_embs), _OUT_DTYPE[output_embs.dtype], None,
                                      N.stream()), "gather_embedding_pooled")


def reduce_grads(reverse_indices, grads, num_unique, batch_size, out_dim, offsets=None, D_offsets=None, combiner=-1, total_D=0) -> torch.Tensor:
    F = (total_D // out_dim) if offsets is not None else 0
    return backward(None, out_dim, reverse_indices, num_unique, None, grads, offsets=offsets, batch_size=batch_size if offsets is not None else 0,
                    num_features=F, combiner=combiner if offsets is not None else -1, want_unique_grads=True)


# ---------------------------------------------------------------------------------------------------
# reference-named ops over per-table base pointers (module_bind.cu: dynamic_emb_op.cu:839-876, optimizer.cu:416-447, initializer.cu:193-210,
# index_calculation.cu:241, table_operation/table.cu:164) — same names, argument order and in-place conventions as the pybind module
# ---------------------------------------------------------------------------------------------------
def _flat_copy(table_ptrs, indices, table_ids, scalar_tid, dense, table_value_dims, table_emb_dims, max_emb_dim, region, to_table):
    assert dense.dtype == torch.float32 and dense.stride(1) == 1, "fp32 rows (bf16/fp16 value rows are not built, DESIGN.md)"
    N.check(N.lib.demb_flat_table_copy(N.ptr(_i64(table_ptrs)), N.ptr(_i64(table_ids)), int(scalar_tid), N.ptr(_i64(indices)), indices.numel(),
                                       N.ptr(_i64(table_value_dims)), N.ptr(_i64(table_emb_dims)), int(max_emb_dim), N.ptr(dense), dense.stride(0),
                                       dense.shape[1], region, 1 if to_table else 0, N.stream()), "flat_table_copy")


def load_from_flat_table_contiguous(table_ptrs, indices, table_id, output, table_value_dims, table_emb_dims, max_emb_dim, all_dims_vec4=True) -> None:
    _flat_copy(table_ptrs, indices, None, table_id, output, table_value_dims, table_emb_dims, max_emb_dim, 0, False)


def load_from_flat_table_emb(table_ptrs, indices, table_ids, output, table_value_dims, table_emb_dims, max_emb_dim, all_dims_vec4=True) -> None:
    _flat_copy(table_ptrs, indices, table_ids, 0, output, table_value_dims, table_emb_dims, max_emb_dim, 1, False)


def load_from_flat_table_value(table_ptrs, indices, table_ids, output, table_value_dims, table_emb_dims, max_emb_dim, all_dims_vec4=True) -> None:
    _flat_copy(table_ptrs, indices, table_ids, 0, output, table_value_dims, table_emb_dims, max_emb_dim, 2, False)


def store_to_flat_table_contiguous(table_ptrs, indices, table_id, input, table_value_dims, table_emb_dims, max_emb_dim, all_dims_vec4=True) -> None:  # noqa: A002
    _flat_copy(table_ptrs, indices, None, table_id, input, table_value_dims, table_emb_dims, max_emb_dim, 0, True)


def store_to_flat_table_value(table_ptrs, indices, table_ids, input, table_value_dims, table_emb_dims, max_emb_dim, all_dims_vec4=True) -> None:  # noqa: A002
    _flat_copy(table_ptrs, indices, table_ids, 0, input, table_value_dims, table_emb_dims, max_emb_dim, 2, True)


def _flat_update(grads, indices, table_ptrs, table_ids, table_value_dims, table_emb_dims, max_emb_dim, opt_type, lr, eps=1e-8, beta1=0.9, beta2=0.999,
                 weight_decay=0.0, bc1=1.0, bc2=1.0):
    grads = grads.contiguous()
    N.check(N.lib.demb_flat_table_update(N.ptr(_i64(table_ptrs)), N.ptr(_i64(table_ids)), N.ptr(_i64(indices)), indices.numel(), N.ptr(_i64(table_value_dims)),
                                         N.ptr(_i64(table_emb_dims)), int(max_emb_dim), N.ptr(grads), grads.stride(0), int(opt_type), float(lr), float(eps),
                                         float(beta1), float(beta2), float(weight_decay), float(bc1), float(bc2), N.stream()), "flat_table_update")


def sgd_update_for_flat_table(grads, indices, table_ptrs, table_ids, table_value_dims, table_emb_dims, max_emb_dim, all_dims_vec4, lr, table_dtype=0) -> None:
    _flat_update(grads, indices, table_ptrs, table_ids, table_value_dims, table_emb_dims, max_emb_dim, OptimizerType.SGD, lr)


def adam_update_for_flat_table(grads, indices, table_ptrs, table_ids, table_value_dims, table_emb_dims, lr, beta1, beta2, eps, weight_decay, iter_num,
                               max_emb_dim, all_dims_vec4, table_dtype=0) -> None:
    it = max(int(iter_num), 1)
    _flat_update(grads, indices, table_ptrs, table_ids, table_value_dims, table_emb_dims, max_emb_dim, OptimizerType.ADAM, lr, eps, beta1, beta2, weight_decay,
                 1.0 - beta1 ** it, 1.0 - beta2 ** it)


def adagrad_update_for_flat_table(grads, indices, table_ptrs, table_ids, table_value_dims, table_emb_dims, lr, eps, max_emb_dim, all_dims_vec4, table_dtype=0) -> None:
    _flat_update(grads, indices, table_ptrs, table_ids, table_value_dims, table_emb_dims, max_emb_dim, OptimizerType.ADAGRAD, lr, eps)


def rowwise_adagrad_for_flat_table(grads, indices, table_ptrs, table_ids, table_value_dims, table_emb_dims, lr, eps, max_emb_dim, all_dims_vec4, table_dtype=0) -> None:
    _flat_update(grads, indices, table_ptrs, table_ids, table_value_dims, table_emb_dims, max_emb_dim, OptimizerType.ROWWISE_ADAGRAD, lr, eps)


def flagged_compact(flags: torch.Tensor, inputs: List[Optional[torch.Tensor]]):
    """index_calculation.cu:130.  Returns (count, indices[count], [compacted inputs or None]) like the reference (which reads the count
    back: ONE host sync); `flagged_compact_device` keeps the count on the device."""
    cnt, idx, outs = flagged_compact_device(flags, inputs)
    c = int(cnt.item())
    return c, idx[:c], [o[:c] if o is not None else None for o in outs]


def flagged_compact_device(flags: torch.Tensor, inputs: List[Optional[torch.Tensor]]):
    """(count[1] device, indices[n], outputs[n]...) — only the first count entries are meaningful; one launch, no host sync."""
    assert flags.dtype == torch.bool, "flags must be bool"
    n, dev = flags.numel(), flags.device
    real = [t for t in inputs if t is not None]
    assert len(real) <= 4 and all(t.dim() == 1 and t.element_size() == 8 and t.numel() == n for t in real), "up to 4 1-D 8-byte inputs of the flags' length"
    cnt = torch.zeros(1, dtype=torch.int64, device=dev)
    idx = torch.empty(n, dtype=torch.int64, device=dev)
    outs_real = [torch.empty_like(t) for t in real]
    ins = (ctypes.c_void_p * 4)(*[t.data_ptr() for t in real] + [None] * (4 - len(real)))
    outs = (ctypes.c_void_p * 4)(*[t.data_ptr() for t in outs_real] + [None] * (4 - len(real)))
    ws = N.workspace(N.lib.demb_flagged_compact_workspace_bytes(n), dev)
    N.check(N.launch("flagged_compact", 2, N.lib.demb_flagged_compact, n, N.ptr(flags), N.ptr(cnt), N.ptr(idx), ctypes.cast(ins, ctypes.c_void_p),
                     ctypes.cast(outs, ctypes.c_void_p), len(real), N.ptr(ws), ws.numel(), N.stream()), "flagged_compact")
    it = iter(outs_real)
    return cnt, idx, [next(it) if t is not None else None for t in inputs]


def bucketize_keys(keys: torch.Tensor, table_ids: torch.Tensor, table_bucket_offsets: torch.Tensor, num_buckets: int, bucket_capacity: int):
    """table_operation/bucketize.cu:111: keys ordered by (global bucket, key), offsets[active buckets + 1] into that order, inverse (sorted
    position -> input position).  Only the reference's DEMB_DETERMINISM_MODE helper calls it; our insert is deterministic by construction."""
    n, dev = keys.numel(), keys.device
    if n == 0:
        z = torch.empty(0, dtype=torch.int64, device=dev)
        return z, z.clone(), z.clone()
    bkt = torch.empty(n, dtype=torch.int64, device=dev)
    N.check(N.lib.demb_bucket_of(N.ptr(table_bucket_offsets), int(bucket_capacity), n, N.ptr(keys.contiguous()), N.ptr(_i64(table_ids)), N.ptr(bkt), N.stream()),
            "bucket_of")
    k_order = torch.sort(keys.view(torch.int64) if keys.dtype == torch.uint64 else keys, stable=True).indices     # signed order for int64 keys
    b_sorted, order2 = torch.sort(bkt[k_order], stable=True)
    inverse = k_order[order2]
    ends = torch.nonzero(torch.cat([b_sorted[1:] != b_sorted[:-1], torch.ones(1, dtype=torch.bool, device=dev)])).flatten() + 1
    offsets = torch.cat([torch.zeros(1, dtype=torch.int64, device=dev), ends])
    return keys[inverse], offsets, inverse


class CurandStateContext:
    """initializer.cu:26-112 keeps a pool of curand states; our initializers are counter-based (Philox keyed by seed, key, column), so the
    context only carries the seed."""

    def __init__(self, seed: int = 0):
        self.seed = seed

    def ptr(self):
        return self


def _init_op(buffer, indices, mode, p, seed=0, keys=None):
    """buffer[indices[i], :] = initializer(...) for indices >= 0 (initializer.cu:114-190).  The random stream is keyed by `keys` when given,
    else by the row index."""
    assert buffer.dim() == 2 and buffer.is_contiguous() and buffer.dtype == torch.float32, "dense fp32 [n, dim] buffer"
    idx = _i64(indices)
    k = keys if keys is not None else idx
    N.check(N.launch("init_rows", 1, N.lib.demb_init_rows, N.ptr(buffer), buffer.stride(0), buffer.shape[1], idx.numel(), N.ptr(idx), N.ptr(k.contiguous()),
                     int(mode), float(p[0]), float(p[1]), float(p[2]), float(p[3]), int(seed), None, None, 0.0, None, None, N.stream()), "init_op")


def normal_init(buffer, indices, curand_state_context, mean, std_dev) -> None:
    _init_op(buffer, indices, InitializerMode.NORMAL, (mean, std_dev, 0.0, 0.0), getattr(curand_state_context, "seed", 0))


def truncated_normal_init(buffer, indices, curand_state_context, mean, std_dev, lower, upper) -> None:
    _init_op(buffer, indices, InitializerMode.TRUNCATED_NORMAL, (mean, std_dev, lower, upper), getattr(curand_state_context, "seed", 0))


def uniform_init(buffer, indices, curand_state_context, lower, upper) -> None:
    _init_op(buffer, indices, InitializerMode.UNIFORM, (lower, upper, 0.0, 0.0), getattr(curand_state_context, "seed", 0))


def const_init(buffer, indices, value) -> None:
    _init_op(buffer, indices, InitializerMode.CONSTANT, (value, 0.0, 0.0, 0.0))


def debug_init(buffer, indices, keys) -> None:
    _init_op(buffer, indices, InitializerMode.DEBUG, (0.0, 0.0, 0.0, 0.0), keys=keys)


# ---------------------------------------------------------------------------------------------------
# input dist
# ---------------------------------------------------------------------------------------------------
DIST_TYPE = {"continuous": 0, "roundrobin": 1, "hash_roundrobin": 2}


def block_bucketize_sparse_features(lengths: torch.Tensor, indices: torch.Tensor, batch_size: int, world_size: int, block_sizes: torch.Tensor,
                                    dist_type_per_feature: Optional[torch.Tensor] = None, weights: Optional[torch.Tensor] = None,
                                    sequence: bool = True):
    """sparse_block_bucketize_features.cu:372.  lengths[F*B] feature-major.  Returns (new_lengths[W*F*B], new_indices[n],
    new_weights or None, unbucketize_permute[n] or None)."""
    dev = indices.device
    S = lengths.numel()
    offsets = torch.zeros(S + 1, dtype=torch.int64, device=dev)
    torch.cumsum(lengths.to(torch.int64), 0, out=offsets[1:])
    n = indices.numel()
    new_lengths = torch.zeros(world_size * S, dtype=torch.int64, device=dev)
    new_ids = torch.empty(n, dtype=torch.int64, device=dev)
    perm = torch.empty(n, dtype=torch.int64, device=dev) if sequence else None
    new_w = torch.empty(n, dtype=torch.float32, device=dev) if weights is not None else None
    ws = N.workspace(N.lib.demb_bucketize_workspace_bytes(S, world_size), dev)
    dt = dist_type_per_feature.to(torch.int32).contiguous() if dist_type_per_feature is not None else None
    N.check(N.launch("bucketize", 2, N.lib.demb_block_bucketize_sparse_features_n, S, batch_size, world_size, n, N.ptr(offsets), N.ptr(_i64(indices)),
                                                       N.ptr(_i64(block_sizes)), N.ptr(dt), N.ptr(weights), N.ptr(new_lengths), N.ptr(new_ids),
                                                       N.ptr(perm), N.ptr(new_w), N.ptr(ws), ws.numel(), N.stream()),
            "block_bucketize_sparse_features")
    return new_lengths, new_ids, new_w, perm
