"""Frequency-based admission: a key enters the embedding table only after it has been seen often enough.

Host mirror of /root/reference/corelib/dynamicemb/dynamicemb/embedding_admission.py:35-180 (KVCounter, MultiTableKVCounter,
FrequencyAdmissionStrategy) and of the admission branch of `_prefetch_hbm_direct_path`
(batched_dynamicemb_function.py:612-689) for the HBM-direct storage tier — SURVEY.md §8(f) row 4, admission part.  The counter is one
more `LinearBucketTable` (ACCUMULATE score = the frequency seen so far), so every device step is one of the table kernels of
csrc/demb_table.cu: there is no admission-specific kernel and nothing here runs on the CPU.

`admission_split` is the host-side decision sequence, written against the table / counter INTERFACES (lookup-found flags in, masks out)
so the CPU tests drive it with dictionary-backed stand-ins and the module drives it with the GPU tables.
"""
from typing import List, Optional, Tuple

import torch

from .dynamicemb_extensions import ScorePolicy
from .scored_hashtable import ScoreArg, ScoreSpec, get_scored_table
from .types import AdmissionStrategy, Counter, DynamicEmbInitializerArgs, MemoryType


class KVCounter:
    """Per-table counter configuration (embedding_admission.py:35-51).  One per logical table; the module wraps the list into ONE
    `MultiTableKVCounter` over a fused multi-table scored hash table."""

    def __init__(self, capacity: int, bucket_capacity: int = 1024, key_type: torch.dtype = torch.int64):
        self.capacity = capacity
        self.bucket_capacity = bucket_capacity
        self.key_type = key_type


class MultiTableKVCounter(Counter):
    """key -> accumulated frequency, per logical table (embedding_admission.py:54-104).

    `add` = insert with ScorePolicy.ACCUMULATE (a key that is new gets `frequencies[i]`, a resident key `old + frequencies[i]`; a full
    bucket evicts its least-frequent key, which is the reference's behaviour too) followed by a read-only lookup of the stored values.
    The reference reads them from the insert's `score_out`; the stored value is the same number, and the read-only lookup is the path
    the GPU parity tests cover for every policy."""

    def __init__(self, kv_counters: List[KVCounter], device: torch.device):
        if not kv_counters:
            raise ValueError("kv_counters must be non-empty")
        self.score_name_ = "counter"
        self.score_specs_ = [ScoreSpec(name=self.score_name_, policy=ScorePolicy.ACCUMULATE)]
        self.table_ = get_scored_table([kv.capacity for kv in kv_counters], kv_counters[0].bucket_capacity, kv_counters[0].key_type,
                                       self.score_specs_, device)          # positional order of the reference's call (:72-78)

    def add(self, keys: torch.Tensor, table_ids: torch.Tensor, frequencies: torch.Tensor) -> torch.Tensor:
        n = keys.numel()
        if n == 0:
            return torch.empty(0, dtype=torch.int64, device=keys.device)
        freq = frequencies.to(torch.int64).contiguous()
        self.table_.insert(keys, table_ids, ScoreArg(name=self.score_name_, value=freq, policy=ScorePolicy.ACCUMULATE))
        stored, founds, _ = self.table_.lookup(keys, table_ids, ScoreArg(name=self.score_name_, policy=ScorePolicy.CONST))
        # a key the counter could not take (every slot of its bucket pinned / illegal key) counts as seen `frequencies[i]` times
        return torch.where(founds, stored, freq)

    def erase(self, keys: torch.Tensor, table_ids: torch.Tensor) -> None:
        if keys.numel():
            self.table_.erase(keys, table_ids)

    def memory_usage(self, mem_type=MemoryType.DEVICE) -> int:
        return self.table_.memory_usage() if mem_type == MemoryType.DEVICE else 0

    def load(self, key_file, counter_file, table_id: int) -> None:
        self.table_.load(key_file, {self.score_name_: counter_file}, table_id=table_id)

    def dump(self, key_file, counter_file, table_id: int) -> None:
        self.table_.dump(key_file, {self.score_name_: counter_file}, table_id=table_id)


class FrequencyAdmissionStrategy(AdmissionStrategy):
    """Admit keys whose accumulated frequency is >= `threshold` (embedding_admission.py:107-180).  `initializer_args` is what the
    reference's generic-storage path uses for rows of keys that are not admitted; on the HBM-direct path (the one built here, like the
    reference's :1090-1097) such keys read the table's own initializer and are not stored."""

    def __init__(self, threshold: int, initializer_args: Optional[DynamicEmbInitializerArgs] = None):
        if threshold < 0:
            raise ValueError(f"Threshold must be non-negative, got {threshold}")
        self.threshold = threshold
        self.initializer_args = initializer_args

    def admit(self, keys: torch.Tensor, frequencies: torch.Tensor) -> torch.Tensor:
        if keys.shape[0] != frequencies.shape[0]:
            raise ValueError(f"Keys and frequencies must have same length, got {keys.shape[0]} and {frequencies.shape[0]}")
        return frequencies >= self.threshold

    def initialize_non_admitted_embeddings(self, buffer: torch.Tensor, indices: torch.Tensor) -> bool:
        """Fill `buffer[indices]` from `initializer_args`; False when there are none (the caller then uses the table's initializer).
        Only the modes that do not need the key are possible here (the reference passes keys=None too): CONSTANT, or UNIFORM /
        NORMAL drawn by torch's generator on the buffer's device."""
        a = self.initializer_args
        if a is None:
            return False
        from .types import DynamicEmbInitializerMode as M
        n, d = indices.numel(), buffer.size(1)
        if n == 0:
            return True
        if a.mode == M.CONSTANT:
            rows = torch.full((n, d), float(a.value), dtype=buffer.dtype, device=buffer.device)
        elif a.mode == M.UNIFORM:
            lo = -1.0 if a.lower is None else a.lower
            up = 1.0 if a.upper is None else a.upper
            rows = torch.empty(n, d, dtype=buffer.dtype, device=buffer.device).uniform_(lo, up)
        elif a.mode == M.NORMAL:
            rows = torch.empty(n, d, dtype=buffer.dtype, device=buffer.device).normal_(a.mean, a.std_dev)
        else:
            raise NotImplementedError(f"non-admitted initializer mode {a.mode} needs the keys; use the table's initializer (initializer_args=None)")
        buffer[indices] = rows
        return True


def admission_split(missing_keys: torch.Tensor, missing_table_ids: torch.Tensor, missing_freq: Optional[torch.Tensor],
                    admit_strategy: AdmissionStrategy, admission_counter: Counter) -> Tuple[torch.Tensor, torch.Tensor]:
    """The admission decision of one prefetch (batched_dynamicemb_function.py:615-641): count the missing keys, ask the strategy,
    forget the admitted ones.  Returns (admit_mask [n] bool, accumulated frequency [n] int64).

    `missing_freq` = how often each missing key occurs in this batch (LFU-style scores) or None = once per batch, as the reference
    counts when the table's score is not a frequency."""
    n = missing_keys.numel()
    counters = (missing_freq.to(torch.int64) if missing_freq is not None
                else torch.ones(n, dtype=torch.int64, device=missing_keys.device))
    freq = admission_counter.add(missing_keys, missing_table_ids, counters)
    admit_mask = admit_strategy.admit(missing_keys, freq).to(torch.bool)
    if n and bool(admit_mask.any()):
        admission_counter.erase(missing_keys[admit_mask], missing_table_ids[admit_mask])
    return admit_mask, freq
