"""LinearBucketTable: the scored GPU hash table (key -> slot index) behind DynamicEmb.

Host-side mirror of /root/reference/corelib/dynamicemb/dynamicemb/scored_hashtable.py:294-1757
(class LinearBucketTable) on top of the B200 kernels: same constructor, same storage image
(`table_storage_`, `keys_`, `digests_`, `scores_list`), same lookup/insert/insert_and_evict/erase
results.  Differences: insert is always deterministic (the reference needs
DEMB_DETERMINISM_MODE and one launch per wave), no overflow bucket, single ScoreSpec.
"""
import enum
import warnings
from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple

import torch

from . import dynamicemb_extensions as ext
from .dynamicemb_extensions import ScorePolicy


@dataclass(frozen=True)
class ScoreSpec:
    name: str
    policy: ScorePolicy
    dtype: torch.dtype = torch.uint64
    priority: int = 0
    is_reduction: bool = True


@dataclass
class ScoreArg:
    name: str
    value: Optional[torch.Tensor] = None
    policy: Optional[ScorePolicy] = None


def score_policy_num_scores(policy) -> int:
    return 2 if policy == ScorePolicy.LRU_LFU else 1


def uint64_to_int64(x: int) -> int:
    return x if x < (1 << 63) else x - (1 << 64)


def murmur3_hash_64bits(key: int) -> int:
    """fmix64 (scored_hashtable.py:279-291); the table uses this & INT64_MAX (types.cuh:123-131)."""
    k = key & 0xFFFFFFFFFFFFFFFF
    k ^= k >> 33
    k = (k * 0xFF51AFD7ED558CCD) & 0xFFFFFFFFFFFFFFFF
    k ^= k >> 33
    k = (k * 0xC4CEB9FE1A85EC53) & 0xFFFFFFFFFFFFFFFF
    k ^= k >> 33
    return k


class LinearBucketTable:
    def __init__(self, capacity: List[int], score_specs: List[ScoreSpec], key_type: torch.dtype = torch.int64,
                 bucket_capacity: Optional[int] = None, device: torch.device = None, enable_overflow: bool = False):
        if enable_overflow:
            raise NotImplementedError("overflow bucket (cache tier) is out of scope; see DESIGN.md")
        self.device = device if device is not None else torch.device("cuda", torch.cuda.current_device())
        assert key_type in (torch.int64, torch.uint64), "Only accept 64 bits integer as key's type."
        assert len(score_specs) == 1, "Only a single ScoreSpec is supported."
        self.key_type_ = key_type
        self.score_specs_ = list(score_specs)
        self.score_names_ = [s.name for s in score_specs]
        self.num_scores_ = score_policy_num_scores(score_specs[0].policy)
        if bucket_capacity is None:
            bucket_capacity = 128
        self.bucket_capacity_ = ((bucket_capacity + 15) // 16) * 16          # digest uint4 load granularity (:362-376)
        if self.bucket_capacity_ != bucket_capacity:
            warnings.warn(f"Bucket capacity is rounded from {bucket_capacity} to {self.bucket_capacity_}.", UserWarning)
        assert isinstance(capacity, list) and len(capacity) >= 1
        self.num_tables_ = len(capacity)
        C = self.bucket_capacity_
        self.per_table_num_buckets_ = [(c + C - 1) // C for c in capacity]
        self.per_table_capacity_ = [nb * C for nb in self.per_table_num_buckets_]
        off = [0]
        for nb in self.per_table_num_buckets_:
            off.append(off[-1] + nb)
        self.num_buckets_ = off[-1]
        self.capacity_ = self.num_buckets_ * C
        self.table_bucket_offsets_cpu_ = torch.tensor(off, dtype=torch.int64)
        self.table_bucket_offsets_ = self.table_bucket_offsets_cpu_.to(self.device)
        self.row_base_ = (self.table_bucket_offsets_[:-1] * C).contiguous()     # first global slot of each table
        if self.capacity_ != sum(capacity):
            warnings.warn(f"Table total capacity is rounded from {sum(capacity)} to {self.capacity_}.", UserWarning)
        self.storage_bytes_ = (9 + 8 * self.num_scores_) * C * self.num_buckets_
        self.table_storage_ = torch.empty(self.storage_bytes_, dtype=torch.uint8, device=self.device)
        fields = [self.key_type_, torch.uint8] + [torch.uint64] * self.num_scores_
        # scores are AoS per key; expose them as one [num_buckets, C*ns] uint64 view like the reference's per-word list
        self.keys_, self.digests_ = ext.table_partition(self.table_storage_, fields, C, self.num_buckets_)[:2]
        self.bucket_sizes = torch.zeros(self.num_buckets_, dtype=torch.int32, device=self.device)
        self._ref_counter = torch.zeros(self.capacity_, dtype=torch.int32, device=self.device)
        self._bucket_heads = torch.empty(self.num_buckets_, dtype=torch.int32, device=self.device)   # per-bucket pending-insert lists (fused prefetch)
        self.reset()

    # ------------------------------------------------------------------ properties
    @property
    def key_type(self) -> torch.dtype:
        return self.key_type_

    @property
    def index_type(self) -> torch.dtype:
        return torch.int64

    @property
    def result_type(self) -> torch.dtype:
        return torch.uint8

    @property
    def score_specs(self) -> List[ScoreSpec]:
        return self.score_specs_

    def reset(self) -> None:
        ext.table_init(self.table_storage_, self.bucket_capacity_, self.num_scores_)
        self.bucket_sizes.zero_()
        self._ref_counter.zero_()
        ext.fill_i32(self._bucket_heads, -1)

    def capacity(self, table_id: Optional[int] = None) -> int:
        return self.capacity_ if table_id is None else self.per_table_capacity_[table_id]

    def size(self, table_id: Optional[int] = None) -> int:
        if table_id is None:
            return int(self.bucket_sizes.sum().item())
        b, e = int(self.table_bucket_offsets_cpu_[table_id]), int(self.table_bucket_offsets_cpu_[table_id + 1])
        return int(self.bucket_sizes[b:e].sum().item())

    def load_factor(self) -> float:
        return self.size() / max(self.capacity_, 1)

    def memory_usage(self) -> int:
        return self.table_storage_.numel() + self.bucket_sizes.numel() * 4 + self._ref_counter.numel() * 4

    def _parse_score(self, score: ScoreArg) -> Tuple[Optional[torch.Tensor], ScorePolicy]:
        index = self.score_names_.index(score.name)
        policy = score.policy if score.policy is not None else self.score_specs_[index].policy
        return score.value, policy

    # ------------------------------------------------------------------ ops
    def lookup(self, keys, table_ids, score: ScoreArg, timestamp: int = 0):
        """(score_out, founds, indices) — scored_hashtable.py:537."""
        value, policy = self._parse_score(score)
        return ext.table_lookup(self.table_storage_, self.table_bucket_offsets_, self.bucket_capacity_, keys, table_ids, value, policy,
                                num_scores=self.num_scores_, timestamp=timestamp)

    def insert(self, keys, table_ids, score: ScoreArg, insert_results=None, score_out=None, timestamp: int = 0) -> torch.Tensor:
        """Keys have to be unique.  Returns indices — scored_hashtable.py:565."""
        value, policy = self._parse_score(score)
        return ext.table_insert(self.table_storage_, self.table_bucket_offsets_, self.bucket_capacity_, self.bucket_sizes, keys, table_ids,
                                value, policy, self._ref_counter, insert_results, score_out, num_scores=self.num_scores_, timestamp=timestamp)

    def insert_and_evict(self, keys, table_ids, score: ScoreArg, insert_results=None, score_out=None, timestamp: int = 0):
        """(indices, num_evicted, evicted_keys, evicted_indices, evicted_scores, evicted_table_ids) — :606 (one host sync, as :658)."""
        value, policy = self._parse_score(score)
        idx, nev, ek, ei, es, et = ext.table_insert_and_evict(
            self.table_storage_, self.table_bucket_offsets_, self.bucket_capacity_, self.bucket_sizes, keys, table_ids, value, policy,
            self._ref_counter, insert_results, score_out, num_scores=self.num_scores_, timestamp=timestamp)
        h = int(nev.cpu().item())
        return idx, h, ek[:h], ei[:h], es[:h], et[:h]

    def erase(self, keys, table_ids) -> None:
        ext.table_erase(self.table_storage_, self.table_bucket_offsets_, self.bucket_capacity_, self.bucket_sizes, keys, table_ids,
                        num_scores=self.num_scores_)

    def increment_counter(self, slot_indices, table_ids) -> None:
        ext.table_update_counter_with_layout(self._ref_counter, slot_indices, 1, self.table_bucket_offsets_, self.bucket_capacity_,
                                             table_ids=table_ids)

    def decrement_counter(self, slot_indices, table_ids) -> None:
        ext.table_update_counter_with_layout(self._ref_counter, slot_indices, -1, self.table_bucket_offsets_, self.bucket_capacity_,
                                             table_ids=table_ids)

    # ------------------------------------------------------------------ whole score blocks (checkpoint of multi-word scores)
    def _score_words(self) -> torch.Tensor:
        """[num_buckets, C, num_scores] int64 strided view of the score words (AoS per key behind keys and digests of a bucket)."""
        C, ns = self.bucket_capacity_, self.num_scores_
        bucket_words = (9 + 8 * ns) * C // 8
        flat = self.table_storage_.view(torch.uint8).reshape(-1)[9 * C:].view(torch.int64)
        return torch.as_strided(flat, (self.num_buckets_, C, ns), (bucket_words, ns, 1))

    def _global_slots(self, table_id: int, indices: torch.Tensor):
        g = indices.to(torch.int64) + int(self.table_bucket_offsets_cpu_[table_id]) * self.bucket_capacity_
        return torch.div(g, self.bucket_capacity_, rounding_mode="floor"), g % self.bucket_capacity_

    def gather_score_blocks(self, table_id: int, indices: torch.Tensor) -> torch.Tensor:
        """All score words of the given table-local slots, [n, num_scores] int64 in device (physical) order
        (reference scored_hashtable.py gather_score_blocks, used by export_keys_values_iter key_value_table.py:1111-1113)."""
        b, s = self._global_slots(table_id, indices)
        return self._score_words()[b, s].contiguous()

    def scatter_score_blocks(self, table_id: int, indices: torch.Tensor, blocks: torch.Tensor) -> None:
        """Write [n, num_scores] score blocks at table-local slots; negative slots (failed inserts) are skipped
        (reference scatter_score_blocks, key_value_table.py:1471-1474)."""
        if blocks.dim() != 2 or blocks.size(1) != self.num_scores_ or blocks.size(0) != indices.numel():
            raise ValueError(f"expects [{indices.numel()}, {self.num_scores_}] scores, got {tuple(blocks.shape)}")
        ok = indices >= 0
        b, s = self._global_slots(table_id, indices[ok])
        self._score_words()[b, s] = blocks[ok].to(torch.int64)

    def export(self, table_id: int = 0, batch: int = 65536, threshold: Optional[int] = None):
        """Scan one logical table; yields (keys, scores, table-local slot indices) batches (BATCH_SIZE_PER_DUMP scan, :export)."""
        C = self.bucket_capacity_
        b, e = int(self.table_bucket_offsets_cpu_[table_id]) * C, int(self.table_bucket_offsets_cpu_[table_id + 1]) * C
        for off in range(b, e, batch):
            n = min(batch, e - off)
            counter, keys, scores, idx = ext.table_export_batch(self.table_storage_, C, n, off, self.key_type_, threshold, table_begin=b,
                                                                num_scores=self.num_scores_, score_index=self.num_scores_ - 1)
            c = int(counter.item())
            if c:
                yield keys[:c], scores[:c], idx[:c]

    def incremental_dump(self, score_threshold: Dict[str, int], table_id: int, batch_size: int = 65536, pg=None, return_index: bool = False):
        """Keys (and scores, optionally slot indices) of one logical table whose score is not below the threshold, as CPU tensors; with a
        process group of more than one rank every rank receives all ranks' entries in rank order (scored_hashtable.py:1128-1260).
        The scan is the export kernel, the filter a torch mask."""
        import torch.distributed as dist
        name, threshold = None, 0
        for score_name, thr in score_threshold.items():
            if score_name not in self.score_names_:
                print(f"Score name {score_name} not existed, will not dump it.")
            else:
                name, threshold = score_name, int(thr)
        ks, ss, ix = [], [], []
        for keys, scores, idx in self.export(table_id, batch=batch_size):
            keep = scores >= threshold
            ks.append(keys[keep]); ss.append(scores[keep]); ix.append(idx[keep])
        cat = lambda xs, dt: torch.cat(xs) if xs else torch.empty(0, dtype=dt, device=self.device)      # noqa: E731
        keys, scores, idx = cat(ks, self.key_type_), cat(ss, torch.int64), cat(ix, torch.int64)
        if pg is not None and dist.is_initialized() and dist.get_world_size(group=pg) > 1:
            from .checkpoint import all_gather_keys_values
            both = torch.stack([scores, idx], dim=1)
            keys, both = all_gather_keys_values(keys, both, pg)
            scores, idx = both[:, 0].contiguous(), both[:, 1].contiguous()
        else:
            keys, scores, idx = keys.cpu(), scores.cpu(), idx.cpu()
        out_scores = {name: scores} if name is not None else {}
        return (keys, out_scores, idx) if return_index else (keys, out_scores)

    # ------------------------------------------------------------------ file round trip of one logical table (admission counters)
    def dump(self, key_file: str, score_files: Dict[str, str], table_id: Optional[int] = None) -> None:
        """ScoredHashTable.dump (scored_hashtable.py:1083-1126): raw little-endian int64 keys and uint64 scores of the live slots, one
        logical table (or all of them, table after table); GLOBAL_TIMER scores are written as ages."""
        name = self.score_names_[0]
        fscore = None
        for score_name, path in score_files.items():
            if score_name != name:
                print(f"Score name {score_name} not existed, will not dump to {path}.")
            else:
                fscore = open(path, "wb")
        timed = self.score_specs_[0].policy == ScorePolicy.GLOBAL_TIMER
        ts = ext.device_timestamp() if timed else 0
        with open(key_file, "wb") as fkey:
            for tid in ([table_id] if table_id is not None else range(self.num_tables_)):
                for keys, scores, _ in self.export(tid):
                    fkey.write(keys.cpu().numpy().tobytes())
                    if fscore is not None:
                        s = scores.view(torch.int64)
                        fscore.write(((ts - s) if timed else s).cpu().numpy().tobytes())
        if fscore is not None:
            fscore.close()

    def load(self, key_file: str, score_files: Dict[str, str], table_id: Optional[int] = None, batch_size: int = 65536) -> None:
        """ScoredHashTable.load (scored_hashtable.py:848-950): insert the keys of the file into logical table `table_id` (default 0) with
        their stored scores (ASSIGN); with more than one rank each rank keeps `key % world_size == rank` like the reference."""
        import os
        import numpy as np
        import torch.distributed as dist
        name = self.score_names_[0]
        tid = table_id if table_id is not None else 0
        score_path = score_files.get(name)
        if score_path is None or not os.path.exists(score_path):
            print(f"Will not load scores for {name}, as not provide the file path or file path not existed.")
            score_path = None
        num_keys = os.path.getsize(key_file) // 8
        if score_path is not None and os.path.getsize(score_path) // 8 != num_keys:
            raise ValueError(f"The number of keys({num_keys}) in {key_file} does not match with number of scores"
                             f"({os.path.getsize(score_path) // 8}) in {score_path}.")
        world, rank = (dist.get_world_size(), dist.get_rank()) if dist.is_initialized() else (1, 0)
        timed = self.score_specs_[0].policy == ScorePolicy.GLOBAL_TIMER
        ts = ext.device_timestamp()
        fscore = open(score_path, "rb") if score_path is not None else None
        with open(key_file, "rb") as fkey:
            for start in range(0, num_keys, batch_size):
                m = min(batch_size, num_keys - start)
                keys = np.frombuffer(fkey.read(8 * m), dtype=np.int64)
                scores = np.frombuffer(fscore.read(8 * m), dtype=np.int64) if fscore is not None else None
                if world > 1:
                    keep = (keys.view(np.uint64) % np.uint64(world)) == np.uint64(rank)
                    keys, scores = keys[keep], (scores[keep] if scores is not None else None)
                if keys.size == 0:
                    continue
                k = torch.from_numpy(keys.copy()).to(self.device)
                t = torch.full((k.numel(),), tid, dtype=torch.int64, device=self.device)
                if scores is None:      # no score file: timed tables stamp the loading time, the others start from 1
                    v = None if timed else torch.ones(k.numel(), dtype=torch.int64, device=self.device)
                    self.insert(k, t, ScoreArg(name=name, value=v), timestamp=ts)
                else:
                    s = torch.from_numpy(scores.copy()).to(self.device)
                    if timed:
                        s = torch.clamp(ts - s, min=0)
                    self.insert(k, t, ScoreArg(name=name, value=s, policy=ScorePolicy.ASSIGN), timestamp=ts)
        if fscore is not None:
            fscore.close()


class ProbingType(enum.Enum):
    """scored_hashtable.py:72-75"""
    LINEAR = "linear"
    CHAINED = "separate_chain"


@enum.unique
class ReductionType(enum.Enum):
    """scored_hashtable.py:77-80"""
    LINEAR = "linear"
    DOUBLY_LINKED = "doubly_linked"


def get_scored_table(capacity: List[int], bucket_capacity: Optional[int] = None, key_type: Optional[torch.dtype] = torch.int64,
                     score_specs: Optional[List[ScoreSpec]] = None, device: torch.device = None, probing_type=ProbingType.LINEAR,
                     reduction_type=ReductionType.LINEAR, bucket_load_factor=0.5, enable_overflow: bool = False) -> LinearBucketTable:
    """scored_hashtable.py:1734-1757 (linear probing + linear reduction is the only table the reference builds either)."""
    if probing_type != ProbingType.LINEAR or reduction_type != ReductionType.LINEAR:
        raise NotImplementedError
    if score_specs is None:
        score_specs = [ScoreSpec(name="timestamp", policy=ScorePolicy.GLOBAL_TIMER)]
    return LinearBucketTable(capacity, score_specs, key_type=key_type, bucket_capacity=bucket_capacity, device=device,
                             enable_overflow=enable_overflow)
