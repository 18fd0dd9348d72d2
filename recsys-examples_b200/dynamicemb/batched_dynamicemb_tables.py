"""BatchedDynamicEmbeddingTablesV2 — the TBE-shaped module TorchRec's row-wise sharding wrapper calls.

Drop-in for /root/reference/corelib/dynamicemb/dynamicemb/batched_dynamicemb_tables.py:452-1440
(ctor :462-508, forward :999-1088, prefetch :1090) for the HBM-direct storage tier
(`DynamicEmbStorage`, key_value_table.py:1654) and, with `caching=True`, an HBM cache in front of a table whose value rows live in pinned
host memory (`DynamicEmbCache`, `_prefetch_cache_path` batched_dynamicemb_function.py:296-556; every row is still moved by the GPU
kernels — "no CPU fallback"; external parameter servers and the cache-less hybrid tier are out of scope).  Orchestration follows batched_dynamicemb_function.py
(dynamicemb_prefetch :699, _prefetch_hbm_direct_path :559, DynamicEmbeddingFunction :1044/:1194)
with the kernel sequence collapsed:

  reference (15-20 launches, 3 host syncs)          here
  ------------------------------------------------  -----------------------------------------------
  get_table_range, segmented_unique(3), expand ids  get_table_range, segmented_unique (ids come out)
  table_lookup, flagged_compact                      table_lookup (+1 host sync for #missing)
  initializer, table_insert(+unlock), store_to_flat  table_insert (deterministic), init_rows (fused)
  increment_counter x2                               counter_update
  load_from_flat, gather_embedding[_pooled]          gather_forward (one pass, no staging copy)
  reduce_grads (sort+2), optimizer, decrement        backward (sort + fused reduce/update), counter
"""
import os
import warnings
from collections import deque
from dataclasses import dataclass
from itertools import accumulate
from typing import Deque, Dict, List, Optional

import torch
from torch import nn

from . import dynamicemb_extensions as ext
from .dynamicemb_extensions import InitializerMode, ScorePolicy
from .optimizer import OptimizerArgs, SparseOptimizer
from .scored_hashtable import LinearBucketTable, ScoreArg, ScoreSpec
from .types import (BoundsCheckMode, DynamicEmbCheckMode, DynamicEmbEvictStrategy, DynamicEmbInitializerArgs, DynamicEmbInitializerMode,
                    DynamicEmbPoolingMode, DynamicEmbScoreStrategy, DynamicEmbTableOptions, EmbOptimType)

_INIT_MODE = {
    DynamicEmbInitializerMode.NORMAL: InitializerMode.NORMAL,
    DynamicEmbInitializerMode.TRUNCATED_NORMAL: InitializerMode.TRUNCATED_NORMAL,
    DynamicEmbInitializerMode.UNIFORM: InitializerMode.UNIFORM,
    DynamicEmbInitializerMode.CONSTANT: InitializerMode.CONSTANT,
    DynamicEmbInitializerMode.DEBUG: InitializerMode.DEBUG,
}


def _init_params(a: DynamicEmbInitializerArgs, num_embeddings_hint: Optional[int] = None):
    """(mode, p0..p3) for demb_init_rows; UNIFORM/TRUNCATED_NORMAL bounds default to +-1/sqrt(N) (dynamicemb_config.py:619-658)."""
    mode = _INIT_MODE[a.mode]
    lo, up = a.lower, a.upper
    if lo is None or up is None:
        bound = 1.0 / (float(num_embeddings_hint) ** 0.5) if num_embeddings_hint else 1.0
        lo = -bound if lo is None else lo
        up = bound if up is None else up
    if a.mode == DynamicEmbInitializerMode.UNIFORM:
        return mode, (lo, up, 0.0, 0.0)
    if a.mode == DynamicEmbInitializerMode.NORMAL:
        return mode, (a.mean, a.std_dev, 0.0, 0.0)
    if a.mode == DynamicEmbInitializerMode.TRUNCATED_NORMAL:
        return mode, (a.mean, a.std_dev, lo, up)
    if a.mode == DynamicEmbInitializerMode.CONSTANT:
        return mode, (a.value, 0.0, 0.0, 0.0)
    return mode, (0.0, 0.0, 0.0, 0.0)


@dataclass
class PrefetchState:
    """What forward/backward need from prefetch (batched_dynamicemb_function.py:166-181)."""
    unique_keys: torch.Tensor
    reverse_indices: torch.Tensor
    unique_table_ids: torch.Tensor
    slot_indices: torch.Tensor      # table-local slot per unique key, -1 = insert failed
    rows: torch.Tensor              # global value row per unique key, -1 = absent
    num_unique_bound: int           # host-known upper bound of the unique count (sort width)
    num_unique_dev: Optional[torch.Tensor] = None   # device-side count (fused path: buffers are sized by the bound)
    bwd_ws: Optional[object] = None                 # ext.PreparedBackward: pre-sorted (unique idx, gradient row) pairs (backward_prepare)
    non_admitted_positions: Optional[torch.Tensor] = None   # unique-list positions of keys the admission strategy kept out (rows = -1)
    cold_slots: Optional[torch.Tensor] = None       # hybrid storage: slot / value row of the keys that live in the HOST tier (-1 for the
    cold_rows: Optional[torch.Tensor] = None        # others; `slot_indices` / `rows` are -1 for them)

    @property
    def num_unique(self) -> int:
        return int(self.num_unique_dev.item()) if self.num_unique_dev is not None else self.num_unique_bound


class _LookupFunction(torch.autograd.Function):
    """DynamicEmbeddingFunction (batched_dynamicemb_function.py:1043-1300): the optimizer update happens
    inside backward (fused); gradients returned to autograd are None."""

    @staticmethod
    def forward(ctx, module, state: PrefetchState, offsets, batch_size, dummy):
        pooled = module.pooling_mode != DynamicEmbPoolingMode.NONE
        combiner = int(module.pooling_mode) if pooled else -1
        n = state.reverse_indices.numel()
        out = ext.gather_forward(module._hot_values, module.max_D, state.rows, state.reverse_indices, n, offsets=offsets if pooled else None,
                                 batch_size=batch_size if pooled else 0, num_features=module.feature_num if pooled else 0,
                                 combiner=combiner, out_dtype=module.output_dtype)
        if state.cold_rows is not None:                             # hybrid storage: the rest of the rows sit in the host tier
            cold = ext.gather_forward(module._values, module.max_D, state.cold_rows, state.reverse_indices, n, offsets=offsets if pooled else None,
                                      batch_size=batch_size if pooled else 0, num_features=module.feature_num if pooled else 0,
                                      combiner=combiner, out_dtype=torch.float32, out=torch.empty(out.shape, dtype=torch.float32, device=out.device))
            out = out + cold if out.dtype == torch.float32 else (out.to(torch.float32) + cold).to(out.dtype)
        if state.non_admitted_positions is not None and state.non_admitted_positions.numel() > 0:
            out = module._add_non_admitted(out, state, offsets, batch_size, combiner)
        if module._mixed_D:
            out = out.index_select(1, module._col_index)          # [B, F*max_D] -> [B, total_D]
        ctx.module, ctx.state, ctx.offsets, ctx.batch_size, ctx.combiner = module, state, offsets, batch_size, combiner
        return out

    @staticmethod
    def backward(ctx, grads):
        m, st = ctx.module, ctx.state
        grads = grads.contiguous().to(torch.float32)
        opt = m._optimizer
        if opt.args.gradient_clipping:
            grads = grads.clamp(-opt.args.max_gradient, opt.args.max_gradient)
        opt.step()
        pooled = ctx.combiner >= 0
        if m._mixed_D:                                             # [B, total_D] -> [B, F*max_D], zeros in the padding columns
            grads = torch.zeros(grads.size(0), m.feature_num * m.max_D, dtype=torch.float32, device=grads.device).index_copy_(1, m._col_index, grads)
        ext.backward(m._hot_values, m.max_D, st.reverse_indices, max(st.num_unique_bound, 1), st.rows, grads, offsets=ctx.offsets if pooled else None,
                     batch_size=ctx.batch_size if pooled else 0, num_features=m.feature_num if pooled else 0, combiner=ctx.combiner,
                     prepared=st.bwd_ws, **opt.kernel_kwargs())
        if st.cold_rows is not None:                                # hybrid storage: the same reduce, applied to the host tier's rows
            ext.backward(m._values, m.max_D, st.reverse_indices, max(st.num_unique_bound, 1), st.cold_rows, grads,
                         offsets=ctx.offsets if pooled else None, batch_size=ctx.batch_size if pooled else 0,
                         num_features=m.feature_num if pooled else 0, combiner=ctx.combiner, **opt.kernel_kwargs())
        m._unpin(st)
        return None, None, None, None, None


class BatchedDynamicEmbeddingTablesV2(nn.Module):
    def __init__(
        self,
        table_options: List[DynamicEmbTableOptions],
        table_names: Optional[List[str]] = None,
        feature_table_map: Optional[List[int]] = None,
        use_index_dedup: bool = False,
        prefetch_pipeline: bool = False,
        pooling_mode: DynamicEmbPoolingMode = DynamicEmbPoolingMode.SUM,
        output_dtype: torch.dtype = torch.float32,
        device: torch.device = None,
        enforce_hbm: bool = False,
        bounds_check_mode: BoundsCheckMode = BoundsCheckMode.WARNING,
        optimizer: EmbOptimType = EmbOptimType.SGD,
        stochastic_rounding: bool = True,
        gradient_clipping: bool = False,
        max_gradient: float = 1.0,
        max_norm: float = 0.0,
        learning_rate: float = 0.01,
        eps: float = 1.0e-8,
        initial_accumulator_value: float = 0.0,
        momentum: float = 0.9,
        weight_decay: float = 0.0,
        weight_decay_mode=None,
        eta: float = 0.001,
        beta1: float = 0.9,
        beta2: float = 0.999,
        counter_based_regularization=None,
        cowclip_regularization=None,
        *args,
        **kwargs,
    ) -> None:
        super().__init__()
        assert len(table_options) >= 1
        opt0 = table_options[0]
        for o in table_options:
            assert o.get_grouped_key() == opt0.get_grouped_key(), "All tables must match in grouped keys."
            # one key map and one value tensor per module: these two also have to agree (the reference keeps a table object per dim)
            assert o.bucket_capacity == opt0.bucket_capacity and o.embedding_dtype == opt0.embedding_dtype, \
                "tables of one module share bucket_capacity and embedding_dtype"
            if o.external_storage is not None:
                raise NotImplementedError("external parameter-server storage is out of scope; see DESIGN.md")
            if o.embedding_dtype != torch.float32:
                raise NotImplementedError("value rows are fp32 in this build")
            if o.eval_initializer_args != opt0.eval_initializer_args:
                raise NotImplementedError("all tables of one module must share eval_initializer_args (one absent-row constant per fused eval lookup)")
        self._dynamicemb_options = table_options
        self.index_type = opt0.index_type
        self.embedding_dtype = opt0.embedding_dtype
        self.output_dtype = output_dtype
        self.pooling_mode = DynamicEmbPoolingMode(pooling_mode)
        self.use_index_dedup = use_index_dedup
        self._enable_prefetch = prefetch_pipeline
        self._table_names = table_names if table_names is not None else [f"t{i}" for i in range(len(table_options))]
        self.device_id = torch.device(device).index if device is not None and torch.device(device).index is not None else torch.cuda.current_device()
        self._device = torch.device("cuda", self.device_id)
        self.dims = [o.dim for o in table_options]
        # mixed embedding dims (pooled modules only, like the reference's D_offsets output, lookup_kernel.cuh:901-962): every table's rows
        # are stored max_D wide and the kernels run at max_D; the module slices / scatters the [B, F*max_D] pooled layout to the
        # [B, total_D] one the caller sees, so a narrower table's padding columns receive zero gradients and are never read
        self._mixed_D = any(d != self.dims[0] for d in self.dims)
        if self._mixed_D and DynamicEmbPoolingMode(pooling_mode) == DynamicEmbPoolingMode.NONE:
            raise NotImplementedError("mixed embedding dims need a pooling mode (sequence outputs have one width); split the tables by dim")
        T_ = len(table_options)
        self.feature_table_map = feature_table_map if feature_table_map is not None else list(range(T_))
        assert sorted(set(self.feature_table_map)) == list(range(T_)), "Each table must have at least one feature!"
        assert self.feature_table_map == sorted(self.feature_table_map), "features must be grouped by table"
        self.feature_num = len(self.feature_table_map)
        self.max_D = max(self.dims)
        self.total_D = sum(self.dims[t] for t in self.feature_table_map)
        offs, old = [], -1
        for i, t in enumerate(self.feature_table_map):
            if t != old:
                offs.append(i)
                old = t
        offs.append(self.feature_num)
        self.table_offsets_in_feature = offs
        self.feature_offsets = torch.tensor(offs, device=self._device, dtype=torch.int64)
        for o in table_options:
            if o.init_capacity is None:
                o.init_capacity = o.max_capacity
        self._optimizer_type = optimizer if opt0.training else EmbOptimType.NONE
        self._optimizer = SparseOptimizer(self._optimizer_type, OptimizerArgs(
            learning_rate=learning_rate, eps=eps, initial_accumulator_value=initial_accumulator_value, beta1=beta1, beta2=beta2,
            weight_decay=weight_decay, gradient_clipping=gradient_clipping, max_gradient=max_gradient))
        if self._mixed_D:
            if self._optimizer_type == EmbOptimType.EXACT_ROWWISE_ADAGRAD:
                raise NotImplementedError("row-wise Adagrad averages g^2 over the table's own dim; not available with mixed dims in one module")
            cols = [f * max(self.dims) + c for f, t in enumerate(self.feature_table_map) for c in range(self.dims[t])]
            self._col_index = torch.tensor(cols, dtype=torch.int64, device=self._device)       # [total_D] -> column of the max_D layout
        self._create_score()
        # --- storage: key index map + value rows [capacity, emb_dim + state_dim] in HBM (key_value_table.py:346-356)
        policy = self._score_policy()
        self._table = LinearBucketTable([o.max_capacity for o in table_options], [ScoreSpec(name="score", policy=policy)],   # LRU_LFU => 2 score words
                                        key_type=self.index_type, bucket_capacity=opt0.bucket_capacity, device=self._device)
        self.value_dim = self.max_D + self._optimizer.get_state_dim(self.max_D)
        self.value_dim = (self.value_dim + 3) // 4 * 4
        self._caching, self._hybrid, self._cache, self._cache_values = False, False, None, None
        total_bytes = sum(o.max_capacity * 4 * self.value_dim for o in table_options)
        local_hbm = sum(o.local_hbm_for_values for o in table_options)
        if any(o.caching for o in table_options):
            self._create_cache_storage(table_options, policy)
        elif 0 < local_hbm < total_bytes:
            self._create_hybrid_storage(table_options, policy, local_hbm / total_bytes)
        else:
            self._values = torch.zeros(self._table.capacity_, self.value_dim, dtype=torch.float32, device=self._device)
        self._seed = int(kwargs.get("seed", 0))
        # one initializer per table (reference: _create_initializers, batched_dynamicemb_tables.py:789-796): mode / bounds from that table's
        # initializer_args (default bound 1/sqrt(that table's capacity)), Philox seed mixed with the table id so the same key in two
        # tables does not get the same row
        self._init_per_table = []
        for t, o in enumerate(table_options):
            mode, p = _init_params(o.initializer_args, o.max_capacity)
            self._init_per_table.append((mode, p, (self._seed + t * 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF))
        self._table_init_dev = ext.make_table_init(self._init_per_table, self._device) if len(table_options) > 1 else None
        self._fused_prefetch = bool(kwargs.get("fused_prefetch", True))     # False = op-by-op path (reference op order, 2 host syncs)
        # admission (batched_dynamicemb_tables.py:526,624,798-812): table 0's strategy, one fused counter over all tables.  A module with
        # a strategy runs the op-by-op prefetch (the decision needs the missing keys on the host side of the op sequence, as in the
        # reference); without one nothing changes.
        self._admit_strategy = opt0.admit_strategy
        self._admission_counter = self._create_admission_counter(table_options)
        if self._admit_strategy is not None and self._admission_counter is None:
            raise ValueError("admit_strategy needs an admission_counter (KVCounter) in every table's options")
        self._bwd_prep = None                                               # ext.BackwardPrep, created on first training prefetch
        self._force_prepare = False
        self._prefetch_states: Deque[PrefetchState] = deque()
        self._empty_tensor = nn.Parameter(torch.empty(10, requires_grad=True, device=self._device, dtype=self.embedding_dtype))
        self.bounds_check_mode_int = int(bounds_check_mode)

    # the tier forward / backward read and write, and whose pin counters protect the rows of a prefetched batch
    @property
    def _hot_values(self) -> torch.Tensor:
        return self._cache_values if (self._caching or self._hybrid) else self._values

    @property
    def _hot_table(self) -> LinearBucketTable:
        return self._cache if (self._caching or self._hybrid) else self._table

    def _create_cache_storage(self, table_options, policy) -> None:
        """caching=True (batched_dynamicemb_tables.py:637-700): when the value rows do not fit `local_hbm_for_values`, an HBM cache of
        `capacity * local_hbm / total` rows per table (1024-slot buckets) fronts the full table, whose value rows are allocated in pinned
        host memory and read / written by the same row-copy kernels over the host link.  If everything fits, the module stays HBM-only."""
        assert all(o.caching for o in table_options), "caching is a grouped option: set it on every table of the module"
        if self._table.num_scores_ != 1:
            raise NotImplementedError("caching with the compound (TIMESTAMP, LFU) score is not built")
        total = sum(o.max_capacity * 4 * self.value_dim for o in table_options)
        local_hbm = sum(o.local_hbm_for_values for o in table_options)
        if total <= local_hbm:
            self._values = torch.zeros(self._table.capacity_, self.value_dim, dtype=torch.float32, device=self._device)
            return
        if local_hbm <= 0:
            raise ValueError("Can't use caching mode as the reserved HBM size is too small.")
        scale = local_hbm / total
        caps = [min(o.max_capacity, max(1, int(o.max_capacity * scale))) for o in table_options]
        # a NO_EVICTION backing table keeps its policy; its cache must be able to evict: LRU (reference :662-669)
        self._cache_policy = (ScorePolicy.GLOBAL_TIMER if table_options[0].score_strategy == DynamicEmbScoreStrategy.NO_EVICTION else policy)
        self._cache = LinearBucketTable(caps, [ScoreSpec(name="score", policy=self._cache_policy)], key_type=self.index_type,
                                        bucket_capacity=1024, device=self._device)
        self._cache_values = torch.zeros(self._cache.capacity_, self.value_dim, dtype=torch.float32, device=self._device)
        self._values = ext.host_values(self._table.capacity_, self.value_dim)
        self._caching = True

    def _create_hybrid_storage(self, table_options, policy, scale: float) -> None:
        """HybridStorage (key_value_table.py:2107-2404; module :701-742): 0 < local_hbm_for_values < table bytes WITHOUT caching = two
        disjoint tiers, an HBM table of `capacity * scale` rows per table (1024-slot buckets) and a host table of `capacity * (1 - scale)`
        rows whose value rows sit in pinned host memory.  New keys enter the HBM tier, what it evicts moves to the host tier, a key found
        in the host tier stays there: forward / fused backward run once per tier (rows of the other tier read zeros / are skipped)."""
        if self._table.num_scores_ != 1:
            raise NotImplementedError("hybrid storage with the compound (TIMESTAMP, LFU) score is not built")
        hbm_caps = [min(o.max_capacity, max(1, int(o.max_capacity * scale))) for o in table_options]
        host_caps = [min(o.max_capacity, max(1, int(o.max_capacity * (1.0 - scale)))) for o in table_options]
        self._cache_policy = (ScorePolicy.GLOBAL_TIMER if table_options[0].score_strategy == DynamicEmbScoreStrategy.NO_EVICTION else policy)
        self._cache = LinearBucketTable(hbm_caps, [ScoreSpec(name="score", policy=self._cache_policy)], key_type=self.index_type,
                                        bucket_capacity=1024, device=self._device)
        self._cache_values = torch.zeros(self._cache.capacity_, self.value_dim, dtype=torch.float32, device=self._device)
        self._table = LinearBucketTable(host_caps, [ScoreSpec(name="score", policy=policy)], key_type=self.index_type,
                                        bucket_capacity=table_options[0].bucket_capacity, device=self._device)
        self._values = ext.host_values(self._table.capacity_, self.value_dim)
        self._hybrid = True

    def _create_admission_counter(self, table_options):
        """One fused counter table for all tables (batched_dynamicemb_tables.py:798-812)."""
        counters = [o.admission_counter for o in table_options]
        if all(c is None for c in counters):
            return None
        assert all(c is not None for c in counters), "All tables must either have or not have an admission counter"
        from .embedding_admission import MultiTableKVCounter
        return MultiTableKVCounter(counters, device=self._device)

    def _add_non_admitted(self, out, state: PrefetchState, offsets, batch_size, combiner):
        """Ids whose key was not admitted read a freshly initialised row that is NOT stored (DynamicEmbeddingFunction.forward,
        batched_dynamicemb_function.py:1090-1097: the table's initializer over the non-admitted positions).  The rows are initialised
        into a scratch [n_na, value_dim] by the same init kernel, gathered / pooled by the same forward kernel (every other id reads
        zeros there) and added to the main result; the stored rows contributed zeros at those ids, so sequence outputs are exact."""
        na = state.non_admitted_positions
        n_na = na.numel()
        keys = state.unique_keys[na].contiguous()
        tids = state.unique_table_ids[na].contiguous() if state.unique_table_ids is not None else None
        scratch = torch.empty(n_na, self.value_dim, dtype=torch.float32, device=self._device)
        mode, p, seed0 = self._init_per_table[0]
        ext.init_rows(scratch, self.max_D, torch.arange(n_na, dtype=torch.int64, device=self._device), keys, mode, *p, seed=seed0, state_init=0.0,
                      table_ids=tids if self._table_init_dev is not None else None, table_init=self._table_init_dev)
        rows2 = torch.full((max(state.num_unique_bound, 1),), -1, dtype=torch.int64, device=self._device)
        rows2[na] = torch.arange(n_na, dtype=torch.int64, device=self._device)
        pooled = combiner >= 0
        n = state.reverse_indices.numel()
        extra = ext.gather_forward(scratch, self.max_D, rows2, state.reverse_indices, n, offsets=offsets if pooled else None,
                                   batch_size=batch_size if pooled else 0, num_features=self.feature_num if pooled else 0,
                                   combiner=combiner, out_dtype=torch.float32)
        return out + extra if out.dtype == torch.float32 else (out.to(torch.float32) + extra).to(out.dtype)

    # ------------------------------------------------------------------ scores (batched_dynamicemb_tables.py:1210-1260)
    def _create_score(self):
        self._scores: Dict[str, int] = {}
        for name, o in zip(self._table_names, self._dynamicemb_options):
            s = o.score_strategy
            if isinstance(s, tuple):
                # compound {TIMESTAMP, LFU} -> LruLfu: word 0 = last-access timestamp, word 1 = frequency, which drives eviction
                # (key_value_table.py:136-177)
                if frozenset(s) != frozenset({DynamicEmbScoreStrategy.TIMESTAMP, DynamicEmbScoreStrategy.LFU}) or len(s) != 2:
                    raise NotImplementedError(f"Unsupported compound score_strategy {s}.")
                o.evict_strategy = DynamicEmbEvictStrategy.LFU
                self._scores[name] = 1
            elif s == DynamicEmbScoreStrategy.TIMESTAMP:
                o.evict_strategy = DynamicEmbEvictStrategy.LRU
                self._scores[name] = 0
            elif s == DynamicEmbScoreStrategy.STEP:
                o.evict_strategy = DynamicEmbEvictStrategy.CUSTOMIZED
                self._scores[name] = 1
            elif s == DynamicEmbScoreStrategy.CUSTOMIZED:
                o.evict_strategy = DynamicEmbEvictStrategy.CUSTOMIZED
            elif s == DynamicEmbScoreStrategy.LFU:
                o.evict_strategy = DynamicEmbEvictStrategy.LFU
                self._scores[name] = 1
            elif s == DynamicEmbScoreStrategy.NO_EVICTION:
                o.evict_strategy = DynamicEmbEvictStrategy.CUSTOMIZED
                self._scores[name] = 0
            else:
                raise NotImplementedError(f"score strategy {s}")

    def _score_policy(self) -> ScorePolicy:
        """key_value_table.py:136-177 strategy -> policy."""
        s = self._dynamicemb_options[0].score_strategy
        if isinstance(s, tuple):
            return ScorePolicy.LRU_LFU
        if s == DynamicEmbScoreStrategy.TIMESTAMP:
            return ScorePolicy.GLOBAL_TIMER
        if s == DynamicEmbScoreStrategy.LFU:
            return ScorePolicy.ACCUMULATE
        return ScorePolicy.ASSIGN

    def _device_scores(self) -> torch.Tensor:
        """Per-table ASSIGN scores kept on the device so a training step never copies host data (STEP scores advance with add_)."""
        if getattr(self, "_scores_dev", None) is None:
            self._scores_dev = torch.tensor([self._scores[nm] & 0x7FFFFFFFFFFFFFFF for nm in self._table_names], dtype=torch.int64, device=self._device)
            self._step_mask = torch.tensor([1 if o.score_strategy == DynamicEmbScoreStrategy.STEP else 0 for o in self._dynamicemb_options],
                                           dtype=torch.int64, device=self._device)
        return self._scores_dev

    def _update_score(self):
        for name, o in zip(self._table_names, self._dynamicemb_options):
            if o.score_strategy == DynamicEmbScoreStrategy.STEP:
                self._scores[name] = (self._scores[name] + 1) & 0xFFFFFFFFFFFFFFFF
        if getattr(self, "_scores_dev", None) is not None:
            self._scores_dev.add_(self._step_mask)

    def set_score(self, named_score: Dict[str, int]) -> None:
        for name, score in named_score.items():
            if not isinstance(score, int):
                raise ValueError(f"Table's score is expect to int but got {type(score)}")
            if score == 0:
                raise ValueError("Can't set table's score to 0.")
            idx = self._table_names.index(name)
            assert self._dynamicemb_options[idx].score_strategy == DynamicEmbScoreStrategy.CUSTOMIZED, \
                "Can only set score for table whose score_strategy is DynamicEmbScoreStrategy.CUSTOMIZED."
            self._scores[name] = score
        self._scores_dev = None          # rebuilt lazily from the host copy

    def get_score(self) -> Dict[str, int]:
        out = {}
        for name, o in zip(self._table_names, self._dynamicemb_options):
            timed = o.score_strategy == DynamicEmbScoreStrategy.TIMESTAMP or isinstance(o.score_strategy, tuple)
            out[name] = ext.device_timestamp() if timed else self._scores[name]
        return out

    def _score_arg(self, n: int, table_ids: torch.Tensor, freq: Optional[torch.Tensor], const: bool = False, policy=None) -> ScoreArg:
        if const:
            return ScoreArg(name="score", policy=ScorePolicy.CONST)
        policy = self._score_policy() if policy is None else policy
        if policy == ScorePolicy.GLOBAL_TIMER:
            return ScoreArg(name="score", policy=policy)
        if policy in (ScorePolicy.ACCUMULATE, ScorePolicy.LRU_LFU):
            v = freq if freq is not None else torch.ones(n, dtype=torch.int64, device=self._device)
            return ScoreArg(name="score", value=v.to(torch.int64), policy=policy)
        for name in self._table_names:
            if name not in self._scores:
                raise RuntimeError(f"Must set score for table '{name}' whose score_strategy is customized.")
        per_table = torch.tensor([self._scores[nm] & 0x7FFFFFFFFFFFFFFF for nm in self._table_names], dtype=torch.int64, device=self._device)
        return ScoreArg(name="score", value=per_table[table_ids], policy=policy)

    # ------------------------------------------------------------------ properties the callers use
    @property
    def optimizer(self) -> SparseOptimizer:
        return self._optimizer

    @property
    def tables(self) -> LinearBucketTable:
        return self._table

    @property
    def table_names(self) -> List[str]:
        return self._table_names

    def set_learning_rate(self, lr: float) -> None:
        self._optimizer.set_learning_rate(lr)

    # --- the rest of the reference module's small public surface (batched_dynamicemb_tables.py:942-1000); the cache-tier entries are
    # no-ops without caching=True
    @property
    def cache(self):
        """The cache tier's table (None without caching; the HBM tier of a hybrid-storage module is not a cache)."""
        return self._cache if self._caching else None

    def reset_cache_states(self) -> None:
        """Empty the cache WITHOUT writing it back (reference :959-962); call flush() first to keep its rows."""
        if self._caching:
            self.reset_prefetch()
            self._cache.reset()

    def set_record_cache_metrics(self, record: bool) -> None:
        pass

    @property
    def enable_prefetch(self) -> bool:
        return self._enable_prefetch

    @enable_prefetch.setter
    def enable_prefetch(self, value: bool) -> None:
        self._enable_prefetch = value

    def split_embedding_weights(self) -> List[torch.Tensor]:
        """One placeholder per table, like the reference (:942-953): dynamic tables have no dense weight to split."""
        return [torch.empty((1, 1), device=self._device, dtype=self.embedding_dtype) for _ in self._dynamicemb_options]

    def fill_tables(self, load_factor: float = 0.95, tolerance: float = 1e-5) -> None:
        """Raise every table's occupancy to `load_factor` (clamped to 0.95) with uniformly random keys — only the key map is written,
        value rows stay as allocated (reference :1182-1209, DynamicEmbStorage.fill_tables key_value_table.py:1669-1790).  Batches of
        random keys are deduplicated, looked up, and the missing ones inserted with the table's current score until the target count (or
        `abs(load - load_factor) <= tolerance`) is reached."""
        if load_factor < 0.0:
            raise ValueError(f"load_factor must be non-negative, got {load_factor}")
        if tolerance < 0.0:
            raise ValueError(f"tolerance must be non-negative, got {tolerance}")
        load_factor = min(load_factor, 0.95)
        tb = self._table
        lfu = self._score_policy() in (ScorePolicy.ACCUMULATE, ScorePolicy.LRU_LFU)
        for t in range(len(self._dynamicemb_options)):
            cap = tb.capacity(t)
            if cap == 0:
                continue
            remaining = max(0, min(cap, int(load_factor * cap)) - tb.size(t))
            stale = 0
            while remaining > 0:
                n_gen = min(min(262144, max(4096, remaining * 4)), remaining * 8 + 1024)
                keys = torch.randint(0, torch.iinfo(torch.int64).max, (n_gen,), device=self._device, dtype=torch.int64).to(self.index_type)
                keys, counts = torch.unique(keys, return_counts=True)
                tids = torch.full((keys.numel(),), t, dtype=torch.int64, device=self._device)
                ts = ext.device_timestamp()
                _, founds, _ = tb.lookup(keys, tids, self._score_arg(keys.numel(), tids, counts if lfu else None), timestamp=ts)
                new = (~founds).nonzero(as_tuple=True)[0][:remaining]
                if new.numel() == 0:
                    stale += 1
                    if stale > 1000:
                        raise RuntimeError(f"fill_tables: stalled on table {t}")
                    continue
                stale = 0
                nk, nt = keys[new].contiguous(), tids[new].contiguous()
                tb.insert(nk, nt, self._score_arg(nk.numel(), nt, counts[new] if lfu else None), timestamp=ts)
                remaining -= int(nk.numel())
                if tolerance > 0.0 and abs(tb.size(t) / cap - load_factor) <= tolerance:
                    break

    def flush(self) -> None:
        """Write the cache back to the backing table (reference :955-957), then wait for the stream."""
        if self._caching:
            self._flush_cache()
        torch.cuda.current_stream(self._device).synchronize()

    def reset_prefetch(self) -> None:
        """Drop the queued prefetches: their rows are unpinned (ref_counter) so they can be evicted again."""
        while self._prefetch_states:
            self._unpin(self._prefetch_states.popleft())

    # ------------------------------------------------------------------ prefetch / forward
    def _split(self, indices: torch.Tensor, offsets: torch.Tensor):
        if indices.dtype != self.index_type:
            indices = indices.to(self.index_type)
        fb = offsets.numel() - 1
        assert fb > 0 and fb % self.feature_num == 0, "feature_batch_size must be divisible by feature_num"
        return indices.contiguous(), offsets.to(torch.int64).contiguous(), fb // self.feature_num

    def prefetch(self, indices, offsets, forward_stream=None, batch_size_per_feature_per_rank=None, frequency_counters=None) -> None:
        """dynamicemb_prefetch + _prefetch_hbm_direct_path: dedup, find, insert+init the missing keys, pin rows."""
        if not self.training:
            return
        indices, offsets, B = self._split(indices, offsets)
        T = len(self._dynamicemb_options)
        tb = self._table
        trange = ext.get_table_range(offsets, self.feature_offsets, self.feature_num) if T > 1 else None
        if self._fused_prefetch and self._admit_strategy is None and not self._caching and not self._hybrid:
            self._prefetch_fused(indices, trange, T, frequency_counters)
            return
        want_freq = self._score_policy() in (ScorePolicy.ACCUMULATE, ScorePolicy.LRU_LFU)
        freq_in = (frequency_counters.to(torch.int64) if frequency_counters is not None
                   else (torch.empty(0, dtype=torch.int64, device=self._device) if want_freq else None))
        num_u, ukeys, reverse, toffs, freq, utids = ext.segmented_unique_cuda(indices, trange, T, freq_in, want_table_ids=True)
        nu = int(num_u.item())                                   # host sync #1 (reference: batched_dynamicemb_function.py:141-142)
        ukeys, utids = ukeys[:nu], utids[:nu]
        freq = freq[:nu] if freq is not None else None
        if self._caching or self._hybrid:
            self._prefetch_cached(ukeys, utids, freq, reverse, nu)
            return
        ts = ext.device_timestamp()
        _, founds, slots = tb.lookup(ukeys, utids, self._score_arg(nu, utids, freq), timestamp=ts)
        tb.increment_counter(slots, utids)                       # pin found rows before anything can evict them (:607)
        miss = (~founds).nonzero(as_tuple=True)[0]               # host sync #2 (reference: flagged_compact)
        non_admitted = None
        if miss.numel() > 0 and self._admit_strategy is not None:
            # _prefetch_hbm_direct_path :615-641: count the missing keys, insert only the admitted ones; the others keep slot -1 (their
            # ids read an initialised row that is not stored, and their gradients are dropped)
            from .embedding_admission import admission_split
            admit_mask, _ = admission_split(ukeys[miss], utids[miss], freq[miss] if freq is not None else None, self._admit_strategy,
                                            self._admission_counter)
            non_admitted = miss[~admit_mask]
            miss = miss[admit_mask]
        if miss.numel() > 0:
            mk, mt = ukeys[miss], utids[miss]
            mf = freq[miss] if freq is not None else None
            new_slots = tb.insert(mk, mt, self._score_arg(mk.numel(), mt, mf), timestamp=ts)
            new_rows = ext.rows_from_slots(new_slots, mt, tb.row_base_)
            mode, p, seed0 = self._init_per_table[0]
            ext.init_rows(self._values, self.max_D, new_rows, mk, mode, *p, seed=seed0, state_init=self._optimizer.initial_state_value,
                          table_ids=mt if self._table_init_dev is not None else None, table_init=self._table_init_dev)
            tb.increment_counter(new_slots, mt)
            slots[miss] = new_slots
        rows = ext.rows_from_slots(slots, utids, tb.row_base_)
        self._prefetch_states.append(PrefetchState(ukeys, reverse, utids, slots, rows, nu, non_admitted_positions=non_admitted))
        self._update_score()

    def _prefetch_cached(self, ukeys, utids, freq, reverse, nu) -> None:
        """_prefetch_cache_path (batched_dynamicemb_function.py:296-556) on the table / row-copy kernels: cache lookup -> backing-table
        lookup of the misses (+ admission of keys found in neither) -> insert into the cache with eviction -> the evicted rows are read
        out of the cache BEFORE their slots are overwritten -> rows found in the backing table are copied into their cache slots, new keys
        initialised there -> evicted rows written back to the backing table -> pin.  Keys the cache cannot place (every slot of the
        bucket pinned) keep slot -1 like a failed insert of the HBM-direct path (the reference parks them in an overflow bucket)."""
        cache, st, V = self._cache, self._table, self.value_dim
        ts = ext.device_timestamp()
        _, founds, cslots = cache.lookup(ukeys, utids, self._score_arg(nu, utids, freq, policy=self._cache_policy), timestamp=ts)
        slots = cslots.clone()
        cache.increment_counter(cslots, utids)                  # pins the hits (slot -1 is skipped)
        miss = (~founds).nonzero(as_tuple=True)[0]
        non_admitted = None
        if miss.numel() > 0:
            mk, mt = ukeys[miss].contiguous(), utids[miss].contiguous()
            mf = freq[miss] if freq is not None else None
            s_score, s_found, s_slots = st.lookup(mk, mt, self._score_arg(mk.numel(), mt, mf), timestamp=ts)
            # cache: rows found in the backing table are PROMOTED into the cache; hybrid: they stay in the host tier (pinned there)
            promote = self._caching
            ins_mask = s_found.clone() if promote else torch.zeros_like(s_found)
            cold_slots = None
            if not promote:
                cold_slots = torch.full_like(slots, -1)
                cold_slots[miss] = s_slots                       # -1 where not found
                st.increment_counter(s_slots, mt)
            new_in_miss = ~s_found
            if bool(new_in_miss.any()):
                if self._admit_strategy is not None:
                    from .embedding_admission import admission_split
                    admit_mask, _ = admission_split(mk[new_in_miss], mt[new_in_miss], mf[new_in_miss] if mf is not None else None,
                                                    self._admit_strategy, self._admission_counter)
                    ins_mask[new_in_miss] = admit_mask
                    rejected = new_in_miss.clone()
                    rejected[new_in_miss] = ~admit_mask
                    if bool(rejected.any()):
                        non_admitted = miss[rejected]
                else:
                    ins_mask[new_in_miss] = True
            ins = ins_mask.nonzero(as_tuple=True)[0]
            if ins.numel() > 0:
                ik, it = mk[ins].contiguous(), mt[ins].contiguous()
                iscore = None
                if mf is not None:                               # a frequency score travels with the row: backing-table count + this batch
                    mf = torch.where(s_found, s_score, mf)
                    iscore = mf[ins]
                cidx, nev, ek, ei, es, et = cache.insert_and_evict(ik, it, self._score_arg(ik.numel(), it, iscore, policy=self._cache_policy),
                                                                   timestamp=ts)
                # A key inserted by this call can be evicted again by a later key of the SAME call when new keys carry the lowest score of
                # the bucket (LFU): its slot index is stale and its "evicted" record has no row yet.  A read-only lookup tells which
                # inserted keys really own their slot; the others count as failed inserts (slot -1) and their records are dropped.
                _, f2, s2 = cache.lookup(ik, it, ScoreArg(name="score", policy=ScorePolicy.CONST))
                lost = ~(f2 & (s2 == cidx)) & (cidx >= 0)
                keep = ei >= 0                                   # BUSY records (the new key itself, negative index) carry no row
                if bool(lost.any()):
                    cidx = torch.where(lost, torch.full_like(cidx, -1), cidx)
                    for t in range(len(self._dynamicemb_options)):
                        keep &= ~((et == t) & torch.isin(ek, ik[lost & (it == t)]))
                ek, ei, es, et = ek[keep].contiguous(), ei[keep].contiguous(), es[keep].contiguous(), et[keep].contiguous()
                ev_vals = torch.empty(ek.numel(), V, dtype=torch.float32, device=self._device)
                if ek.numel() > 0:
                    ext.copy_rows(self._cache_values, V, ext.rows_from_slots(ei, et, cache.row_base_), ev_vals, to_table=False)
                crow = ext.rows_from_slots(cidx, it, cache.row_base_)
                sf = s_found[ins]
                if bool(sf.any()):
                    tmp = torch.empty(int(sf.sum()), V, dtype=torch.float32, device=self._device)
                    ext.copy_rows(self._values, V, ext.rows_from_slots(s_slots[ins][sf].contiguous(), it[sf].contiguous(), st.row_base_), tmp,
                                  to_table=False)
                    ext.copy_rows(self._cache_values, V, crow[sf].contiguous(), tmp, to_table=True)
                nw = ~sf
                if bool(nw.any()):
                    mode, p, seed0 = self._init_per_table[0]
                    ext.init_rows(self._cache_values, self.max_D, crow[nw].contiguous(), ik[nw].contiguous(), mode, *p, seed=seed0,
                                  state_init=self._optimizer.initial_state_value,
                                  table_ids=it[nw].contiguous() if self._table_init_dev is not None else None, table_init=self._table_init_dev)
                if ek.numel() > 0:
                    self._write_back(ek, et, es, ev_vals, ts)
                cache.increment_counter(cidx, it)
                slots[miss[ins]] = cidx
        rows = ext.rows_from_slots(slots, utids, cache.row_base_)
        state = PrefetchState(ukeys, reverse, utids, slots, rows, nu, non_admitted_positions=non_admitted)
        if self._hybrid and miss.numel() > 0:
            state.cold_slots, state.cold_rows = cold_slots, ext.rows_from_slots(cold_slots, utids, st.row_base_)
        self._prefetch_states.append(state)
        self._update_score()

    def _write_back(self, keys, tids, scores, vals, ts) -> None:
        """Rows leaving the cache go to the backing table under their cache score (storage.insert of the evicted, :527-530)."""
        st = self._table
        sslots = st.insert(keys, tids, ScoreArg(name="score", value=scores.to(torch.int64).contiguous(), policy=ScorePolicy.ASSIGN), timestamp=ts)
        ext.copy_rows(self._values, self.value_dim, ext.rows_from_slots(sslots, tids, st.row_base_), vals, to_table=True)

    def _flush_cache(self) -> None:
        """flush_cache (key_value_table.py): every cached row is written to the backing table; the cache keeps its (now clean) content."""
        cache, V = self._cache, self.value_dim
        ts = ext.device_timestamp()
        for t in range(len(self._dynamicemb_options)):
            base = int(cache.table_bucket_offsets_cpu_[t]) * cache.bucket_capacity_
            for keys, scores, idx in cache.export(t):
                vals = torch.empty(keys.numel(), V, dtype=torch.float32, device=self._device)
                ext.copy_rows(self._cache_values, V, (idx + base).contiguous(), vals, to_table=False)
                tids = torch.full((keys.numel(),), t, dtype=torch.int64, device=self._device)
                self._write_back(keys, tids, scores, vals, ts)

    def _eval_forward_cached(self, indices, offsets, B) -> torch.Tensor:
        """Read-only lookup through both tiers (dynamicemb_eval_forward with a cache, :836-1040): an id reads its cached row, else its
        row of the backing table, else the eval initializer's constant.  Two gathers (ids absent from a tier read zeros there) + the constant."""
        cache, st, D = self._cache, self._table, self.max_D
        T = len(self._dynamicemb_options)
        n = indices.numel()
        if T > 1:
            trange = ext.get_table_range(offsets, self.feature_offsets, self.feature_num)
            tids = torch.repeat_interleave(torch.arange(T, device=self._device), trange[1:] - trange[:-1], output_size=n)
        else:
            tids = torch.zeros(n, dtype=torch.int64, device=self._device)
        const = ScoreArg(name="score", policy=ScorePolicy.CONST)
        _, cf, cs = cache.lookup(indices, tids, const)
        _, sf, ss = st.lookup(indices, tids, const)
        crow = ext.rows_from_slots(cs, tids, cache.row_base_)
        srow = ext.rows_from_slots(torch.where(cf, torch.full_like(ss, -1), ss), tids, st.row_base_)
        pooled = self.pooling_mode != DynamicEmbPoolingMode.NONE
        kw = dict(offsets=offsets if pooled else None, batch_size=B if pooled else 0, num_features=self.feature_num if pooled else 0,
                  combiner=int(self.pooling_mode) if pooled else -1, out_dtype=torch.float32)
        shape = (B, self.feature_num * D) if pooled else (n, D)
        new = lambda: torch.empty(shape, dtype=torch.float32, device=self._device)       # noqa: E731  (the host-resident tier cannot size `out`)
        out = ext.gather_forward(self._cache_values, D, crow, None, n, out=new(), **kw) + ext.gather_forward(self._values, D, srow, None, n, out=new(), **kw)
        ea = self._dynamicemb_options[0].eval_initializer_args
        absent = ea.value if ea.mode == DynamicEmbInitializerMode.CONSTANT else 0.0
        if absent != 0.0:
            fill = torch.empty(1, self.value_dim, dtype=torch.float32, device=self._device).fill_(absent)
            arow = torch.where(cf | sf, torch.full_like(cs, -1), torch.zeros_like(cs))
            out = out + ext.gather_forward(fill, D, arow, None, n, out=new(), **kw)
        if self._mixed_D:
            out = out.index_select(1, self._col_index)
        return out.to(self.output_dtype)

    def _unique_scratch(self, n: int) -> torch.Tensor:
        """Module-owned persistent dedup scratch (left clean by every call): no per-step re-initialisation of 32 B per id."""
        sc = getattr(self, "_uscratch", None)
        if sc is None or self._uscratch_n < n:
            assert not torch.cuda.is_current_stream_capturing(), "first use / growth of the dedup scratch must happen outside graph capture"
            self._uscratch_n = max(n, 1024)
            self._uscratch = sc = ext.unique_scratch(self._uscratch_n, len(self._dynamicemb_options), self._device)
        return sc

    def _prefetch_device_count(self, indices, trange, T, n_dev, scratch) -> PrefetchState:
        """Fused prefetch of `indices[:n_dev]` (count on the device; indices.numel() bounds it) — the owner side of the row-wise sharded
        step.  Returns the state instead of queueing it."""
        self._prefetch_fused(indices, trange, T, None, n_dev=n_dev, scratch=scratch, prepare=False)
        return self._prefetch_states.pop()

    def _prefetch_fused(self, indices, trange, T, frequency_counters, n_dev=None, scratch=None, prepare=True) -> None:
        """Same table image / rows as the op-by-op path, one C call, no host sync (csrc/demb_train.cu)."""
        tb = self._table
        if scratch is None:
            scratch = self._unique_scratch(indices.numel())
        policy = self._score_policy()
        scores = None
        if policy == ScorePolicy.ASSIGN:
            for name in self._table_names:
                if name not in self._scores:
                    raise RuntimeError(f"Must set score for table '{name}' whose score_strategy is customized.")
            scores = self._device_scores()
        mode, p, seed0 = self._init_per_table[0]
        uk, rev, utids, slots, rows, nu = ext.train_prefetch(
            tb.table_storage_, tb.table_bucket_offsets_, tb.bucket_capacity_, tb.bucket_sizes, tb._ref_counter, tb._bucket_heads, self._values,
            self.max_D, tb.row_base_, indices, trange, T, policy, scores, ext.device_timestamp(), mode, p, seed0,
            self._optimizer.initial_state_value, freq_in=frequency_counters, num_scores=tb.num_scores_, table_init=self._table_init_dev,
            n_dev=n_dev, unique_scratch=scratch)
        st = PrefetchState(uk, rev, utids if T > 1 else None, slots, rows, indices.numel(), nu)
        if prepare and ((self.training and self.pooling_mode == DynamicEmbPoolingMode.NONE and torch.is_grad_enabled()) or self._force_prepare):
            # the gradient-independent half of the fused backward (pair list + sort) starts now, on a side stream, under the forward gather
            if self._bwd_prep is None:
                self._bwd_prep = ext.BackwardPrep(self._device)
            st.bwd_ws = ext.backward_prepare(self._bwd_prep, self.max_D, rev, max(st.num_unique_bound, 1))
        self._prefetch_states.append(st)
        self._update_score()

    def _unpin(self, st: PrefetchState) -> None:
        if st.cold_slots is not None:
            self._table.decrement_counter(st.cold_slots, st.unique_table_ids)
        tb = self._hot_table
        if st.num_unique_dev is not None:
            ext.table_update_counter_n(tb._ref_counter, st.slot_indices, -1, tb.table_bucket_offsets_, tb.bucket_capacity_, st.num_unique_dev,
                                       table_ids=st.unique_table_ids)
        else:
            tb.decrement_counter(st.slot_indices, st.unique_table_ids)

    def forward(self, indices, offsets, per_sample_weights=None, feature_requires_grad=None, batch_size_per_feature_per_rank=None,
                total_unique_indices=None) -> torch.Tensor:
        indices, offsets, B = self._split(indices, offsets)
        if not self.training:
            return self._eval_forward(indices, offsets, B)
        if not self._prefetch_states:
            self.prefetch(indices, offsets, frequency_counters=per_sample_weights)
        state = self._prefetch_states.popleft()
        if not torch.is_grad_enabled():
            out = _LookupFunction.forward(_NoCtx(), self, state, offsets, B, None)
            self._unpin(state)
            return out
        return _LookupFunction.apply(self, state, offsets, B, self._empty_tensor)

    def _eval_forward(self, indices, offsets, B) -> torch.Tensor:
        """dynamicemb_eval_forward (batched_dynamicemb_function.py:836): read-only fused probe+gather; absent ids take the
        eval initializer (constant, default 0)."""
        if self._caching or self._hybrid:
            return self._eval_forward_cached(indices, offsets, B)
        tb = self._table
        T = len(self._dynamicemb_options)
        trange = ext.get_table_range(offsets, self.feature_offsets, self.feature_num) if T > 1 else None
        pooled = self.pooling_mode != DynamicEmbPoolingMode.NONE
        ea = self._dynamicemb_options[0].eval_initializer_args
        absent = ea.value if ea.mode == DynamicEmbInitializerMode.CONSTANT else 0.0
        out = ext.lookup_forward(tb.table_storage_, tb.table_bucket_offsets_, tb.bucket_capacity_, self._values, self.max_D, indices,
                                 row_base=tb.row_base_, table_range=trange, num_tables=T, offsets=offsets if pooled else None,
                                 batch_size=B if pooled else 0, num_features=self.feature_num if pooled else 0,
                                 combiner=int(self.pooling_mode) if pooled else -1, out_dtype=self.output_dtype, absent_value=absent,
                                 num_scores=tb.num_scores_)
        return out.index_select(1, self._col_index) if self._mixed_D else out

    # ------------------------------------------------------------------ CUDA-graph training step
    def make_graphed_step(self, ids_static: torch.Tensor, offsets: torch.Tensor, grad_static: torch.Tensor, with_loss: bool = True):
        """Capture prefetch -> forward -> loss stand-in -> fused backward as ONE CUDA graph over static buffers.

        Possible because the fused prefetch never synchronises with the host (unique counts stay on the device) and every kernel
        takes its sizes from device memory or from the static shapes.  Replay costs one launch: the ~0.3 ms of per-step Python /
        launch overhead disappears from the host timeline.  Returns (graph, out, loss): copy new ids into `ids_static` (and new
        upstream gradients into `grad_static`), call graph.replay(), read `out` / `loss` (loss = out.sum(), a stand-in for the model's
        scalar result; None with with_loss=False).
        Restrictions: training mode, fused prefetch, optimizers whose kernel arguments do not depend on the step count
        (SGD / Adagrad / row-wise Adagrad — Adam's bias correction is a host value), scores without a host timestamp (not TIMESTAMP /
        compound LRU-LFU).
        Side effects: the three warm-up steps before capture are REAL training steps on whatever `ids_static` / `grad_static` hold
        (they insert those ids, apply the optimizer and advance STEP scores) — pass the first real batch, or zero `grad_static`.
        The returned graph object's replay() also advances the host-side STEP scores and the optimizer's step count, so
        get_score() and a later eager step stay consistent with the device."""
        from .types import EmbOptimType
        assert self.training and self._fused_prefetch
        assert self._admit_strategy is None, "admission decides on the host side of the op sequence: no CUDA-graph step"
        assert not self._caching and not self._hybrid, "the tiered paths compact their misses on the host: no CUDA-graph step"
        assert not self._mixed_D, "mixed dims slice / scatter the pooled layout in torch: use forward() / backward()"
        assert self._optimizer_type in (EmbOptimType.SGD, EmbOptimType.EXACT_SGD, EmbOptimType.EXACT_ADAGRAD, EmbOptimType.EXACT_ROWWISE_ADAGRAD)
        assert self._score_policy() not in (ScorePolicy.GLOBAL_TIMER, ScorePolicy.LRU_LFU)
        indices, offsets_i, B = self._split(ids_static, offsets)
        assert indices.data_ptr() == ids_static.data_ptr(), "ids_static must already be contiguous int64"

        def step():
            # the same kernels as forward() + backward(), called directly: autograd would make the capture stream wait on
            # events of the streams that produced `grad_static` (cudaErrorStreamCaptureIsolation)
            self._force_prepare = self.pooling_mode == DynamicEmbPoolingMode.NONE     # no autograd here, but the backward does follow
            try:
                self.prefetch(indices, offsets_i)
            finally:
                self._force_prepare = False
            st = self._prefetch_states.popleft()
            out_ = _LookupFunction.forward(_NoCtx(), self, st, offsets_i, B, None)
            loss_ = out_.sum() if with_loss else None
            self._optimizer.step()
            pooled = self.pooling_mode != DynamicEmbPoolingMode.NONE
            ext.backward(self._values, self.max_D, st.reverse_indices, max(st.num_unique_bound, 1), st.rows, grad_static,
                         offsets=offsets_i if pooled else None, batch_size=B if pooled else 0, num_features=self.feature_num if pooled else 0,
                         combiner=int(self.pooling_mode) if pooled else -1, prepared=st.bwd_ws,
                         **self._optimizer.kernel_kwargs())
            self._unpin(st)
            return out_, loss_

        cur = torch.cuda.current_stream(self._device)
        side = torch.cuda.Stream(self._device)
        side.wait_stream(cur)
        with torch.no_grad(), torch.cuda.stream(side):
            for _ in range(3):                                    # warm-up on a side stream: workspaces, lazy attributes
                step()
        cur.wait_stream(side)
        torch.cuda.synchronize(self._device)
        graph = torch.cuda.CUDAGraph()
        host_scores, host_iter = dict(self._scores), self._optimizer.iter
        with torch.no_grad(), torch.cuda.graph(graph):
            out, loss = step()
        # capture ran step() on the host without executing it: undo its host-side bookkeeping, and roll the device STEP scores back too
        # (the captured add_ advances them at every replay)
        self._scores, self._optimizer.iter = host_scores, host_iter
        return _GraphedStep(self, graph, keepalive=(indices, offsets_i, grad_static)), out, loss

    # ------------------------------------------------------------------ export / checkpoint (batched_dynamicemb_tables.py:1262-1440)
    def _table_id(self, table) -> int:
        return self._table_names.index(table) if isinstance(table, str) else int(table)

    def _export_batches(self, table_id: int, batch_size: int = 65536):
        """(keys, value rows [n, value_dim], scores) batches of one table: table scan by the export kernel, rows by the row-copy kernel.
        scores: int64 [n], or [n, num_scores] in device order for multi-word scores (export_keys_values_iter, key_value_table.py:1073-1131)."""
        # a cache is flushed into the backing table before any export (flush()); hybrid storage holds disjoint key sets in its two tiers
        tiers = [(self._cache, self._cache_values), (self._table, self._values)] if self._hybrid else [(self._table, self._values)]
        for tb, values in tiers:
            base = int(tb.table_bucket_offsets_cpu_[table_id]) * tb.bucket_capacity_
            for keys, scores, idx in tb.export(table_id, batch=batch_size):
                dense = torch.empty(keys.numel(), self.value_dim, dtype=torch.float32, device=self._device)
                ext.copy_rows(values, self.value_dim, (idx + base).contiguous(), dense, to_table=False)
                sc = tb.gather_score_blocks(table_id, idx) if tb.num_scores_ > 1 else scores.view(torch.int64)
                yield keys, dense, sc

    def export_keys_values(self, table_name=0, device: Optional[torch.device] = None, batch_size: int = 65536):
        """All (keys, rows) of one table, named as in the reference (:1411, `table_name`) or by index.  With `device` given the reference's
        result is returned — (keys, embeddings [n, dim]) on that device; without it (keys, full value rows [n, dim + state]) on the GPU,
        which is what the tests inspect."""
        self.flush()
        t = self._table_id(table_name)
        ks, vs = [], []
        for keys, dense, _ in self._export_batches(t, batch_size):
            ks.append(keys)
            vs.append(dense)
        if not ks:
            ks, vs = [torch.empty(0, dtype=self.index_type, device=self._device)], [torch.empty(0, self.value_dim, device=self._device)]
        keys, vals = torch.cat(ks), torch.cat(vs)
        if device is None:
            return keys, vals
        return keys.to(device), vals[:, :self.dims[t]].contiguous().to(device)

    def incremental_dump(self, named_thresholds: Dict[str, int] = None, pg=None):
        """Keys / embeddings whose score is not below the table's threshold (batched_dynamicemb_tables.py:1432-1482,
        key_value_table.py:1977-2036).  Tables whose score carries a timestamp (TIMESTAMP, or the compound (TIMESTAMP, LFU): word 0)
        are thresholded on it — "touched since `threshold`", pass what the previous call returned; the others on the score value.
        Returns ({table: (keys, embeddings [n, dim])}, {table: next threshold}): CPU tensors; with a process group of more than one rank
        every rank gets all ranks' rows, concatenated in rank order.  The table scan is the export kernel, the filter a torch mask."""
        from . import checkpoint as ck
        import torch.distributed as dist
        self.flush()
        gather = pg is not None and dist.is_initialized() and dist.get_world_size(group=pg) > 1
        ret_tensors, ret_scores, ts = {}, {}, None
        for name, threshold in (named_thresholds or {}).items():
            if name not in self._table_names:
                warnings.warn(f"incremental_dump: table_name '{name}' is not in this module (available: {self._table_names}); skipping.",
                              UserWarning, stacklevel=2)
                continue
            t = self._table_names.index(name)
            D = self.dims[t]
            ks, vs = [], []
            for keys, dense, scores in self._export_batches(t):
                word = scores[:, 0] if scores.dim() == 2 else scores
                keep = word >= int(threshold)
                ks.append(keys[keep])
                vs.append(dense[keep][:, :D])
            keys = torch.cat(ks) if ks else torch.empty(0, dtype=self.index_type, device=self._device)
            vals = torch.cat(vs) if vs else torch.empty(0, D, dtype=torch.float32, device=self._device)
            ret_tensors[name] = ck.all_gather_keys_values(keys, vals, pg) if gather else (keys.cpu(), vals.cpu())
            s = self._dynamicemb_options[t].score_strategy
            if s == DynamicEmbScoreStrategy.TIMESTAMP or (isinstance(s, tuple) and DynamicEmbScoreStrategy.TIMESTAMP in s):
                ts = ext.device_timestamp() if ts is None else ts
                ret_scores[name] = ts
            else:
                ret_scores[name] = self._scores[name]
        return ret_tensors, ret_scores

    def _state_columns(self, dense: torch.Tensor, t: int) -> torch.Tensor:
        """[n, state_dim(dims[t])] optimizer state of table t out of full value rows.  A narrower table of a mixed-dim module keeps its
        state in blocks laid out for max_D (Adagrad: [M, M+D); Adam: m [M, M+D), v [2M, 2M+D))."""
        D, M = self.dims[t], self.max_D
        sdim = self._optimizer.get_state_dim(D)
        if D == M:
            return dense[:, D:D + sdim]
        return torch.cat([dense[:, (b + 1) * M:(b + 1) * M + D] for b in range(sdim // D)], dim=1)

    def _evict_strategy_str(self, table_id: int) -> str:
        # str() of the reference's pybind enum, which is what its meta json holds ("EvictStrategy.KLru", dynamic_emb_op.cu:814-819)
        return f"EvictStrategy.{self._dynamicemb_options[table_id].evict_strategy.value.name}"

    @staticmethod
    def _rank_world(pg):
        import torch.distributed as dist
        if pg is None and not dist.is_initialized():          # single process: the reference's dump asserts, its load treats it as 1 rank
            return 0, 1, False
        return dist.get_rank(group=pg), dist.get_world_size(group=pg), True

    def dump(self, save_dir: str, optim: bool = False, counter: bool = False, table_names: Optional[List[str]] = None, pg=None) -> None:
        """Write every selected table of this rank in the reference's file layout (dynamicemb/checkpoint.py)."""
        from . import checkpoint as ck
        import torch.distributed as dist
        names = set(self._table_names if table_names is None else table_names)
        rank, world, distributed = self._rank_world(pg)
        os.makedirs(save_dir, exist_ok=True)
        self.flush()
        for t, name in enumerate(self._table_names):
            if name not in names:
                continue
            if distributed:
                dist.barrier(group=pg)
            ts = ext.device_timestamp()
            o = self._dynamicemb_options[t]
            if rank == 0:
                meta = dict(self._optimizer.get_opt_args())
                meta["evict_strategy"] = self._evict_strategy_str(t)
                meta["dist_type"] = o.dist_type
                if self._scores.get(name) is not None:
                    meta["step_score"] = self._scores[name]
                ck.save_to_json(meta, ck.encode_meta_json_file_path(save_dir, name))
            D = self.dims[t]
            sdim = self._optimizer.get_state_dim(D)
            perm = ck.score_dump_permutation(o.score_strategy)
            lru = o.evict_strategy == DynamicEmbEvictStrategy.LRU
            path = lambda item: ck.encode_checkpoint_file_path(save_dir, name, rank, world, item)     # noqa: E731
            with ck.TableFileWriter(path("keys"), path("values"), path("scores"), path("opt_values") if optim else None) as w:
                for keys, dense, scores in self._export_batches(t):
                    if scores.dim() == 2 and perm != list(range(scores.size(1))):
                        scores = scores[:, perm].contiguous()
                    if lru:
                        scores = ts - scores                                     # age; restored relative to the loading time
                    opt = None
                    if optim and sdim > 0:
                        opt = ck.truncate_optimizer_states_for_checkpoint(self._optimizer, D, self._state_columns(dense, t))
                    w.write(keys, dense[:, :D], scores, opt)
            if counter:
                if self._admission_counter is not None:
                    self._admission_counter.dump(ck.encode_counter_checkpoint_file_path(save_dir, name, rank, world, "keys"),
                                                 ck.encode_counter_checkpoint_file_path(save_dir, name, rank, world, "frequencies"), t)
                else:
                    warnings.warn(f"Counter table is none and will not dump it for table: {name}")

    def load(self, save_dir: str, optim: bool = False, counter: bool = False, table_names: Optional[List[str]] = None, pg=None) -> None:
        """Read a checkpoint in the reference's layout: this rank's own files when the world size matches, otherwise every file filtered
        to the keys this rank owns."""
        from . import checkpoint as ck
        import torch.distributed as dist
        names = set(self._table_names if table_names is None else table_names)
        rank, world, distributed = self._rank_world(pg)
        if self._caching:
            self.reset_cache_states()          # the files replace the backing table's content: cached copies would be stale
        for t, name in enumerate(self._table_names):
            if name not in names:
                continue
            kf, vf, sf, of, ckf, cff = ck.get_loading_files(save_dir, name, rank=rank, world_size=world)
            if not kf:
                continue
            if distributed:
                dist.barrier(group=pg)
            ts = ext.device_timestamp()
            own_files = len(kf) == 1 and kf[0] == ck.encode_checkpoint_file_path(save_dir, name, rank, world, "keys")
            for i in range(len(kf)):
                loaded = self._load_table_files(t, ck.encode_meta_json_file_path(save_dir, name), kf[i], vf[i], sf[i] if sf else None,
                                                of[i] if of else None, include_optim=optim, timestamp=ts,
                                                filter_rank=None if own_files or world == 1 else (rank, world))
                if loaded is not None and name in self._scores:
                    self._scores[name] = loaded
                    self._scores_dev = None
            if counter:
                if self._admission_counter is None:
                    warnings.warn(f"Counter table is none and will not load for table: {name}")
                    continue
                for i in range(len(ckf)):
                    self._admission_counter.load(ckf[i], cff[i], t)

    def _load_table_files(self, t: int, meta_path, key_path, value_path, score_path, opt_path, include_optim, timestamp, filter_rank=None):
        """DynamicEmbStorage.load + _load_key_values (key_value_table.py:1909-1976, :1402-1517) on the native ops: insert the keys with the
        stored scores (ASSIGN), copy [emb | state] rows to the slots they got."""
        from . import checkpoint as ck
        tb, o, D = self._table, self._dynamicemb_options[t], self.dims[t]
        ns = tb.num_scores_
        meta = ck.load_from_json(meta_path)
        if score_path is None:
            print(f"Score file {score_path} does not exist. Will not load score states.")
        include_optim, file_sdim, _ = ck.validate_load_meta(meta, self._optimizer, self._evict_strategy_str(t), o.dist_type, D, ns, key_path,
                                                           value_path, score_path, opt_path, include_optim)
        sdim = self._optimizer.get_state_dim(D)
        perm = ck.score_load_permutation(o.score_strategy) if ns > 1 else None
        lru = o.evict_strategy == DynamicEmbEvictStrategy.LRU
        base = int(tb.table_bucket_offsets_cpu_[t]) * tb.bucket_capacity_
        rank, world = filter_rank if filter_rank is not None else (0, 1)
        for keys, emb, scores, opt in ck.iter_batches_from_files(key_path, value_path, score_path, opt_path if include_optim else None, D,
                                                                 file_sdim, self._device, num_scores=ns, rank=rank, world_size=world,
                                                                 dist_type=o.dist_type):
            n = keys.numel()
            if n == 0:
                continue
            if scores is not None and perm is not None and perm != list(range(ns)):
                scores = scores[:, perm].contiguous()
            if scores is not None and lru:
                scores = torch.clamp(timestamp - scores, min=0)
            dense = torch.empty(n, self.value_dim, dtype=torch.float32, device=self._device)
            state = None
            if sdim > 0:
                state = (self._optimizer.initial_state_value if opt is None else
                         ck.pad_optimizer_states_from_checkpoint(self._optimizer, D, opt, self._optimizer.initial_state_value, torch.float32,
                                                                 self._device))
            if D != self.max_D:                                     # narrower table of a mixed-dim module: blocks laid out for max_D
                dense.zero_()
                dense[:, :D] = emb
                if sdim > 0:
                    for b in range(sdim // D):
                        lo = (b + 1) * self.max_D
                        dense[:, lo:lo + D] = state if opt is None else state[:, b * D:(b + 1) * D]
            else:
                dense[:, :D] = emb
                if sdim > 0:
                    dense[:, D:D + sdim] = state
                if self.value_dim > D + sdim:
                    dense[:, D + sdim:] = 0
            tids = torch.full((n,), t, dtype=torch.int64, device=self._device)
            keys = keys.to(self.index_type)
            if self._hybrid:                                        # the file is authoritative: no second copy in the HBM tier
                self._cache.erase(keys, tids)
            if ns > 1:
                if scores is None or scores.dim() != 2 or scores.size(1) != ns or scores.size(0) != n:
                    raise ValueError(f"multi-word load expects [{n}, {ns}] scores, got {None if scores is None else tuple(scores.shape)}")
                slots = tb.insert(keys, tids, ScoreArg(name="score", policy=ScorePolicy.CONST), timestamp=timestamp)
                tb.scatter_score_blocks(t, slots, scores)
            elif scores is None:
                assert lru, "scores is None is only allowed for the LRU evict strategy"
                slots = tb.insert(keys, tids, ScoreArg(name="score", policy=ScorePolicy.GLOBAL_TIMER), timestamp=timestamp)
            else:
                slots = tb.insert(keys, tids, ScoreArg(name="score", value=scores.to(torch.int64).contiguous(), policy=ScorePolicy.ASSIGN),
                                  timestamp=timestamp)
            rows = torch.where(slots >= 0, slots + base, slots).contiguous()
            ext.copy_rows(self._values, self.value_dim, rows, dense, to_table=True)
        return meta.get("step_score", None)


class _GraphedStep:
    """graph.replay() + the host-side bookkeeping of one training step (STEP scores, optimizer step count)."""

    def __init__(self, module: BatchedDynamicEmbeddingTablesV2, graph: torch.cuda.CUDAGraph, keepalive=()):
        # the graph holds RAW POINTERS of every tensor its kernels read: anything created outside the capture (offsets, converted ids)
        # must live as long as the graph does
        self.module, self.graph, self._keepalive = module, graph, tuple(keepalive)

    def replay(self) -> None:
        self.graph.replay()
        m = self.module
        for name, o in zip(m._table_names, m._dynamicemb_options):
            if o.score_strategy == DynamicEmbScoreStrategy.STEP:
                m._scores[name] = (m._scores[name] + 1) & 0xFFFFFFFFFFFFFFFF
        m._optimizer.iter += 1


class _NoCtx:
    """Stand-in ctx so the no-grad path reuses _LookupFunction.forward."""
    pass
