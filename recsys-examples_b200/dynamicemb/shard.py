"""Row-wise model-parallel wrapper around BatchedDynamicEmbeddingTablesV2.

`RowWiseShardedDynamicEmbedding` is the executable equivalent of what TorchRec builds from the reference's sharders
(/root/reference/corelib/dynamicemb/dynamicemb/shard/embedding.py:78 ShardedDynamicEmbeddingCollection.input_dist :277 ->
compute -> output_dist; planner/rw_sharding.py:83,189): optional per-rank index dedup -> block bucketize -> all_to_all(lengths, ids)
-> local lookup -> all_to_all(rows) -> un-bucketize (-> un-dedup).  Each rank owns an independent hash table of
get_sharded_table_capacity(num_embeddings, world_size) rows; NCCL all_to_all over NVLink is the only collective.

The TorchRec-facing class names of the reference (DynamicEmbeddingCollectionSharder, ...) are kept for drop-in imports; they need
torchrec at construction time (not installed in the build image) and delegate to this wrapper.
"""
from typing import List, Optional

import torch
import torch.distributed as dist
from torch import nn

from . import dynamicemb_extensions as ext
from .batched_dynamicemb_tables import BatchedDynamicEmbeddingTablesV2
from .input_dist import rw_sharded_lookup
from .types import DynamicEmbPoolingMode


class RowWiseShardedDynamicEmbedding(nn.Module):
    """Sequence-mode (EmbeddingCollection) row-wise sharded lookup.  forward(ids[n], lengths[F*B]) -> [n, D] rows in id order."""

    def __init__(self, local: BatchedDynamicEmbeddingTablesV2, process_group=None, dist_type: str = "hash_roundrobin",
                 num_embeddings_per_feature: Optional[List[int]] = None, use_index_dedup: bool = True):
        super().__init__()
        assert local.pooling_mode == DynamicEmbPoolingMode.NONE, "sequence mode wrapper"
        self.local = local
        self.group = process_group
        self.world_size = dist.get_world_size(process_group)
        self.use_index_dedup = use_index_dedup
        F = local.feature_num
        dev = local._device
        self._dist_type = torch.full((F,), ext.DIST_TYPE[dist_type], dtype=torch.int32, device=dev)
        if num_embeddings_per_feature is None:
            num_embeddings_per_feature = [1 << 62] * F
        # continuous: block size = ceil(hash_size / W) (input_dist.py:231-234)
        self._block_sizes = torch.tensor([(n + self.world_size - 1) // self.world_size for n in num_embeddings_per_feature], dtype=torch.int64, device=dev)

    def _bucketize(self, lengths, ids):
        B = lengths.numel() // self.local.feature_num
        nl, ni, _, perm = ext.block_bucketize_sparse_features(lengths, ids, B, self.world_size, self._block_sizes, self._dist_type, sequence=True)
        return nl, ni, perm

    def _unique(self, ids, trange, F):
        num_u, uk, reverse, toffs, _ = ext.segmented_unique_cuda(ids, trange, F, None)
        return int(num_u.item()), uk, reverse, toffs                 # host sync: the exchange sizes depend on it

    def forward(self, ids: torch.Tensor, lengths: torch.Tensor) -> torch.Tensor:
        return rw_sharded_lookup(ids, lengths, self.local.feature_num, self.group, local_fn=self.local, bucketize_fn=self._bucketize,
                                 unique_fn=self._unique if self.use_index_dedup else None, reduce_fn=self._reduce_rows)

    @staticmethod
    def _reduce_rows(idx, grad, num_rows):
        """sum of the gradient rows per source row: the reduce_grads kernels (sort -> tiles -> windows -> spans, fixed order)"""
        return ext.reduce_grads(idx, grad, num_rows, 0, grad.shape[1])


def _need_torchrec():
    try:
        import torchrec  # noqa: F401
    except ImportError as e:
        raise ImportError("the TorchRec-facing sharder classes need `torchrec` (release/V1.5.0 in the reference image, docker/Dockerfile:33); "
                          "use RowWiseShardedDynamicEmbedding directly when torchrec is not installed") from e


class DynamicEmbeddingCollectionSharder:
    """shard/embedding.py:343 — (qcomm_codecs_registry, use_index_dedup, fused_params)."""

    def __init__(self, qcomm_codecs_registry=None, use_index_dedup: bool = False, fused_params=None):
        _need_torchrec()
        self.qcomm_codecs_registry, self.use_index_dedup, self.fused_params = qcomm_codecs_registry, use_index_dedup, fused_params or {}


class DynamicEmbeddingBagCollectionSharder:
    """shard/embeddingbag.py:79"""

    def __init__(self, qcomm_codecs_registry=None, fused_params=None):
        _need_torchrec()
        self.qcomm_codecs_registry, self.fused_params = qcomm_codecs_registry, fused_params or {}


class DynamicEmbParameterConstraints:
    """planner/planner.py: TorchRec ParameterConstraints + use_dynamicemb + dynamicemb_options."""

    def __init__(self, use_dynamicemb: bool = False, dynamicemb_options=None, **kwargs):
        self.use_dynamicemb, self.dynamicemb_options, self.kwargs = use_dynamicemb, dynamicemb_options, kwargs


class DynamicEmbeddingShardingPlanner:
    """planner/planner.py:213"""

    def __init__(self, *args, **kwargs):
        _need_torchrec()
        self.args, self.kwargs = args, kwargs


ShardedDynamicEmbedding = RowWiseShardedDynamicEmbedding
