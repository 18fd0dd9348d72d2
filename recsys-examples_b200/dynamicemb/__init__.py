"""dynamicemb — B200-native DynamicEmb hot path behind the reference's operator surface.

Import name and public symbols mirror /root/reference/corelib/dynamicemb/dynamicemb/__init__.py for
the parts on the hot path: the TBE-shaped module, its options, the scored hash table
and the TorchRec sharding wrapper (imported lazily: torchrec is optional at import time).
"""
from . import dynamicemb_extensions  # noqa: F401  (raises if librecsys_b200.so is missing: no CPU fallback)
from .batched_dynamicemb_tables import BatchedDynamicEmbeddingTablesV2
from .dynamicemb_extensions import EvictStrategy, InsertResult, ScorePolicy
from .optimizer import OptimizerArgs, SparseOptimizer, get_optimizer_state_dim
from .scored_hashtable import LinearBucketTable, ScoreArg, ScoreSpec, get_scored_table, murmur3_hash_64bits
from .embedding_admission import FrequencyAdmissionStrategy, KVCounter, MultiTableKVCounter
from .dump_load import DynamicEmbDump, DynamicEmbLoad      # get_score / set_score / incremental_dump: `dynamicemb.incremental_dump`, as in the reference
from .types import (BATCH_SIZE_PER_DUMP, BUCKET_ALIGNMENT, DEMB_TABLE_ALIGN_SIZE, MAX_BUCKET_CAPACITY, DynamicEmbDataType, ScoreStrategy,
                    align_to_table_size, data_type_to_dtype, data_type_to_dyn_emb, dyn_emb_to_torch, get_table_value_bytes,
                    string_to_evict_strategy, torch_to_dyn_emb)
from .types import (AdmissionStrategy, BoundsCheckMode, Counter, DynamicEmbCheckMode, DynamicEmbEvictStrategy, DynamicEmbInitializerArgs, DynamicEmbInitializerMode,
                    DynamicEmbPoolingMode, DynamicEmbScoreStrategy, DynamicEmbTableOptions, EmbOptimType, get_sharded_table_capacity)

__all__ = [
    "BatchedDynamicEmbeddingTablesV2", "DynamicEmbTableOptions", "DynamicEmbInitializerArgs", "DynamicEmbInitializerMode",
    "DynamicEmbPoolingMode", "DynamicEmbScoreStrategy", "DynamicEmbEvictStrategy", "DynamicEmbCheckMode", "EmbOptimType",
    "BoundsCheckMode", "LinearBucketTable", "ScoreArg", "ScoreSpec", "ScorePolicy", "InsertResult", "EvictStrategy",
    "get_scored_table", "murmur3_hash_64bits", "get_sharded_table_capacity", "OptimizerArgs", "SparseOptimizer",
    "get_optimizer_state_dim", "DynamicEmbDump", "DynamicEmbLoad", "BATCH_SIZE_PER_DUMP",
    "BUCKET_ALIGNMENT", "DEMB_TABLE_ALIGN_SIZE", "MAX_BUCKET_CAPACITY", "DynamicEmbDataType", "ScoreStrategy", "align_to_table_size",
    "data_type_to_dtype", "data_type_to_dyn_emb", "dyn_emb_to_torch", "get_table_value_bytes", "string_to_evict_strategy", "torch_to_dyn_emb",
    "AdmissionStrategy", "Counter", "FrequencyAdmissionStrategy", "KVCounter", "MultiTableKVCounter",
]


def __getattr__(name):
    # TorchRec-facing classes need torchrec; resolve them on first use (shard/, planner/ in the reference).
    if name in ("DynamicEmbeddingCollectionSharder", "DynamicEmbeddingBagCollectionSharder", "DynamicEmbeddingShardingPlanner",
                "DynamicEmbParameterConstraints", "ShardedDynamicEmbedding", "RowWiseShardedDynamicEmbedding"):
        from . import shard
        return getattr(shard, name)
    raise AttributeError(name)
