"""`dynamicemb.dynamicemb_config` — import path of the reference's configuration module (dynamicemb_config.py); the definitions live in
`dynamicemb.types`."""
from .types import (BATCH_SIZE_PER_DUMP, DEFAULT_BUCKET_CAPACITY, DEFAULT_INDEX_TYPE, SUPPORTED_DIST_TYPES, BoundsCheckMode, DynamicEmbCheckMode,  # noqa: F401
                    DynamicEmbDataType, DynamicEmbEvictStrategy, DynamicEmbInitializerArgs, DynamicEmbInitializerMode, DynamicEmbPoolingMode,
                    DynamicEmbScoreStrategy, DynamicEmbTableOptions, ScoreStrategy, align_to_table_size, data_type_to_dtype, data_type_to_dyn_emb,
                    dtype_to_bytes, dyn_emb_to_torch, get_sharded_table_capacity, get_table_value_bytes, normalize_score_strategy,
                    string_to_evict_strategy)
