"""Config enums / dataclasses of the DynamicEmb boundary.  Field names and defaults follow
/root/reference/corelib/dynamicemb/dynamicemb/{types.py:33-110, dynamicemb_config.py:62-165,448-478}
so user code written against the reference constructs the same objects."""
import abc
import enum
import math
from dataclasses import dataclass, field
from typing import Any, Optional, Tuple, Union

import torch

from .dynamicemb_extensions import EvictStrategy

DEFAULT_INDEX_TYPE = torch.int64
DEFAULT_BUCKET_CAPACITY = 128
BUCKET_ALIGNMENT = 16
DEMB_TABLE_ALIGN_SIZE = 16
MAX_BUCKET_CAPACITY: int = 2 ** 63 - 1        # sentinel bucket_capacity: the whole per-rank shard is one bucket (types.py:118-120)
DEBUG_EMB_INITIALIZER_MOD = 100_000
BATCH_SIZE_PER_DUMP = 65536
SUPPORTED_DIST_TYPES = ("continuous", "roundrobin", "hash_roundrobin")
DTYPE_NUM_BYTES = {torch.float32: 4, torch.float16: 2, torch.bfloat16: 2}


class EmbOptimType(enum.Enum):
    """Subset of fbgemm_gpu.split_embedding_configs.EmbOptimType the reference dispatches on
    (dynamicemb/optimizer.py:36-57); string values are fbgemm's."""
    SGD = "sgd"
    EXACT_SGD = "exact_sgd"
    ADAM = "adam"
    EXACT_ADAGRAD = "exact_adagrad"
    EXACT_ROWWISE_ADAGRAD = "exact_row_wise_adagrad"
    NONE = "none"


class DynamicEmbInitializerMode(enum.Enum):
    NORMAL = "normal"
    TRUNCATED_NORMAL = "truncated_normal"
    UNIFORM = "uniform"
    CONSTANT = "constant"
    DEBUG = "debug"


@dataclass
class DynamicEmbInitializerArgs:
    mode: DynamicEmbInitializerMode = DynamicEmbInitializerMode.UNIFORM
    mean: float = 0.0
    std_dev: float = 1.0
    lower: Optional[float] = None
    upper: Optional[float] = None
    value: float = 0.0


@enum.unique
class DynamicEmbCheckMode(enum.IntEnum):
    ERROR = 0
    WARNING = 1
    IGNORE = 2


class DynamicEmbPoolingMode(enum.IntEnum):
    SUM = 0
    MEAN = 1
    NONE = 2


@enum.unique
class DynamicEmbEvictStrategy(enum.Enum):
    LRU = EvictStrategy.KLru
    LFU = EvictStrategy.KLfu
    EPOCH_LRU = EvictStrategy.KEpochLru
    EPOCH_LFU = EvictStrategy.KEpochLfu
    CUSTOMIZED = EvictStrategy.KCustomized


class DynamicEmbScoreStrategy(enum.IntEnum):
    TIMESTAMP = 0
    STEP = 1
    CUSTOMIZED = 2
    LFU = 3
    NO_EVICTION = 4


ScoreStrategy = Union[DynamicEmbScoreStrategy, Tuple[DynamicEmbScoreStrategy, ...]]


class BoundsCheckMode(enum.IntEnum):
    FATAL = 0
    WARNING = 1
    IGNORE = 2
    NONE = 3


@dataclass
class DynamicEmbTableOptions:
    """dynamicemb_config.py:307-520 (fields :448-478)."""
    embedding_dtype: Optional[torch.dtype] = None
    dim: Optional[int] = None
    max_capacity: Optional[int] = None
    evict_strategy: DynamicEmbEvictStrategy = DynamicEmbEvictStrategy.LRU
    local_hbm_for_values: int = 0
    device_id: Optional[int] = None
    training: bool = True
    initializer_args: DynamicEmbInitializerArgs = field(default_factory=DynamicEmbInitializerArgs)
    eval_initializer_args: DynamicEmbInitializerArgs = field(
        default_factory=lambda: DynamicEmbInitializerArgs(mode=DynamicEmbInitializerMode.CONSTANT, value=0.0))
    caching: bool = False
    init_capacity: Optional[int] = None
    max_load_factor: float = 0.5
    score_strategy: Optional[ScoreStrategy] = DynamicEmbScoreStrategy.TIMESTAMP
    bucket_capacity: int = DEFAULT_BUCKET_CAPACITY
    safe_check_mode: DynamicEmbCheckMode = DynamicEmbCheckMode.IGNORE
    global_hbm_for_values: int = 0
    external_storage: Any = None
    index_type: Optional[torch.dtype] = None
    dist_type: str = "roundrobin"
    admit_strategy: Any = None
    admission_counter: Any = None

    def __post_init__(self):
        # reference :480-492: the eval initializer is a constant, dist_type is one of the three routings, score_strategy is canonical
        assert self.eval_initializer_args.mode == DynamicEmbInitializerMode.CONSTANT, "eval_initializer_args must be constant initialization"
        if self.index_type is None:
            self.index_type = DEFAULT_INDEX_TYPE
        if self.embedding_dtype is None:
            self.embedding_dtype = torch.float32
        if self.dist_type not in SUPPORTED_DIST_TYPES:
            raise ValueError(f"Unsupported dist_type {self.dist_type!r}. Supported values: {SUPPORTED_DIST_TYPES}.")
        self.score_strategy = normalize_score_strategy(self.score_strategy)

    # tables are grouped into one module by these fields only (reference :494-520); options compare / hash by them
    def get_grouped_key(self):
        return {"training": self.training, "caching": self.caching, "external_storage": self.external_storage, "index_type": self.index_type,
                "dist_type": self.dist_type, "score_strategy": self.score_strategy, "admit_strategy": self.admit_strategy}

    def __eq__(self, other):
        if not isinstance(other, DynamicEmbTableOptions):
            return NotImplemented
        return self.get_grouped_key() == other.get_grouped_key()

    def __ne__(self, other):
        if not isinstance(other, DynamicEmbTableOptions):
            return NotImplemented
        return not (self == other)

    def __hash__(self):
        return hash(tuple(self.get_grouped_key().items()))


SUPPORTED_COMPOUND_SCORE_STRATEGIES = (frozenset({DynamicEmbScoreStrategy.TIMESTAMP, DynamicEmbScoreStrategy.LFU}),)


def normalize_score_strategy(score_strategy):
    """dynamicemb_config.py:166-212: None passes through; `(X,)` -> X; a longer tuple must be a supported compound set without duplicates."""
    if score_strategy is None:
        return None
    if isinstance(score_strategy, tuple):
        for element in score_strategy:
            if not isinstance(element, DynamicEmbScoreStrategy):
                raise TypeError(f"score_strategy tuple elements must be DynamicEmbScoreStrategy, got {type(element)}.")
        if len(score_strategy) == 0:
            raise NotImplementedError("score_strategy tuple must be non-empty.")
        if len(score_strategy) == 1:
            return score_strategy[0]
        if len(score_strategy) != len(frozenset(score_strategy)) or frozenset(score_strategy) not in SUPPORTED_COMPOUND_SCORE_STRATEGIES:
            raise NotImplementedError(f"Unsupported compound score_strategy {score_strategy}.")
        return score_strategy
    if not isinstance(score_strategy, DynamicEmbScoreStrategy):
        raise TypeError(f"score_strategy must be a DynamicEmbScoreStrategy or a tuple of them, got {type(score_strategy)}.")
    return score_strategy


def align_to_table_size(n: int, alignment: int) -> int:
    return ((int(n) + alignment - 1) // alignment) * alignment


def _sharded_table_bucket_layout(embedding_config, world_size: int, bucket_capacity: int) -> Tuple[int, int]:
    """(num_buckets, effective bucket width) of one rank's shard (dynamicemb_config.py:696-731).  `embedding_config` is the reference's
    argument — a TorchRec table config, of which only `.num_embeddings` is read — or the row count itself."""
    if world_size <= 0:
        raise ValueError(f"world_size must be positive, got {world_size}")
    num_global = int(getattr(embedding_config, "num_embeddings", embedding_config))
    shard_rows = math.ceil(num_global / world_size)
    if bucket_capacity == MAX_BUCKET_CAPACITY:               # sentinel: the whole shard is one bucket
        return 1, align_to_table_size(shard_rows, BUCKET_ALIGNMENT)
    if bucket_capacity <= 0:
        raise ValueError(f"bucket_capacity must be positive when not MAX_BUCKET_CAPACITY, got {bucket_capacity}")
    if bucket_capacity % BUCKET_ALIGNMENT != 0:
        raise ValueError(f"bucket_capacity ({bucket_capacity}) must be a multiple of BUCKET_ALIGNMENT ({BUCKET_ALIGNMENT}) when not using "
                         "MAX_BUCKET_CAPACITY.")
    return align_to_table_size(shard_rows, bucket_capacity) // bucket_capacity, bucket_capacity


def get_sharded_table_capacity(embedding_config, world_size: int, bucket_capacity: int = DEFAULT_BUCKET_CAPACITY) -> int:
    """dynamicemb_config.py:733-765: shard_rows = ceil(N/W); capacity = round_up(shard_rows, bucket_capacity) — what the planner writes to
    `DynamicEmbTableOptions.max_capacity`."""
    num_buckets, width = _sharded_table_bucket_layout(embedding_config, world_size, bucket_capacity)
    return int(num_buckets * width)


def dtype_to_bytes(dtype: torch.dtype) -> int:
    return torch.empty((), dtype=dtype).element_size()


def get_table_value_bytes(embedding_config, optimizer_type: EmbOptimType, world_size: int, bucket_capacity: int = DEFAULT_BUCKET_CAPACITY) -> int:
    """Bytes of value storage (embedding + optimizer state) of one table over all ranks (dynamicemb_config.py:768-803)."""
    from .optimizer import get_optimizer_state_dim
    total_rows = get_sharded_table_capacity(embedding_config, world_size, bucket_capacity) * world_size
    dim = embedding_config.embedding_dim
    dtype = data_type_to_dtype(embedding_config.data_type) if hasattr(embedding_config, "data_type") else torch.float32
    return int(dtype_to_bytes(dtype) * (dim + get_optimizer_state_dim(optimizer_type, dim, dtype)) * total_rows)


class DynamicEmbDataType(enum.IntEnum):
    """src/utils.h:31-40 (pybind enum of the reference)."""
    Float32 = 0
    Float16 = 1
    BFloat16 = 2
    Int64 = 3
    UInt64 = 4
    Int32 = 5
    UInt32 = 6
    Size_t = 7


_TORCH_OF_DYN = {DynamicEmbDataType.Float32: torch.float32, DynamicEmbDataType.BFloat16: torch.bfloat16, DynamicEmbDataType.Float16: torch.float16,
                 DynamicEmbDataType.Int64: torch.int64, DynamicEmbDataType.UInt64: torch.uint64, DynamicEmbDataType.Int32: torch.int32,
                 DynamicEmbDataType.UInt32: torch.uint32, DynamicEmbDataType.Size_t: torch.int64}
# torchrec.types.DataType member NAME -> (torch dtype, DynamicEmbDataType); matched by name so torchrec is not needed at import time
_OF_TORCHREC = {"FP32": (torch.float32, DynamicEmbDataType.Float32), "FP16": (torch.float16, DynamicEmbDataType.Float16),
                "BF16": (torch.bfloat16, DynamicEmbDataType.BFloat16), "INT64": (torch.int64, DynamicEmbDataType.Int64),
                "INT32": (torch.int32, DynamicEmbDataType.Int32), "INT8": (torch.int8, None), "UINT8": (torch.uint8, None)}


def dyn_emb_to_torch(data_type: DynamicEmbDataType) -> torch.dtype:
    """dynamicemb_config.py:562-580"""
    if data_type not in _TORCH_OF_DYN:
        raise ValueError(f"Unsupported DynamicEmbDataType: {data_type}")
    return _TORCH_OF_DYN[data_type]


def torch_to_dyn_emb(torch_dtype: torch.dtype) -> DynamicEmbDataType:
    """utils.py:41-57"""
    for k, v in _TORCH_OF_DYN.items():
        if v == torch_dtype and k != DynamicEmbDataType.Size_t:
            return k
    raise ValueError(f"Unsupported torch dtype: {torch_dtype}")


def data_type_to_dtype(data_type) -> torch.dtype:
    """TorchRec DataType -> torch dtype (dynamicemb_config.py:543-559)."""
    name = getattr(data_type, "name", str(data_type)).upper()
    if name not in _OF_TORCHREC:
        raise ValueError(f"DataType {data_type} cannot be converted to torch.dtype")
    return _OF_TORCHREC[name][0]


def data_type_to_dyn_emb(data_type) -> DynamicEmbDataType:
    """TorchRec DataType -> DynamicEmbDataType (dynamicemb_config.py:522-540)."""
    name = getattr(data_type, "name", str(data_type)).upper()
    if name not in _OF_TORCHREC or _OF_TORCHREC[name][1] is None:
        raise ValueError(f"DataType {data_type} cannot be converted to DynamicEmbDataType")
    return _OF_TORCHREC[name][1]


def string_to_evict_strategy(strategy_str: str) -> EvictStrategy:
    """dynamicemb_config.py:604-616"""
    if strategy_str not in ("KLru", "KLfu", "KEpochLru", "KEpochLfu", "KCustomized"):
        raise ValueError(f"Invalid EvictStrategy string: {strategy_str}")
    return EvictStrategy[strategy_str]


class MemoryType(enum.Enum):
    """types.py:26-30 of the reference."""
    DEVICE = "device"
    MANAGED = "managed"
    PINNED_HOST = "pinned_host"
    HOST = "host"


class Counter(abc.ABC):
    """key -> counter map with logical tables (reference types.py:333-398); `MultiTableKVCounter` is the shipped one."""

    @abc.abstractmethod
    def add(self, keys: torch.Tensor, table_ids: torch.Tensor, frequencies: torch.Tensor) -> torch.Tensor:
        """Add `frequencies` to the (unique) `keys`; returns the accumulated counter of every key."""

    @abc.abstractmethod
    def erase(self, keys: torch.Tensor, table_ids: torch.Tensor) -> None:
        ...

    @abc.abstractmethod
    def memory_usage(self, mem_type=MemoryType.DEVICE) -> int:
        ...

    @abc.abstractmethod
    def load(self, key_file, counter_file, table_id: int) -> None:
        ...

    @abc.abstractmethod
    def dump(self, key_file, counter_file, table_id: int) -> None:
        ...


class AdmissionStrategy(abc.ABC):
    """Decides which missing keys may enter the table (reference types.py:401-420)."""

    @abc.abstractmethod
    def admit(self, keys: torch.Tensor, frequencies: torch.Tensor) -> torch.Tensor:
        """Boolean mask over `keys`: True = insert."""

    @abc.abstractmethod
    def initialize_non_admitted_embeddings(self, buffer: torch.Tensor, indices: torch.Tensor) -> bool:
        """Fill the rows of keys that were not admitted; False = nothing done (the table's initializer is used)."""
