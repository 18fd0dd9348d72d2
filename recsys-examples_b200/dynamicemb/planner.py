"""`dynamicemb.planner` — import path of the reference's planner package (corelib/dynamicemb/dynamicemb/planner/__init__.py:17-30).
The planner and the constraints class live in `dynamicemb.shard`; the two remaining names of the reference's `__all__` are here."""
from dataclasses import dataclass, field, fields
from typing import Dict, List, Optional

from .shard import DynamicEmbeddingShardingPlanner, DynamicEmbParameterConstraints  # noqa: F401
from .types import DynamicEmbTableOptions

DynamicEmbKernel = "DynamicEmb"          # planner.py:71 customized_compute_kernel name


@dataclass
class DynamicEmbParameterSharding:
    """One table's entry of a sharding plan (planner/planner.py:82-106; TorchRec's ParameterSharding fields first).  The plans of
    `DynamicEmbeddingShardingPlanner.plan()` are dictionaries with these keys; `from_plan_entry` turns one into this object."""
    sharding_type: str = "row_wise"
    compute_kernel: str = "customized_kernel"
    ranks: Optional[List[int]] = None
    sharding_spec: Optional[object] = None
    cache_params: Optional[object] = None
    enforce_hbm: Optional[bool] = None
    stochastic_rounding: Optional[bool] = None
    bounds_check_mode: Optional[object] = None
    output_dtype: Optional[object] = None
    key_value_params: Optional[object] = None
    customized_compute_kernel: Optional[str] = DynamicEmbKernel
    dist_type: str = "roundrobin"
    dynamicemb_options: Optional[DynamicEmbTableOptions] = field(default_factory=DynamicEmbTableOptions)

    _ADDED = ("customized_compute_kernel", "dist_type", "dynamicemb_options")

    def get_additional_fused_params(self) -> Dict[str, object]:
        """The fields DynamicEmb adds to ParameterSharding (what the reference forwards as fused params, :97-106)."""
        return {f.name: getattr(self, f.name) for f in fields(self) if f.name in self._ADDED}

    @classmethod
    def from_plan_entry(cls, entry: Dict[str, object]) -> "DynamicEmbParameterSharding":
        known = {f.name for f in fields(cls)}
        return cls(**{k: v for k, v in entry.items() if k in known})


class DynamicEmbeddingEnumerator:
    """planner/enumerators.py:207-237: TorchRec's EmbeddingEnumerator restricted to what DynamicEmb tables allow (row-wise sharding, the
    customized kernel).  Constructor arguments are the reference's; enumerating a module's sharding options is TorchRec's machinery and
    needs torchrec — the row-wise plan itself comes from `DynamicEmbeddingShardingPlanner.plan()` without it."""

    def __init__(self, topology, batch_size: Optional[int] = 512, constraints: Optional[Dict[str, DynamicEmbParameterConstraints]] = None,
                 estimator=None, use_exact_enumerate_order: Optional[bool] = False) -> None:
        self._topology, self._batch_size, self._constraints = topology, batch_size, constraints
        self._estimator, self._use_exact_enumerate_order = estimator, use_exact_enumerate_order

    def enumerate(self, module, sharders):
        from .shard import _need_torchrec
        _need_torchrec()
        raise NotImplementedError("sharding-option enumeration is TorchRec's; DynamicEmb tables are planned row-wise by DynamicEmbeddingShardingPlanner")


__all__ = ["DynamicEmbeddingEnumerator", "DynamicEmbeddingShardingPlanner", "DynamicEmbParameterConstraints", "DynamicEmbParameterSharding"]
