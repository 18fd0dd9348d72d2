"""`get_planner` — the convenience constructor the reference's examples call (corelib/dynamicemb/dynamicemb/get_planner.py:60-131): per
table a `DynamicEmbParameterConstraints` (data-parallel tables / DynamicEmb tables row-wise with `use_dynamicemb=True` / other
model-parallel tables row-wise), a topology and a `DynamicEmbeddingShardingPlanner` over them.  With torchrec installed the topology and
the storage reservation are TorchRec's; without it a plain record with the same fields stands in (this package's planner reads only the
world size from it and plans the DynamicEmb tables itself)."""
from dataclasses import dataclass
from typing import Dict, List, Set

import torch
import torch.distributed as dist

from .planner import DynamicEmbeddingEnumerator, DynamicEmbeddingShardingPlanner, DynamicEmbParameterConstraints
from .types import BoundsCheckMode, DynamicEmbTableOptions

_pipeline_type_to_model_parallel_allowed_compute_kernels = {"prefetch": ["fused_uvm_caching"], "native": ["fused", "fused_uvm"], "none": []}
_pipeline_type_to_data_parallel_allowed_compute_kernels = {"prefetch": ["dense"], "native": ["dense"], "none": []}


@dataclass
class _Topology:
    """The fields of torchrec.distributed.planner.Topology the reference sets (:108-116)."""
    local_world_size: int
    world_size: int
    compute_device: str
    hbm_cap: int
    ddr_cap: int
    intra_host_bw: float
    inter_host_bw: float


def get_planner(eb_configs: List, data_parallel_embedding_table_names: Set[str], dynamicemb_options_dict: Dict[str, DynamicEmbTableOptions],
                device: torch.device, pipeline_type: str = "none", ddr_cap: int = 512 * 1024 * 1024 * 1024, intra_host_bw: int = 450e9,
                inter_host_bw: int = 25e9):
    constraints = {}
    for config in eb_configs:
        if config.name in data_parallel_embedding_table_names:
            constraints[config.name] = DynamicEmbParameterConstraints(
                sharding_types=["data_parallel"], bounds_check_mode=BoundsCheckMode.NONE, use_dynamicemb=False,
                compute_kernels=_pipeline_type_to_data_parallel_allowed_compute_kernels[pipeline_type])
        elif config.name in dynamicemb_options_dict:
            constraints[config.name] = DynamicEmbParameterConstraints(
                sharding_types=["row_wise"], bounds_check_mode=BoundsCheckMode.NONE, enforce_hbm=True, use_dynamicemb=True,
                dynamicemb_options=dynamicemb_options_dict[config.name])
        else:
            constraints[config.name] = DynamicEmbParameterConstraints(
                sharding_types=["row_wise"], bounds_check_mode=BoundsCheckMode.NONE, use_dynamicemb=False,
                compute_kernels=_pipeline_type_to_model_parallel_allowed_compute_kernels[pipeline_type])
    world = dist.get_world_size() if dist.is_initialized() else 1
    hbm_cap = torch.cuda.get_device_properties(0).total_memory if torch.cuda.is_available() else 0
    storage_reservation = None
    try:
        from torchrec.distributed.comm import get_local_size
        from torchrec.distributed.planner import Topology
        from torchrec.distributed.planner.storage_reservations import HeuristicalStorageReservation
        topology = Topology(local_world_size=get_local_size(), world_size=world, compute_device=device.type, hbm_cap=hbm_cap, ddr_cap=ddr_cap,
                            intra_host_bw=intra_host_bw, inter_host_bw=inter_host_bw)
        storage_reservation = HeuristicalStorageReservation(percentage=0.05)
    except ImportError:
        local = int(torch.cuda.device_count()) if torch.cuda.is_available() else 1
        topology = _Topology(min(local, world) or 1, world, device.type, hbm_cap, ddr_cap, intra_host_bw, inter_host_bw)
    enumerator = DynamicEmbeddingEnumerator(topology=topology, constraints=constraints)
    return DynamicEmbeddingShardingPlanner(eb_configs=eb_configs, topology=topology, constraints=constraints, enumerator=enumerator,
                                           storage_reservation=storage_reservation)
