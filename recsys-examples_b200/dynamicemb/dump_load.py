"""Model-level checkpoint and score API: walk a model, find its sharded dynamic-embedding collections, call the modules' own
dump / load / set_score / get_score / incremental_dump.

Mirrors /root/reference/corelib/dynamicemb/dynamicemb/dump_load.py:31-287 (find_sharded_modules, get_dynamic_emb_module, DynamicEmbDump,
DynamicEmbLoad) and incremental_dump.py:25-348 (set_score, get_score, incremental_dump): same function names, arguments, directory layout
(`<path>/<collection path>/<table>_emb_keys.rank_r.world_size_W` ..., the collection path being the dotted module path prefixed with
"model") and the same nesting of the returned dictionaries.  What the reference finds by type (TorchRec's ShardedEmbeddingCollection /
ShardedEmbeddingBagCollection) is found here by type too: this package's row-wise sharded wrappers, which play that role, plus any module
that carries the marker attribute `_is_dynamicemb_collection` (an adapter around TorchRec's classes can set it).
"""
import logging
import os
import warnings
from collections import deque
from typing import Any, Dict, List, Optional, Set, Tuple, Union

import torch
import torch.distributed as dist
from torch import nn

from .batched_dynamicemb_tables import BatchedDynamicEmbeddingTablesV2


def _is_collection(module: nn.Module) -> bool:
    from . import shard
    return isinstance(module, (shard.RowWiseShardedDynamicEmbedding, shard.RowWiseShardedDynamicEmbeddingA2A)) or \
        bool(getattr(module, "_is_dynamicemb_collection", False))


def _unwrap(module: nn.Module) -> nn.Module:
    """get_unwrapped_module of TorchRec: DistributedModelParallel / DDP / Float16Module keep the wrapped model in `.module`."""
    while hasattr(module, "module") and isinstance(getattr(module, "module"), nn.Module) and not _is_collection(module):
        module = module.module
    return module


def find_sharded_modules(module: nn.Module, path: str = "") -> List[Tuple[str, str, nn.Module]]:
    """(collection path, attribute name, module) of every sharded embedding collection under `module` (dump_load.py:31-50)."""
    found: List[Tuple[str, str, nn.Module]] = []
    stack = deque([(module, path + "model", "model")])
    while stack:
        cur, cur_path, cur_name = stack.pop()
        cur = _unwrap(cur)
        if _is_collection(cur):
            found.append((cur_path, cur_name, cur))
        else:
            for name, child in cur.named_children():
                stack.append((child, cur_path + ("." if cur_path else "") + name, name))
    return found


def check_emb_collection_modules(module: nn.Module, ret_list: List[nn.Module], visited: Optional[Set[int]] = None) -> None:
    """Collect the BatchedDynamicEmbeddingTablesV2 modules under a collection (dump_load.py:53-88): nn.Module children plus the private
    containers TorchRec keeps as plain lists."""
    visited = set() if visited is None else visited
    if id(module) in visited:
        return
    visited.add(id(module))
    if isinstance(module, BatchedDynamicEmbeddingTablesV2):
        ret_list.append(module)
        return
    if isinstance(module, nn.Module):
        for attr in ("_lookups", "_emb_modules", "_emb_module"):
            child = getattr(module, attr, None)
            if child is None:
                continue
            for item in (child if isinstance(child, (list, nn.ModuleList)) else [child]):
                check_emb_collection_modules(item, ret_list, visited)
        for child in module.children():
            check_emb_collection_modules(child, ret_list, visited)


def get_dynamic_emb_module(model: nn.Module) -> List[nn.Module]:
    out: List[nn.Module] = []
    check_emb_collection_modules(model, out)
    return out


def _barrier(pg) -> None:
    if dist.is_initialized():
        dist.barrier(group=pg)


def _sync() -> None:
    if torch.cuda.is_available():
        torch.cuda.synchronize()


def DynamicEmbDump(path: str, model: nn.Module, table_names: Optional[Dict[str, List[str]]] = None, optim: Optional[bool] = False,
                   counter: Optional[bool] = False, pg=None, allow_overwrite: bool = False) -> None:
    """Dump every dynamic embedding table of `model` (and optimizer state / admission counters on request) under
    `path/<collection path>/` in the per-rank file layout of `module.dump` (dump_load.py:103-203).  `table_names` maps a collection
    path to the tables to dump (default: all).  A non-empty `path` is refused unless `allow_overwrite`."""
    _sync()
    if not os.path.exists(path):
        try:
            os.makedirs(path, exist_ok=True)
        except Exception as e:
            raise Exception("can't build path:", path) from e
    elif not os.path.isdir(path):
        raise Exception(f"The path '{path}' exists and is not a directory.")
    elif os.listdir(path):
        if not allow_overwrite:
            raise Exception(f"DynamicEmb Cannot dump to {path} because it already contains files, "
                            "as it may cause overwriting of existing files with the same name.")
        logging.warning(f"DynamicEmb Overwriting existing files in {path}")
    _barrier(pg)
    collections = find_sharded_modules(model, "")
    if len(collections) == 0:
        warnings.warn("Input model don't have any sharded dynamic embedding collection, will not dump any embedding tables to filesystem!",
                      UserWarning)
        return
    for collection_path, _, _ in collections:
        os.makedirs(os.path.join(path, collection_path), exist_ok=True)
    _barrier(pg)
    for collection_path, _, collection in collections:
        names = table_names.get(collection_path, None) if table_names else None
        for m in get_dynamic_emb_module(collection):
            m.dump(os.path.join(path, collection_path), optim=optim, counter=counter, table_names=names, pg=pg)
    _sync()
    _barrier(pg)


def DynamicEmbLoad(path: str, model: nn.Module, table_names: Optional[Dict[str, List[str]]] = None, optim: bool = False, counter: bool = False,
                   pg=None) -> None:
    """Load what DynamicEmbDump wrote (by this package or by the reference) into the model's tables (dump_load.py:211-287)."""
    _sync()
    if not os.path.exists(path):
        raise Exception("can't find path to load, path:", path)
    collections = find_sharded_modules(model, "")
    if len(collections) == 0:
        warnings.warn("Input model don't have any sharded dynamic embedding collection, can't load any embedding tables from filesystem!",
                      UserWarning)
        return
    for collection_path, _, collection in collections:
        names = table_names.get(collection_path, None) if table_names else None
        for m in get_dynamic_emb_module(collection):
            m.load(os.path.join(path, collection_path), optim=optim, counter=counter, table_names=names, pg=pg)
    _sync()


# ---------------------------------------------------------------------------------------------------------------------------------
# scores and incremental dump over a whole model (incremental_dump.py)
def is_valid_score_threshold(score_threshold: Any) -> bool:
    """`Dict[str, Dict[str, int]]`?  (incremental_dump.py:25-44)"""
    if not isinstance(score_threshold, dict):
        return False
    for key, value in score_threshold.items():
        if not isinstance(key, str) or not isinstance(value, dict):
            return False
        for inner_key, inner_value in value.items():
            if not isinstance(inner_key, str) or not isinstance(inner_value, int):
                return False
    return True


def _collections_with_tables(model: nn.Module, what: str):
    collections = find_sharded_modules(model, "")
    if len(collections) == 0:
        warnings.warn(f"Input model don't have any sharded dynamic embedding collection, can't {what}!", UserWarning)
        return None
    if not any(get_dynamic_emb_module(c) for _, _, c in collections):
        warnings.warn(f"Input model don't have any Dynamic embedding tables, can't {what}!", UserWarning)
        return None
    return collections


def set_score(model: nn.Module, table_score: Union[int, Dict[str, Dict[str, int]]]) -> None:
    """Register the score the next forward passes assign (CUSTOMIZED tables): one integer for every table, or
    {collection NAME: {table: score}} (incremental_dump.py:47-150; the reference keys this call by the collection's attribute name)."""
    if isinstance(table_score, int):
        set_all = True
    elif is_valid_score_threshold(table_score):
        set_all = False
    else:
        raise ValueError("DynamicEmb Error:table_score should be int or Dict")
    collections = _collections_with_tables(model, "set score")
    if collections is None:
        return
    if not set_all:
        present = {name for _, name, _ in collections}
        for wanted in table_score:
            if wanted not in present:
                warnings.warn(f"sharded module '{wanted}' specified in table_score not found in the model", UserWarning)
        collections = [c for c in collections if c[1] in table_score]
    for _, name, collection in collections:
        for m in get_dynamic_emb_module(collection):
            if set_all:
                scores = {t: table_score for t in m.table_names}
            else:
                scores = {t: s for t, s in table_score[name].items() if t in m.table_names}
            if scores:
                m.set_score(scores)


def get_score(model: nn.Module) -> Optional[Dict[str, Dict[str, int]]]:
    """{collection path: {table: score}} — device timer for TIMESTAMP tables, the step for STEP, the registered score for CUSTOMIZED
    (incremental_dump.py:153-214)."""
    collections = _collections_with_tables(model, "get score")
    if collections is None:
        return None
    out: Dict[str, Dict[str, int]] = {}
    for collection_path, _, collection in collections:
        scores: Dict[str, int] = {}
        for m in get_dynamic_emb_module(collection):
            scores.update(m.get_score())
        out[collection_path] = scores
    return out


def incremental_dump(model: nn.Module, score_threshold: Union[int, Dict[str, Dict[str, int]]], pg=None):
    """Keys / embeddings whose score is not below the threshold, for every table (int) or the listed ones
    ({collection PATH: {table: threshold}}): ({collection path: {table: (keys, values)}}, {collection path: {table: next threshold}})
    (incremental_dump.py:217-348)."""
    if isinstance(score_threshold, int):
        set_all = True
    elif is_valid_score_threshold(score_threshold):
        set_all = False
    else:
        raise ValueError("DynamicEmb Error:score_threshold should be int or Dict")
    collections = _collections_with_tables(model, "incremental dump")
    if collections is None:
        return None
    if not set_all:
        present = {p for p, _, _ in collections}
        for wanted in score_threshold:
            if wanted not in present:
                warnings.warn(f"sharded module '{wanted}' specified in score_threshold not found in the model", UserWarning)
        collections = [c for c in collections if c[0] in score_threshold]
    ret_tensors: Dict[str, Dict[str, Tuple[torch.Tensor, torch.Tensor]]] = {}
    ret_scores: Dict[str, Dict[str, int]] = {}
    for collection_path, _, collection in collections:
        tensors: Dict[str, Tuple[torch.Tensor, torch.Tensor]] = {}
        scores: Dict[str, int] = {}
        for m in get_dynamic_emb_module(collection):
            if set_all:
                thresholds = {t: score_threshold for t in m.table_names}
            else:
                thresholds = {t: s for t, s in score_threshold[collection_path].items() if t in m.table_names}
            if not thresholds:
                continue
            t_, s_ = m.incremental_dump(thresholds, pg)
            tensors.update(t_)
            scores.update(s_)
        ret_tensors[collection_path] = tensors
        ret_scores[collection_path] = scores
    return ret_tensors, ret_scores
