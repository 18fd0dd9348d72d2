"""Checkpoint wire format of DynamicEmb tables: files a table trained here loads into the reference and back.

Restates the file layout of /root/reference/corelib/dynamicemb/dynamicemb/batched_dynamicemb_tables.py:73-210 (file names, discovery,
world-size re-sharding rule) and key_value_table.py:1134-1520 (`_dump_table`, `_iter_batches_from_files`, `_validate_load_meta`,
`_load_key_values`).  Per table and per rank a dump writes raw little-endian arrays, no header:

    <save_dir>/<table>_opt_args.json                                    optimizer arguments + "evict_strategy" + "dist_type" (+ "step_score")
    <save_dir>/<table>_emb_keys.rank_<r>.world_size_<W>                  int64  [n]
    <save_dir>/<table>_emb_values.rank_<r>.world_size_<W>                float32 [n, dim]
    <save_dir>/<table>_emb_scores.rank_<r>.world_size_<W>                int64  [n] or [n, num_scores] in the user's (logical) score order;
                                                                          LRU tables store `dump timestamp - score` (age), so a load at a later
                                                                          time restores `load timestamp - age`
    <save_dir>/<table>_emb_opt_values.rank_<r>.world_size_<W>            float32 [n, ckpt_state_dim]   (row-wise Adagrad: 1 column)

The functions below move bytes between those files and device tensors; the device work (table scan, insert, row copies) is done by the
module through the native ops.  Nothing here computes on the CPU beyond numpy (de)serialisation.
"""
import glob
import json
import os
from typing import Iterator, List, Optional, Tuple

import numpy as np
import torch

from .types import DynamicEmbScoreStrategy

KEY_TYPE = torch.int64          # reference types.py:108-111
EMBEDDING_TYPE = torch.float32
SCORE_TYPE = torch.int64
OPT_STATE_TYPE = torch.float32

_ITEMS = ("keys", "values", "scores", "opt_values")
_COUNTER_ITEMS = ("keys", "frequencies")


def encode_meta_json_file_path(root_path: str, table_name: str) -> str:
    return os.path.join(root_path, f"{table_name}_opt_args.json")


def encode_checkpoint_file_path(root_path: str, table_name: str, rank, world_size, item: str) -> str:
    assert item in _ITEMS
    return os.path.join(root_path, f"{table_name}_emb_{item}.rank_{rank}.world_size_{world_size}")


def encode_counter_checkpoint_file_path(root_path: str, table_name: str, rank, world_size, item: str) -> str:
    assert item in _COUNTER_ITEMS
    return os.path.join(root_path, f"{table_name}_counter_{item}.rank_{rank}.world_size_{world_size}")


_SUFFIX = {
    "emb_keys": (encode_checkpoint_file_path, "keys"), "emb_values": (encode_checkpoint_file_path, "values"),
    "emb_scores": (encode_checkpoint_file_path, "scores"), "opt_values": (encode_checkpoint_file_path, "opt_values"),
    "counter_keys": (encode_counter_checkpoint_file_path, "keys"), "counter_frequencies": (encode_counter_checkpoint_file_path, "frequencies"),
}


def find_files(root_path: str, table_name: str, suffix: str) -> Tuple[List[str], int]:
    """All rank files of one item; the world size is read from the file name and every rank must be present."""
    if suffix not in _SUFFIX:
        raise RuntimeError(f"Invalid suffix: {suffix}")
    enc, item = _SUFFIX[suffix]
    files = sorted(glob.glob(enc(root_path, table_name, "*", "*", item)))
    if not files:
        return [], 0
    world_size = int(files[0].rsplit(".", 1)[-1].rsplit("_", 1)[-1])
    if len(files) != world_size:
        raise RuntimeError(f"Checkpoints is corrupted. Found {len(files)} under path {root_path} for table {table_name}, "
                           f"but the number of checkpointed world size is {world_size}.")
    have = set(files)
    for r in range(world_size):
        want = enc(root_path, table_name, r, world_size, item)
        if want not in have:
            raise RuntimeError(f"Checkpoints is corrupted. Expected file path {want} for table {table_name}, but it is not found.")
    return files, len(files)


def get_loading_files(root_path: str, name: str, rank: int, world_size: int):
    """(key, value, score, opt, counter key, counter frequency) file lists this rank has to read: its own files when the checkpoint was
    written by the same world size, otherwise ALL files (the reader then keeps `key % world_size == rank`)."""
    if not os.path.exists(root_path):
        raise RuntimeError(f"can't find path to load, path: {root_path}")
    key_files, nk = find_files(root_path, name, "emb_keys")
    value_files, nv = find_files(root_path, name, "emb_values")
    score_files, nsf = find_files(root_path, name, "emb_scores")
    opt_files, no = find_files(root_path, name, "opt_values")
    if nk != nv:
        assert nk > 0, f"No key files found under path {root_path} for table {name}"
        raise RuntimeError(f"The number of key files under path {root_path} for table {name} does not match the number of value files.")
    ck_files, nck = find_files(root_path, name, "counter_keys")
    cf_files, ncf = find_files(root_path, name, "counter_frequencies")
    if nck != ncf:
        raise RuntimeError(f"The number of key files of admission counter under path {root_path} for table {name} does not match the "
                           f"number of frequency files({nck}/{ncf}).")
    if nck > 0 and nck != nk:
        raise RuntimeError(f"The number of key files under path {root_path} for table {name} does not match the number of keys files of "
                           f"admission counter({nk}/{nck}).")
    if world_size == nk:
        own = lambda item: [encode_checkpoint_file_path(root_path, name, rank, world_size, item)]              # noqa: E731
        cown = lambda item: [encode_counter_checkpoint_file_path(root_path, name, rank, world_size, item)]     # noqa: E731
        return (own("keys"), own("values"), own("scores") if nsf == nk else [], own("opt_values") if no == nk else [],
                cown("keys") if nck == nk else [], cown("frequencies") if ncf == nk else [])
    return key_files, value_files, score_files, opt_files, ck_files, cf_files


def save_to_json(data: dict, file_path: str) -> None:
    with open(file_path, "w") as f:
        json.dump(data, f, indent=4)


def load_from_json(file_path: str) -> dict:
    with open(file_path, "r") as f:
        return json.load(f)


# ---------------------------------------------------------------------------------------------------------------- score column order
def get_physical_score_order(score_strategy) -> tuple:
    """Device order of the score words: the compound {TIMESTAMP, LFU} is always (timestamp, frequency) whatever order the user wrote
    (dynamicemb_config.py:257-274)."""
    if isinstance(score_strategy, tuple):
        if frozenset(score_strategy) == frozenset({DynamicEmbScoreStrategy.TIMESTAMP, DynamicEmbScoreStrategy.LFU}):
            return (DynamicEmbScoreStrategy.TIMESTAMP, DynamicEmbScoreStrategy.LFU)
        raise NotImplementedError(f"No physical score layout for compound score_strategy {score_strategy}.")
    return (score_strategy,)


def score_dump_permutation(score_strategy) -> List[int]:
    """file[:, j] = device[:, perm[j]]."""
    if not isinstance(score_strategy, tuple):
        return [0]
    physical = get_physical_score_order(score_strategy)
    return [physical.index(s) for s in score_strategy]


def score_load_permutation(score_strategy) -> List[int]:
    """device[:, p] = file[:, perm[p]] — inverse of score_dump_permutation."""
    if not isinstance(score_strategy, tuple):
        return [0]
    physical = get_physical_score_order(score_strategy)
    logical = list(score_strategy)
    return [logical.index(s) for s in physical]


# ---------------------------------------------------------------------------------------------------------------- optimizer state width
def truncate_optimizer_states_for_checkpoint(optimizer, emb_dim: int, opt_states_runtime: torch.Tensor) -> torch.Tensor:
    ckpt_dim = optimizer.get_ckpt_state_dim(emb_dim)
    n = opt_states_runtime.size(1)
    if ckpt_dim == 0 or n == ckpt_dim:
        return opt_states_runtime
    if n < ckpt_dim:
        raise ValueError(f"Runtime optimizer state width {n} is less than checkpoint width {ckpt_dim}.")
    return opt_states_runtime[:, :ckpt_dim].contiguous()


def pad_optimizer_states_from_checkpoint(optimizer, emb_dim: int, opt_states_from_file: torch.Tensor, initial_accumulator_value: float,
                                         values_dtype: torch.dtype, device) -> torch.Tensor:
    runtime_dim = optimizer.get_state_dim(emb_dim)
    file_dim = opt_states_from_file.size(1)
    if runtime_dim == 0:
        return opt_states_from_file
    if file_dim >= runtime_dim:
        return opt_states_from_file[:, :runtime_dim].contiguous().to(values_dtype)
    out = torch.full((opt_states_from_file.size(0), runtime_dim), initial_accumulator_value, dtype=values_dtype, device=device)
    out[:, :file_dim] = opt_states_from_file.to(values_dtype)
    return out


# ---------------------------------------------------------------------------------------------------------------- key -> owning rank
def _lsr(x: torch.Tensor, k: int) -> torch.Tensor:
    """Logical right shift of the 64-bit patterns held in an int64 tensor."""
    return (x >> k) & ((1 << (64 - k)) - 1)


def _i64const(u: int) -> int:
    return u - (1 << 64) if u >= (1 << 63) else u


def fmix64_tensor(keys: torch.Tensor) -> torch.Tensor:
    """murmur3 fmix64 on the bit patterns of an int64 tensor (wrapping multiplies) — same bits as csrc/demb_common.cuh fmix64."""
    k = keys.to(torch.int64)
    k = k ^ _lsr(k, 33)
    k = k * _i64const(0xFF51AFD7ED558CCD)
    k = k ^ _lsr(k, 33)
    k = k * _i64const(0xC4CEB9FE1A85EC53)
    return k ^ _lsr(k, 33)


def _u64_mod(x: torch.Tensor, m: int) -> torch.Tensor:
    """(unsigned 64-bit x) % m for bit patterns in an int64 tensor."""
    return ((_lsr(x, 1) % m) * 2 + (x & 1)) % m


def owner_rank(keys: torch.Tensor, world_size: int, dist_type: str) -> torch.Tensor:
    """Rank that owns each key under row-wise sharding (sparse_block_bucketize_features.cu:30-37,254-259; csrc/demb_dist.cu)."""
    if dist_type == "roundrobin":
        return _u64_mod(keys.to(torch.int64), world_size)
    if dist_type == "hash_roundrobin":
        return _u64_mod(fmix64_tensor(keys), world_size)
    raise NotImplementedError("re-sharding a 'continuous' checkpoint needs the table's hash_size; load it with the world size it was dumped with")


# ---------------------------------------------------------------------------------------------------------------- file <-> tensors
_NP = {torch.int64: np.int64, torch.float32: np.float32}


class TableFileWriter:
    """Appends (keys, embeddings, scores, optimizer states) batches to the four files of one table and rank."""

    def __init__(self, key_path: str, value_path: str, score_path: str, opt_path: Optional[str], append: bool = False):
        mode = "ab" if append else "wb"
        self._f = [open(key_path, mode), open(value_path, mode), open(score_path, mode), open(opt_path, mode) if opt_path else None]

    def write(self, keys: torch.Tensor, embeddings: torch.Tensor, scores: torch.Tensor, opt_states: Optional[torch.Tensor]) -> None:
        for f, t, dt in zip(self._f, (keys, embeddings, scores, opt_states), (KEY_TYPE, EMBEDDING_TYPE, SCORE_TYPE, OPT_STATE_TYPE)):
            if f is not None and t is not None:
                f.write(t.detach().to(dt).contiguous().cpu().numpy().tobytes())

    def close(self) -> None:
        for f in self._f:
            if f is not None:
                f.close()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


def iter_batches_from_files(key_path: str, value_path: str, score_path: Optional[str], opt_path: Optional[str], dim: int, optstate_dim: int,
                            device, batch_size: int = 65536, num_scores: int = 1, rank: int = 0, world_size: int = 1,
                            dist_type: str = "roundrobin") -> Iterator[Tuple[torch.Tensor, torch.Tensor, Optional[torch.Tensor], Optional[torch.Tensor]]]:
    """(keys, embeddings, scores, opt_states) batches of one file set on `device`.  With world_size > 1 (a checkpoint written by a
    different world size: every rank reads all files) only the keys this rank owns are kept.  The reference keeps
    `key % world_size == rank` (key_value_table.py:1263-1270), which is the owner under `roundrobin`; here the owner follows the table's
    dist_type (`owner_rank`), so a re-sharded `hash_roundrobin` table finds its keys again."""
    num_keys = os.path.getsize(key_path) // 8
    fs = [open(key_path, "rb"), open(value_path, "rb"),
          open(score_path, "rb") if score_path and os.path.exists(score_path) else None, open(opt_path, "rb") if opt_path else None]
    widths = (1, dim, num_scores, optstate_dim)
    dts = (KEY_TYPE, EMBEDDING_TYPE, SCORE_TYPE, OPT_STATE_TYPE)
    try:
        for start in range(0, num_keys, batch_size):
            n = min(num_keys - start, batch_size)
            out = []
            for f, w, dt in zip(fs, widths, dts):
                if f is None:
                    out.append(None)
                    continue
                a = np.frombuffer(f.read(n * w * (8 if dt == torch.int64 else 4)), dtype=_NP[dt])
                if a.size != n * w:
                    raise ValueError(f"{f.name}: short read ({a.size} of {n * w} elements)")
                t = torch.from_numpy(a.copy()).to(device)
                out.append(t.view(n, w) if w != 1 or dt != KEY_TYPE else t)
            keys, emb, scores, opt = out
            if scores is not None and num_scores == 1:
                scores = scores.view(-1)
            if world_size > 1:
                keep = owner_rank(keys, world_size, dist_type) == rank
                keys, emb = keys[keep], emb[keep]
                scores = scores[keep] if scores is not None else None
                opt = opt[keep] if opt is not None else None
            yield keys, emb, scores, opt
    finally:
        for f in fs:
            if f is not None:
                f.close()


def validate_load_meta(meta: dict, optimizer, evict_strategy_str: str, dist_type: str, dim: int, num_scores: int, key_path: str,
                       value_path: str, score_path: Optional[str], opt_path: Optional[str], include_optim: bool):
    """The checks of `_validate_load_meta` (key_value_table.py:1294-1400).  Returns (include_optim, file_optstate_dim, num_keys)."""
    opt_type = meta.get("opt_type", None)
    if opt_type and optimizer.get_opt_args().get("opt_type", None) != opt_type:
        include_optim = False
        print(f"Optimizer type mismatch: {opt_type} != {optimizer.get_opt_args().get('opt_type')}. Will not load optimizer states.")
    ck_evict = meta.get("evict_strategy", None)
    if ck_evict and evict_strategy_str != ck_evict:
        raise ValueError(f"Evict strategy mismatch: {ck_evict} != {evict_strategy_str}")
    ck_dist = meta.get("dist_type", "roundrobin")
    if dist_type != ck_dist:
        raise ValueError(f"Input dist_type mismatch: checkpoint was dumped with {ck_dist!r}, but runtime table is configured with "
                         f"{dist_type!r}. Please load with a matching dist_type.")
    if not opt_path or not os.path.exists(opt_path):
        include_optim = False
    if optimizer.get_state_dim(dim) == 0:
        include_optim = False
    if include_optim:
        optimizer.set_opt_args(meta)
    num_keys = os.path.getsize(key_path) // 8
    if num_keys != os.path.getsize(value_path) // 4 // dim:
        raise ValueError(f"The number of keys in {key_path} does not match with number of embeddings in {value_path}.")
    if score_path and os.path.exists(score_path) and os.path.getsize(score_path) // 8 != num_keys * num_scores:
        raise ValueError(f"The number of keys in {key_path} does not match with number of scores in {score_path}.")
    file_optstate_dim = 0
    if include_optim:
        nbytes = os.path.getsize(opt_path)
        if num_keys == 0:
            if nbytes != 0:
                raise ValueError(f"Optimizer state file {opt_path} is non-empty but key file has no keys.")
        else:
            if nbytes % (num_keys * 4) != 0:
                raise ValueError(f"Optimizer state file {opt_path} size {nbytes} is not divisible by {num_keys * 4} (num_keys={num_keys}).")
            file_optstate_dim = nbytes // (num_keys * 4)
            want = optimizer.get_ckpt_state_dim(dim)
            if file_optstate_dim != want:
                raise ValueError(f"Optimizer state width in checkpoint is {file_optstate_dim}; expected {want}.")
    return include_optim, file_optstate_dim, num_keys


def all_gather_keys_values(keys: torch.Tensor, values: torch.Tensor, pg) -> Tuple[torch.Tensor, torch.Tensor]:
    """Every rank's (keys [n_r], values [n_r, D]) concatenated in rank order, on the CPU (key_value_table.py:75-111: counts first, then
    the rows padded to the largest count — all_gather needs equal shapes)."""
    import torch.distributed as dist
    world = dist.get_world_size(group=pg)
    n = keys.numel()
    count = torch.tensor([n], dtype=torch.long, device=keys.device)
    counts = [torch.empty_like(count) for _ in range(world)]
    dist.all_gather(counts, count, group=pg)
    counts = [int(c.item()) for c in counts]
    max_n = max(counts)
    kpad = torch.zeros(max_n, dtype=torch.int64, device=keys.device)
    vpad = torch.zeros(max_n, values.shape[1], dtype=values.dtype, device=values.device)
    if n > 0:
        kpad[:n] = keys.view(torch.int64) if keys.dtype != torch.int64 else keys
        vpad[:n] = values
    gk = [torch.empty_like(kpad) for _ in range(world)]
    gv = [torch.empty_like(vpad) for _ in range(world)]
    dist.all_gather(gk, kpad, group=pg)
    dist.all_gather(gv, vpad, group=pg)
    return (torch.cat([gk[i][:counts[i]] for i in range(world)]).cpu(), torch.cat([gv[i][:counts[i]] for i in range(world)]).cpu())
