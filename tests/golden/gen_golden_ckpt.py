"""Golden-vector generator for the checkpoint wire format: executes the reference's OWN Python (function source exec'd unmodified from
/root/reference) in this container and commits what it produced as `checkpoint.json`.  Not run by the test suite; rerun by hand:
    python tests/golden/gen_golden_ckpt.py

  names      : encode_meta_json_file_path / encode_checkpoint_file_path / encode_counter_checkpoint_file_path
               (corelib/dynamicemb/dynamicemb/batched_dynamicemb_tables.py:73-93)
  loading    : get_loading_files / find_files (:95-210) on a synthetic directory written by 2 ranks, asked for by world sizes 2, 1 and 3
  perms      : score_dump_permutation / score_load_permutation / get_physical_score_order (dynamicemb_config.py:257-304)
  opt        : truncate_optimizer_states_for_checkpoint / pad_optimizer_states_from_checkpoint (optimizer.py:514-560) on a 4-wide run-time state
  reader     : the bytes `_dump_table` writes for a batch (`tensor.cpu().numpy().tobytes()`, key_value_table.py:1166-1181) read back by the
               reference's `_iter_batches_from_files` (:1190-1282), single- and two-word scores, optimizer states, batch size 3 (so that the
               batching is exercised) — the test feeds the same bytes to our reader and our writer's bytes must equal them
"""
import ast
import glob
import json
import os
import sys
import tempfile
from functools import partial

import numpy as np
import torch

REF = "/root/reference/corelib/dynamicemb/dynamicemb"
OUT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(OUT)), "recsys-examples_b200"))


def extract(path, names, g):
    tree = ast.parse(open(path).read())
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in names:
            exec(compile(ast.Module(body=[node], type_ignores=[]), path, "exec"), g)
    return [g[n] for n in names]


class _NoDist:
    @staticmethod
    def is_initialized():
        return False


def main():
    from typing import Any, Dict, Iterator, List, Optional, Tuple
    from torch import Tensor
    import enum

    class DynamicEmbScoreStrategy(enum.IntEnum):      # values of dynamicemb_config.py:113-140
        TIMESTAMP = 0
        STEP = 1
        CUSTOMIZED = 2
        LFU = 3
        NO_EVICTION = 4

    g = {"os": os, "glob": glob, "partial": partial, "np": np, "torch": torch, "dist": _NoDist, "Tensor": Tensor, "Any": Any, "Dict": Dict,
         "Iterator": Iterator, "List": List, "Optional": Optional, "Tuple": Tuple, "KEY_TYPE": torch.int64, "EMBEDDING_TYPE": torch.float32,
         "SCORE_TYPE": torch.int64, "OPT_STATE_TYPE": torch.float32,
         "torch_dtype_to_np_dtype": {torch.uint64: np.uint64, torch.int64: np.int64, torch.float32: np.float32},
         "DynamicEmbScoreStrategy": DynamicEmbScoreStrategy, "ScoreStrategy": Any, "BaseDynamicEmbeddingOptimizer": Any}
    enc_meta, enc_ckpt, enc_cnt, find_files, get_loading_files = extract(
        f"{REF}/batched_dynamicemb_tables.py",
        ["encode_meta_json_file_path", "encode_checkpoint_file_path", "encode_counter_checkpoint_file_path", "find_files", "get_loading_files"], g)
    phys, dump_perm, load_perm = extract(f"{REF}/dynamicemb_config.py", ["get_physical_score_order", "score_dump_permutation", "score_load_permutation"], g)
    (iter_batches,) = extract(f"{REF}/key_value_table.py", ["_iter_batches_from_files"], g)
    trunc, pad = extract(f"{REF}/optimizer.py", ["truncate_optimizer_states_for_checkpoint", "pad_optimizer_states_from_checkpoint"], g)

    rec = {}
    rec["names"] = {
        "meta": enc_meta("ROOT", "t_user"),
        "ckpt": {item: enc_ckpt("ROOT", "t_user", 3, 8, item) for item in ("keys", "values", "scores", "opt_values")},
        "counter": {item: enc_cnt("ROOT", "t_user", 1, 2, item) for item in ("keys", "frequencies")},
    }

    # ---- discovery
    with tempfile.TemporaryDirectory() as d:
        for r in range(2):
            for item in ("keys", "values", "scores", "opt_values"):
                open(enc_ckpt(d, "tab", r, 2, item), "wb").close()
            for item in ("keys", "values"):                       # a second table without scores / optimizer files
                open(enc_ckpt(d, "bare", r, 2, item), "wb").close()
        rel = lambda lists: [[os.path.relpath(p, d) for p in l] for l in lists]      # noqa: E731
        rec["loading"] = {
            "tab_rank1_world2": rel(get_loading_files(d, "tab", rank=1, world_size=2)),
            "tab_rank0_world1": rel(get_loading_files(d, "tab", rank=0, world_size=1)),
            "tab_rank2_world3": rel(get_loading_files(d, "tab", rank=2, world_size=3)),
            "bare_rank0_world2": rel(get_loading_files(d, "bare", rank=0, world_size=2)),
            "missing_rank0_world1": rel(get_loading_files(d, "nothing", rank=0, world_size=1)),
        }
        os.remove(enc_ckpt(d, "tab", 1, 2, "keys"))
        try:
            get_loading_files(d, "tab", rank=0, world_size=2)
            rec["loading"]["corrupt_error"] = None
        except RuntimeError as e:
            rec["loading"]["corrupt_error"] = str(e).replace(d, "ROOT")

    # ---- score column order
    S = DynamicEmbScoreStrategy
    rec["perms"] = {}
    for name, st in [("timestamp", S.TIMESTAMP), ("lfu", S.LFU), ("ts_lfu", (S.TIMESTAMP, S.LFU)), ("lfu_ts", (S.LFU, S.TIMESTAMP))]:
        rec["perms"][name] = {"physical": [int(x) for x in phys(st)], "dump": dump_perm(st), "load": load_perm(st)}

    # ---- optimizer state width
    class _Opt:                                                       # row-wise Adagrad: 4 floats at run time, 1 in the file
        def get_ckpt_state_dim(self, emb_dim):
            return 1

        def get_state_dim(self, emb_dim):
            return 4
    rt = torch.arange(12, dtype=torch.float32).view(3, 4)
    t = trunc(_Opt(), 8, rt)
    p = pad(_Opt(), 8, t, 0.5, torch.float32, torch.device("cpu"))
    rec["opt"] = {"runtime": rt.tolist(), "truncated": t.tolist(), "padded_init_0.5": p.tolist()}

    # ---- reader
    rng = np.random.default_rng(7)
    n, dim, sdim = 8, 4, 4
    keys = rng.integers(-(1 << 62), 1 << 62, size=n, dtype=np.int64)
    emb = rng.standard_normal((n, dim)).astype(np.float32)
    sc1 = rng.integers(0, 1 << 40, size=n, dtype=np.int64)
    sc2 = rng.integers(0, 1 << 40, size=(n, 2), dtype=np.int64)
    opt = rng.standard_normal((n, sdim)).astype(np.float32)
    rec["reader"] = {"keys": keys.tolist(), "emb": emb.tolist(), "scores1": sc1.tolist(), "scores2": sc2.tolist(), "opt": opt.tolist(),
                     "bytes": {k: torch.from_numpy(v).cpu().numpy().tobytes().hex() for k, v in
                               (("keys", keys), ("emb", emb), ("scores1", sc1), ("scores2", sc2), ("opt", opt))}}
    with tempfile.TemporaryDirectory() as d:
        paths = {k: os.path.join(d, k) for k in ("keys", "emb", "scores1", "scores2", "opt")}
        for k, v in (("keys", keys), ("emb", emb), ("scores1", sc1), ("scores2", sc2), ("opt", opt)):
            with open(paths[k], "wb") as f:
                f.write(torch.from_numpy(v).cpu().numpy().tobytes())
        for tag, spath, ns in (("read1", "scores1", 1), ("read2", "scores2", 2)):
            got = list(iter_batches(paths["keys"], paths["emb"], paths[spath], paths["opt"], dim, sdim, torch.device("cpu"), batch_size=3, num_scores=ns))
            rec["reader"][tag] = [[x.tolist() if x is not None else None for x in b] for b in got]
        got = list(iter_batches(paths["keys"], paths["emb"], None, None, dim, 0, torch.device("cpu"), batch_size=5))
        rec["reader"]["read_no_scores_no_opt"] = [[x.tolist() if x is not None else None for x in b] for b in got]
        # cross-check at generation time: files written by OUR writer are read identically by the reference's reader
        from dynamicemb import checkpoint as ck
        ours = {k: os.path.join(d, "ours_" + k) for k in ("keys", "emb", "scores2", "opt")}
        with ck.TableFileWriter(ours["keys"], ours["emb"], ours["scores2"], ours["opt"]) as w:
            for a in range(0, n, 5):
                w.write(torch.from_numpy(keys[a:a + 5]), torch.from_numpy(emb[a:a + 5]), torch.from_numpy(sc2[a:a + 5]), torch.from_numpy(opt[a:a + 5]))
        got = list(iter_batches(ours["keys"], ours["emb"], ours["scores2"], ours["opt"], dim, sdim, torch.device("cpu"), batch_size=3, num_scores=2))
        assert [[x.tolist() for x in b] for b in got] == rec["reader"]["read2"], "reference reader disagrees on files written by our writer"

    with open(os.path.join(OUT, "checkpoint.json"), "w") as f:
        json.dump(rec, f, indent=1)
    print("checkpoint.json written;", len(json.dumps(rec)), "bytes")


if __name__ == "__main__":
    main()
