"""Model-level checkpoint / score API on the CPU shim: DynamicEmbDump / DynamicEmbLoad / set_score / get_score / incremental_dump walk a
model, find its dynamic-embedding collections and use the directory layout and dictionary nesting of the reference
(corelib/dynamicemb/dynamicemb/dump_load.py:103-287, incremental_dump.py:47-348)."""
import os

import pytest
import torch
from torch import nn

from tests.cpu_ext_shim import patched_module
from tests.test_admission_cpu import _module


class _Collection(nn.Module):
    """Stand-in for a sharded embedding collection (the role TorchRec's ShardedEmbeddingCollection plays in the reference)."""
    _is_dynamicemb_collection = True

    def __init__(self, emb):
        super().__init__()
        self._emb_module = emb          # the private attribute the reference's walker follows (dump_load.py:73)


class _Model(nn.Module):
    def __init__(self, a, b):
        super().__init__()
        self.sparse = nn.Module()
        self.sparse.ec = _Collection(a)
        self.ebc = _Collection(b)
        self.dense = nn.Linear(4, 4)


def _train(m, ids):
    m.train()
    x = torch.tensor(ids, dtype=torch.int64)
    out = m(x, torch.arange(0, len(ids) + 1, dtype=torch.int64))
    out.backward(torch.ones_like(out))


def test_model_level_dump_load_scores_cpu_shim(tmp_path):
    from dynamicemb import DynamicEmbDump, DynamicEmbLoad, DynamicEmbScoreStrategy as S
    from dynamicemb.incremental_dump import get_score, incremental_dump, set_score
    from dynamicemb.dump_load import find_sharded_modules, get_dynamic_emb_module
    with patched_module():
        mk = lambda strat: _module({"fused_prefetch": False}, None, T=2, score_strategy=strat)          # noqa: E731
        a, b = mk(S.STEP), mk(S.CUSTOMIZED)
        model = _Model(a, b)
        found = find_sharded_modules(model, "")
        assert sorted(p for p, _, _ in found) == ["model.ebc", "model.sparse.ec"] and sorted(n for _, n, _ in found) == ["ebc", "ec"]
        assert get_dynamic_emb_module(model) == [a, b] or get_dynamic_emb_module(model) == [b, a]
        # scores: by collection NAME for set_score, by collection PATH in what get_score returns
        with pytest.raises(RuntimeError):
            _train(b, [1, 2])                                      # CUSTOMIZED tables need a score first
        b.reset_prefetch()
        set_score(model, {"ebc": {"t0": 7, "t1": 9}})
        assert get_score(model) == {"model.sparse.ec": {"t0": 1, "t1": 1}, "model.ebc": {"t0": 7, "t1": 9}}
        with pytest.warns(UserWarning):
            set_score(model, {"nope": {"t0": 1}})
        with pytest.raises(ValueError):
            set_score(model, {"ebc": {"t0": "x"}})
        _train(a, [1, 2, 3, 4])                                    # table t0 gets 1, 2 (first half of the feature-major batch), t1 gets 3, 4
        _train(a, [2, 5, 4, 6])
        _train(b, [10, 11, 12, 13])
        # incremental dump: one threshold for all tables / selected tables of one collection
        tensors, nxt = incremental_dump(model, 2)
        assert sorted(tensors["model.sparse.ec"]["t0"][0].tolist()) == [2, 5] and sorted(tensors["model.sparse.ec"]["t1"][0].tolist()) == [4, 6]
        assert sorted(tensors["model.ebc"]["t0"][0].tolist()) == [10, 11] and nxt["model.sparse.ec"] == {"t0": 3, "t1": 3}
        tensors, nxt = incremental_dump(model, {"model.ebc": {"t1": 9}})
        assert list(tensors) == ["model.ebc"] and list(tensors["model.ebc"]) == ["t1"] and sorted(tensors["model.ebc"]["t1"][0].tolist()) == [12, 13]
        with pytest.warns(UserWarning):
            assert incremental_dump(nn.Linear(2, 2), 1) is None
        # dump -> directory per collection path, files per table; refuses a non-empty directory; selected tables only
        d = str(tmp_path / "ckpt")
        DynamicEmbDump(d, model, optim=True)
        assert sorted(os.listdir(d)) == ["model.ebc", "model.sparse.ec"]
        assert "t0_emb_keys.rank_0.world_size_1" in os.listdir(os.path.join(d, "model.sparse.ec"))
        with pytest.raises(Exception, match="already contains files"):
            DynamicEmbDump(d, model)
        d2 = str(tmp_path / "only_t1")
        DynamicEmbDump(d2, model, table_names={"model.ebc": ["t1"]})
        assert all(f.startswith("t1_") for f in os.listdir(os.path.join(d2, "model.ebc"))) and os.listdir(os.path.join(d2, "model.ebc"))
        # load into a fresh model: identical tables
        a2, b2 = mk(S.STEP), mk(S.CUSTOMIZED)
        model2 = _Model(a2, b2)
        DynamicEmbLoad(d, model2, optim=True)
        for src, dst in ((a, a2), (b, b2)):
            for t in range(2):
                k1, v1 = src.export_keys_values(t)
                k2, v2 = dst.export_keys_values(t)
                i1, i2 = torch.argsort(k1), torch.argsort(k2)
                assert torch.equal(k1[i1], k2[i2]) and torch.equal(v1[i1], v2[i2])
        assert a2.get_score() == {"t0": 3, "t1": 3}                                 # step_score of the meta file
        with pytest.raises(AssertionError):
            set_score(model2, 5)                                                    # one integer = every table; STEP tables refuse a manual score
        set_score(model2, {"ebc": {"t0": 5, "t1": 5}})
        assert get_score(model2)["model.ebc"] == {"t0": 5, "t1": 5}
        with pytest.raises(Exception, match="can't find path"):
            DynamicEmbLoad(str(tmp_path / "missing"), model2)
        with pytest.warns(UserWarning):
            DynamicEmbDump(str(tmp_path / "empty"), nn.Linear(2, 2))
