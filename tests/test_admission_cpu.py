"""Admission host logic on the CPU (admission part).

The decision sequence (`admission_split`), the strategy / counter classes and the module's op-by-op prefetch with admission are
exercised without a GPU: `tests/cpu_ext_shim.py` restates the native ops the host logic calls on the oracle, so what runs here is the
product's Python above the op layer.  The GPU test of the same scenario (tests/test_zz_admission_gpu.py) runs it on the kernels.

Model of the expected behaviour = the reference's own test (corelib/dynamicemb/test/unit_tests/test_embedding_admission.py:165-220):
count, per key, how often it was presented while not in the table; only keys whose count reached the threshold may be stored.
"""
import numpy as np
import pytest
import torch

from tests.cpu_ext_shim import patched_module


class DictCounter:
    """Dictionary stand-in for MultiTableKVCounter (the `Counter` interface)."""

    def __init__(self):
        self.d = {}
        self.erased = []

    def add(self, keys, table_ids, frequencies):
        out = []
        for k, t, f in zip(keys.tolist(), table_ids.tolist(), frequencies.tolist()):
            self.d[(t, k)] = self.d.get((t, k), 0) + f
            out.append(self.d[(t, k)])
        return torch.tensor(out, dtype=torch.int64)

    def erase(self, keys, table_ids):
        for k, t in zip(keys.tolist(), table_ids.tolist()):
            self.erased.append((t, k))
            self.d.pop((t, k), None)


def test_admission_split_counts_and_forgets():
    from dynamicemb.embedding_admission import FrequencyAdmissionStrategy, admission_split
    strat, ctr = FrequencyAdmissionStrategy(threshold=3), DictCounter()
    keys = torch.tensor([10, 11, 12], dtype=torch.int64)
    tids = torch.tensor([0, 0, 1], dtype=torch.int64)
    m1, f1 = admission_split(keys, tids, None, strat, ctr)                       # seen once each
    assert m1.tolist() == [False, False, False] and f1.tolist() == [1, 1, 1]
    m2, f2 = admission_split(keys, tids, torch.tensor([2, 1, 5]), strat, ctr)    # batch frequencies (LFU-style)
    assert f2.tolist() == [3, 2, 6] and m2.tolist() == [True, False, True]
    assert sorted(ctr.erased) == [(0, 10), (1, 12)] and ctr.d == {(0, 11): 2}    # admitted keys leave the counter
    m3, _ = admission_split(keys[1:2], tids[1:2], None, strat, ctr)
    assert m3.tolist() == [True]
    e = torch.empty(0, dtype=torch.int64)
    m0, f0 = admission_split(e, e, None, strat, ctr)
    assert m0.numel() == 0 and f0.numel() == 0


def test_strategy_validation_and_non_admitted_initializer():
    from dynamicemb import DynamicEmbInitializerArgs, DynamicEmbInitializerMode, FrequencyAdmissionStrategy
    with pytest.raises(ValueError):
        FrequencyAdmissionStrategy(threshold=-1)
    s = FrequencyAdmissionStrategy(threshold=0)
    with pytest.raises(ValueError):
        s.admit(torch.zeros(3, dtype=torch.int64), torch.zeros(2, dtype=torch.int64))
    assert s.admit(torch.zeros(2, dtype=torch.int64), torch.zeros(2, dtype=torch.int64)).all()     # threshold 0 admits everything
    buf = torch.ones(4, 8)
    assert s.initialize_non_admitted_embeddings(buf, torch.tensor([1, 3])) is False and bool((buf == 1).all())
    s2 = FrequencyAdmissionStrategy(2, DynamicEmbInitializerArgs(mode=DynamicEmbInitializerMode.CONSTANT, value=0.5))
    assert s2.initialize_non_admitted_embeddings(buf, torch.tensor([1, 3])) is True
    assert bool((buf[[1, 3]] == 0.5).all()) and bool((buf[[0, 2]] == 1).all())


def _module(btm_kwargs, threshold, T=1, pooling=None, score_strategy=None, cap=2048, counter_cap=4096, dim=32, caching=False, local_hbm=0):
    from dynamicemb import (BatchedDynamicEmbeddingTablesV2, DynamicEmbInitializerArgs, DynamicEmbInitializerMode, DynamicEmbPoolingMode,
                            DynamicEmbScoreStrategy, DynamicEmbTableOptions, EmbOptimType, FrequencyAdmissionStrategy, KVCounter)
    strat = FrequencyAdmissionStrategy(threshold=threshold) if threshold is not None else None
    opts = [DynamicEmbTableOptions(dim=dim, max_capacity=cap, bucket_capacity=128,
                                   initializer_args=DynamicEmbInitializerArgs(mode=DynamicEmbInitializerMode.DEBUG),
                                   score_strategy=score_strategy if score_strategy is not None else DynamicEmbScoreStrategy.STEP,
                                   caching=caching, local_hbm_for_values=local_hbm,
                                   admit_strategy=strat, admission_counter=KVCounter(counter_cap, bucket_capacity=128) if strat is not None else None)
            for _ in range(T)]
    return BatchedDynamicEmbeddingTablesV2(opts, table_names=[f"t{i}" for i in range(T)], feature_table_map=list(range(T)),
                                           pooling_mode=pooling if pooling is not None else DynamicEmbPoolingMode.NONE,
                                           optimizer=EmbOptimType.SGD, learning_rate=0.5, **btm_kwargs)


def _stored_keys(m, t=0):
    keys, _ = m.export_keys_values(t)
    return set(keys.tolist())


def _snapshot(m, t=0):
    keys, rows = m.export_keys_values(t)
    return {k: rows[i].clone() for i, k in enumerate(keys.tolist())}


def scenario_sequence(dev):
    """Keys enter the table exactly when their presentation count reaches the threshold; until then their ids read the initializer's
    row (DEBUG: key % 100000) without being stored, and their gradients are dropped; admitted rows train as usual."""
    rng = np.random.default_rng(0)
    thr, D = 3, 32          # 32 / 64 / 128 are the embedding widths the GPU suite validates
    if True:
        m = _module({"device": dev}, thr, dim=D)
        m.train()
        seen, stored_expect = {}, set()
        for step in range(8):
            ids = rng.integers(1, 60, size=50).astype(np.int64)
            x = torch.from_numpy(ids).to(dev)
            offsets = torch.arange(0, ids.size + 1, dtype=torch.int64, device=dev)
            before = _snapshot(m)
            out = m(x, offsets)
            # model: unique missing keys are counted once per step; a key whose count reaches thr is inserted (and forgotten by the counter)
            for k in dict.fromkeys(ids.tolist()):
                if k in stored_expect:
                    continue
                seen[k] = seen.get(k, 0) + 1
                if seen[k] >= thr:
                    stored_expect.add(k)
                    del seen[k]
            # forward values: a stored key reads its row BEFORE this step's update, everything else reads the DEBUG initializer
            for i, k in enumerate(ids.tolist()):
                want = before[k][:D] if k in before else torch.full((D,), float(k % 100000), device=dev)
                assert torch.equal(out[i], want), (step, k)
            out.backward(torch.ones_like(out))
            assert _stored_keys(m) == stored_expect, step
            # SGD lr 0.5, gradient = occurrences of the key: rows of keys stored during or before this step moved by -0.5 * count
            now = _snapshot(m)
            cnt = {k: int((ids == k).sum()) for k in set(ids.tolist())}
            for k, v in now.items():
                base = before[k][:D] if k in before else torch.full((D,), float(k % 100000), device=dev)
                assert torch.allclose(v[:D], base - 0.5 * cnt.get(k, 0)), (step, k)
        assert stored_expect and seen, "the scenario must contain admitted and still-waiting keys"
        # the counter holds exactly the waiting keys with their counts
        ck, cs, _ = next(m._admission_counter.table_.export(0))
        assert dict(zip(ck.tolist(), cs.tolist())) == seen


def scenario_pooled_two_tables(dev, mean):
    """Pooled output with non-admitted ids = pooling of (stored row | initializer row) per id, two tables with separate counters."""
    from dynamicemb import DynamicEmbPoolingMode
    rng = np.random.default_rng(1)
    thr, D, B, T = 2, 32, 6, 2
    if True:
        m = _module({"device": dev}, thr, T=T, pooling=DynamicEmbPoolingMode.MEAN if mean else DynamicEmbPoolingMode.SUM, dim=D)
        m.train()
        stored = [set(), set()]
        seen = [{}, {}]
        for step in range(5):
            lens = rng.integers(0, 4, size=T * B)
            ids = rng.integers(1, 25, size=int(lens.sum())).astype(np.int64)
            offsets = torch.from_numpy(np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)).to(dev)
            out = m(torch.from_numpy(ids).to(dev), offsets)
            want = torch.zeros(B, T * D)
            pos = 0
            for f in range(T):
                for b in range(B):
                    n = int(lens[f * B + b])
                    for k in ids[pos:pos + n].tolist():
                        want[b, f * D:(f + 1) * D] += float(k % 100000) / (n if mean else 1)     # no update has run: every row is its init value
                    pos += n
            assert torch.allclose(out.cpu(), want, rtol=1e-6, atol=1e-4), step
            out.backward(torch.zeros_like(out))          # unpins; a zero gradient leaves the rows at their init values
            pos = 0
            for f in range(T):
                n = int(lens[f * B:(f + 1) * B].sum())
                for k in dict.fromkeys(ids[pos:pos + n].tolist()):
                    if k not in stored[f]:
                        seen[f][k] = seen[f].get(k, 0) + 1
                        if seen[f][k] >= thr:
                            stored[f].add(k)
                            del seen[f][k]
                pos += n
            assert [_stored_keys(m, t) for t in range(T)] == stored, step


def scenario_counter_checkpoint(dev, tmp_path):
    """Counter files of a dump (`<table>_counter_keys / _counter_frequencies`, the reference's names) restore the waiting keys."""
    if True:
        m = _module({"device": dev}, 3)
        m.train()
        ids = torch.tensor([5, 6, 7, 5, 6, 5], dtype=torch.int64, device=dev)
        off = torch.arange(0, ids.numel() + 1, dtype=torch.int64, device=dev)
        for _ in range(2):
            out = m(ids, off)
            out.backward(torch.zeros_like(out))
        m.dump(str(tmp_path), optim=True, counter=True)
        import os
        names = sorted(os.listdir(tmp_path))
        assert "t0_counter_keys.rank_0.world_size_1" in names and "t0_counter_frequencies.rank_0.world_size_1" in names
        m2 = _module({"device": dev}, 3)
        m2.load(str(tmp_path), optim=True, counter=True)
        ck, cs, _ = next(m2._admission_counter.table_.export(0))
        assert dict(zip(ck.tolist(), cs.tolist())) == {5: 2, 6: 2, 7: 2}
        m2.train()
        out = m2(ids, off)                    # third presentation: all three are admitted now
        out.backward(torch.zeros_like(out))
        assert _stored_keys(m2) == {5, 6, 7}


def test_admission_needs_counter_and_excludes_graph_step():
    from dynamicemb import DynamicEmbTableOptions, FrequencyAdmissionStrategy, BatchedDynamicEmbeddingTablesV2
    with patched_module():
        with pytest.raises(ValueError):
            BatchedDynamicEmbeddingTablesV2([DynamicEmbTableOptions(dim=8, max_capacity=256, admit_strategy=FrequencyAdmissionStrategy(2))])
        m = _module({}, 2)
        m.train()
        with pytest.raises(AssertionError):
            m.make_graphed_step(torch.zeros(4, dtype=torch.int64), torch.arange(5), torch.zeros(4, 32))


def scenario_lfu(dev):
    """With a frequency score (LFU) the counter advances by the key's occurrences in the batch (batched_dynamicemb_function.py:616-625),
    so a key seen `threshold` times inside ONE batch is admitted at once; the table's LFU score starts at that count."""
    from dynamicemb import DynamicEmbScoreStrategy
    if True:
        m = _module({"device": dev}, 3, score_strategy=DynamicEmbScoreStrategy.LFU)
        m.train()
        ids = torch.tensor([7, 7, 7, 8, 8, 9], dtype=torch.int64, device=dev)
        off = torch.arange(0, ids.numel() + 1, dtype=torch.int64, device=dev)
        out = m(ids, off)
        out.backward(torch.zeros_like(out))
        assert _stored_keys(m) == {7}
        keys, scores, _ = next(m.tables.export(0))
        assert keys.tolist() == [7] and scores.tolist() == [3]
        out = m(ids, off)
        out.backward(torch.zeros_like(out))
        assert _stored_keys(m) == {7, 8}                      # 8: 2 + 2 >= 3; 9: 1 + 1 < 3
        ck, cs, _ = next(m._admission_counter.table_.export(0))
        assert dict(zip(ck.tolist(), cs.tolist())) == {9: 2}


def scenario_counter_dictionary(cuda):
    """MultiTableKVCounter.add / erase against a dictionary over random batches of unique keys in two logical tables."""
    import numpy as np
    from dynamicemb import KVCounter, MultiTableKVCounter
    rng = np.random.default_rng(3)
    c = MultiTableKVCounter([KVCounter(8192), KVCounter(4096)], device=cuda)
    model = {}
    for _ in range(6):
        n = 500
        keys = rng.choice(3000, size=n, replace=False).astype(np.int64)
        tids = np.sort(rng.integers(0, 2, size=n)).astype(np.int64)
        freq = rng.integers(1, 5, size=n).astype(np.int64)
        got = c.add(torch.from_numpy(keys).to(cuda), torch.from_numpy(tids).to(cuda), torch.from_numpy(freq).to(cuda)).cpu().numpy()
        for i in range(n):
            model[(int(tids[i]), int(keys[i]))] = model.get((int(tids[i]), int(keys[i])), 0) + int(freq[i])
            assert got[i] == model[(int(tids[i]), int(keys[i]))]
        drop = rng.permutation(n)[:100]
        c.erase(torch.from_numpy(keys[drop]).to(cuda), torch.from_numpy(tids[drop]).to(cuda))
        for i in drop:
            model.pop((int(tids[i]), int(keys[i])), None)


# ---------------------------------------------------------------------------------------------------------------------------------
# the scenarios above on the CPU shim (tests/test_zz_admission_gpu.py runs the same functions on the kernels)
CPU = torch.device("cpu")


def test_module_admission_sequence_cpu_shim():
    with patched_module():
        scenario_sequence(CPU)


@pytest.mark.parametrize("mean", [False, True])
def test_module_admission_pooled_two_tables_cpu_shim(mean):
    with patched_module():
        scenario_pooled_two_tables(CPU, mean)


def test_counter_dump_load_and_module_checkpoint_cpu_shim(tmp_path):
    with patched_module():
        scenario_counter_checkpoint(CPU, tmp_path)


def test_module_admission_lfu_counts_occurrences_cpu_shim():
    with patched_module():
        scenario_lfu(CPU)


def test_counter_matches_dictionary_cpu_shim():
    with patched_module():
        scenario_counter_dictionary(CPU)


def test_module_without_admission_op_by_op_cpu_shim():
    """The op-by-op prefetch without a strategy: every new key is inserted at once and trains (host logic of the reference's op order)."""
    with patched_module():
        m = _module({"fused_prefetch": False}, None)
        m.train()
        ids = torch.tensor([3, 4, 3, 5], dtype=torch.int64)
        off = torch.arange(0, 5, dtype=torch.int64)
        out = m(ids, off)
        assert out[:, 0].tolist() == [3.0, 4.0, 3.0, 5.0]
        out.backward(torch.ones_like(out))
        snap = _snapshot(m)
        assert set(snap) == {3, 4, 5} and snap[3][0].item() == 3.0 - 0.5 * 2 and snap[4][0].item() == 4.0 - 0.5


def test_incremental_dump_cpu_shim():
    """STEP scores: rows touched at or after the given step are returned with their embeddings; unknown tables warn."""
    with patched_module():
        m = _module({"fused_prefetch": False}, None)
        m.train()
        off = lambda n: torch.arange(0, n + 1, dtype=torch.int64)       # noqa: E731
        for ids in ([1, 2, 3], [3, 4], [5]):
            x = torch.tensor(ids, dtype=torch.int64)
            out = m(x, off(len(ids)))
            out.backward(torch.zeros_like(out))
        # STEP score of a row = the step of its last access (1, 2, 3)
        tensors, nxt = m.incremental_dump({"t0": 2})
        keys, vals = tensors["t0"]
        assert sorted(keys.tolist()) == [3, 4, 5] and vals.shape == (3, 32) and keys.device.type == "cpu"
        assert all(float(v[0]) == float(k) for k, v in zip(keys.tolist(), vals))
        assert nxt == {"t0": m.get_score()["t0"]}
        assert sorted(m.incremental_dump({"t0": 0})[0]["t0"][0].tolist()) == [1, 2, 3, 4, 5]
        with pytest.warns(UserWarning):
            assert m.incremental_dump({"nope": 1}) == ({}, {})
        # the table-level form (ScoredHashTable.incremental_dump): keys, {score name: scores}, slot indices
        k, sc, ix = m.tables.incremental_dump({"score": 2}, table_id=0, return_index=True)
        assert sorted(zip(k.tolist(), sc["score"].tolist())) == [(3, 2), (4, 2), (5, 3)] and ix.numel() == 3 and k.device.type == "cpu"
        k2, sc2 = m.tables.incremental_dump({"unknown": 0}, 0)
        assert k2.numel() == 5 and sc2 == {}


def test_small_public_surface_and_fill_tables_cpu_shim():
    """cache-tier no-ops, enable_prefetch, split_embedding_weights, fill_tables (reference batched_dynamicemb_tables.py:942-1000,:1182)."""
    with patched_module():
        m = _module({"fused_prefetch": False}, None, T=2, cap=1024)
        assert m.cache is None and m.enable_prefetch is False
        m.enable_prefetch = True
        assert m.enable_prefetch is True
        m.reset_cache_states(), m.set_record_cache_metrics(True)
        assert [tuple(w.shape) for w in m.split_embedding_weights()] == [(1, 1), (1, 1)]
        torch.manual_seed(0)
        m.fill_tables(0.5)
        for t in range(2):
            assert abs(m.tables.size(t) - 512) <= 8, m.tables.size(t)       # a full bucket evicts instead of growing: a few keys may be lost
        m.fill_tables(2.0)                                                   # clamped to 0.95
        for t in range(2):
            assert 0.9 * 1024 <= m.tables.size(t) <= int(0.95 * 1024)
        with pytest.raises(ValueError):
            m.fill_tables(-0.1)
