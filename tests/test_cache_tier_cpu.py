"""Cache tier host logic on the CPU (cache half): `caching=True` = an HBM cache table in front of a backing table
whose value rows live in host memory (reference `_prefetch_cache_path`, batched_dynamicemb_function.py:296-556).  The module runs on
the oracle-backed op-layer shim (tests/cpu_ext_shim.py); the model is a dictionary key -> row that knows nothing about tiers, so every
assertion crosses cache eviction, write-back and re-fetch.  tests/test_zz_cache_tier_gpu.py runs the same scenarios on the kernels."""
import numpy as np
import pytest
import torch

from tests.cpu_ext_shim import patched_module
from tests.test_admission_cpu import _module, _snapshot

D, LR = 32, 0.5          # 32 / 64 / 128 are the embedding widths the GPU suite validates


def _cached_module(dev, threshold=None, T=1, pooling=None, score_strategy=None, cap=8192):
    # value row = 32 fp32 (SGD): HBM budget of 1024 rows per table => cache of 1024 slots (one 1024-slot bucket) per table
    return _module({"device": dev}, threshold, T=T, pooling=pooling, score_strategy=score_strategy, cap=cap, dim=D, caching=True,
                   local_hbm=1024 * D * 4)


def scenario_cache_train_evict_refetch(dev, score_strategy=None, key_space=3000):
    """Six training steps of 600 distinct keys out of `key_space` through a 1024-row cache: outputs always equal the dictionary model
    (rows come back from the backing table with every update they received while cached), and after flush() the backing table holds
    exactly the model.  With a frequency score new keys carry the LOWEST score of a full bucket and would evict each other inside one
    insert call (the reference's kernel has the same property); that strategy runs with a key space the cache can hold and checks that
    the frequencies reach the backing table instead."""
    rng = np.random.default_rng(0)
    m = _cached_module(dev, score_strategy=score_strategy)
    assert m.cache is not None and m.cache.capacity() == 1024 and m._values.shape[0] == 8192
    m.train()
    model, seen = {}, {}
    for step in range(6):
        uniq = rng.choice(np.arange(1, key_space + 1), size=600, replace=False)
        ids = np.concatenate([uniq, rng.choice(uniq, size=200)]).astype(np.int64)       # some repeats
        rng.shuffle(ids)
        x = torch.from_numpy(ids).to(dev)
        off = torch.arange(0, ids.size + 1, dtype=torch.int64, device=dev)
        out = m(x, off)
        want = torch.tensor([model.get(int(k), float(k % 100000)) for k in ids], dtype=torch.float32)
        assert torch.equal(out[:, 0].cpu(), want), f"step {step}: forward differs from the model"
        assert bool((out == out[:, :1]).all())
        out.backward(torch.ones_like(out))
        for k in ids.tolist():
            model[k] = model.get(k, float(k % 100000)) - LR
            seen[k] = seen.get(k, 0) + 1
        assert m.cache.size() <= 1024
    assert (len(model) > 1024) == (key_space > 1024), "the scenario must overflow the cache exactly when the key space does"
    m.flush()
    if key_space <= 1024:           # frequency scores: the count of every key travelled cache -> backing table
        keys, scores, _ = next(m.tables.export(0))
        assert dict(zip(keys.tolist(), scores.tolist())) == seen
    snap = _snapshot(m)
    assert set(snap) == set(model)
    for k, row in snap.items():
        assert float(row[0]) == model[k] and bool((row[:D] == row[0]).all()), k
    # eval: cached, backing-only and absent ids in one batch
    m.eval()
    some = list(model)[:50] + [5000, 5001]
    out = m(torch.tensor(some, dtype=torch.int64, device=dev), torch.arange(0, len(some) + 1, dtype=torch.int64, device=dev))
    want = torch.tensor([model.get(k, 0.0) for k in some], dtype=torch.float32)
    assert torch.equal(out[:, 0].cpu(), want)
    # reset_cache_states drops the cache; the backing table alone still answers
    m.reset_cache_states()
    assert m.cache.size() == 0
    out = m(torch.tensor(some, dtype=torch.int64, device=dev), torch.arange(0, len(some) + 1, dtype=torch.int64, device=dev))
    assert torch.equal(out[:, 0].cpu(), want)


def scenario_cache_pooled_two_tables(dev, mean):
    """SUM / MEAN pooling over two cached tables, forward values against the model; checkpoint round trip through the backing table."""
    from dynamicemb import DynamicEmbPoolingMode
    rng = np.random.default_rng(2)
    T, B = 2, 64
    m = _cached_module(dev, T=T, pooling=DynamicEmbPoolingMode.MEAN if mean else DynamicEmbPoolingMode.SUM)
    m.train()
    model = [{}, {}]
    for step in range(4):
        lens = rng.integers(0, 12, size=T * B)
        ids = rng.integers(1, 2500, size=int(lens.sum())).astype(np.int64)
        off_np = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
        out = m(torch.from_numpy(ids).to(dev), torch.from_numpy(off_np).to(dev))
        want = torch.zeros(B, T * D, dtype=torch.float64)
        for f in range(T):
            for b in range(B):
                s, e = off_np[f * B + b], off_np[f * B + b + 1]
                for k in ids[s:e].tolist():
                    want[b, f * D:(f + 1) * D] += model[f].get(k, float(k % 100000)) / ((e - s) if mean else 1)
        assert torch.allclose(out.cpu().double(), want, rtol=1e-5, atol=1e-2), f"step {step}"
        g = torch.ones_like(out)
        out.backward(g)
        for f in range(T):
            for b in range(B):
                s, e = off_np[f * B + b], off_np[f * B + b + 1]
                for k in ids[s:e].tolist():
                    model[f][k] = model[f].get(k, float(k % 100000)) - LR / ((e - s) if mean else 1)
    m.flush()
    for f in range(T):
        snap = _snapshot(m, f)
        assert set(snap) == set(model[f])
        for k, row in snap.items():
            assert abs(float(row[0]) - model[f][k]) <= 1e-3 * max(1.0, abs(model[f][k])), (f, k)


def scenario_cache_with_admission(dev):
    """Admission in front of the cache (reference :377-410): a key enters the cache at its second presentation, not before; keys the
    backing table already holds are never counted."""
    m = _cached_module(dev, threshold=2)
    m.train()
    ids = torch.arange(1, 301, dtype=torch.int64, device=dev)
    off = torch.arange(0, 301, dtype=torch.int64, device=dev)
    out = m(ids, off)
    assert torch.equal(out[:, 0].cpu(), torch.arange(1, 301, dtype=torch.float32))      # initializer rows, unstored
    out.backward(torch.ones_like(out))
    assert m.cache.size() == 0 and m.tables.size() == 0
    out = m(ids, off)
    out.backward(torch.ones_like(out))
    assert m.cache.size() == 300
    out = m(ids, off)
    assert torch.equal(out[:, 0].cpu(), torch.arange(1, 301, dtype=torch.float32) - LR)  # trained once (step 2)
    out.backward(torch.zeros_like(out))
    m.flush()
    assert set(_snapshot(m)) == set(range(1, 301))


CPU = torch.device("cpu")


@pytest.mark.parametrize("strategy", ["step", "lfu", "timestamp", "no_eviction"])
def test_cache_train_evict_refetch_cpu_shim(strategy):
    from dynamicemb import DynamicEmbScoreStrategy as S
    with patched_module():
        scenario_cache_train_evict_refetch(CPU, {"step": S.STEP, "lfu": S.LFU, "timestamp": S.TIMESTAMP, "no_eviction": S.NO_EVICTION}[strategy],
                                           key_space=900 if strategy == "lfu" else 3000)


@pytest.mark.parametrize("mean", [False, True])
def test_cache_pooled_two_tables_cpu_shim(mean):
    with patched_module():
        scenario_cache_pooled_two_tables(CPU, mean)


def test_cache_with_admission_cpu_shim():
    with patched_module():
        scenario_cache_with_admission(CPU)


def test_cache_configuration_rules_cpu_shim():
    """Budget rules of the reference (batched_dynamicemb_tables.py:637-700): everything fits -> HBM only; no budget -> error."""
    from dynamicemb import DynamicEmbScoreStrategy as S
    with patched_module():
        m = _module({}, None, cap=1024, caching=True, local_hbm=1 << 30)
        assert m.cache is None and m._values.shape[0] == 1024
        with pytest.raises(ValueError):
            _module({}, None, cap=8192, caching=True, local_hbm=0)
        with pytest.raises(NotImplementedError):
            _module({}, None, cap=8192, caching=True, local_hbm=1024 * D * 4, score_strategy=(S.TIMESTAMP, S.LFU))
        m = _cached_module(CPU)
        m.train()
        with pytest.raises(AssertionError):
            m.make_graphed_step(torch.zeros(4, dtype=torch.int64), torch.arange(5), torch.zeros(4, D))


def test_cache_checkpoint_round_trip_cpu_shim(tmp_path):
    """dump() flushes the cache first; load() drops a stale cache: a cached module's checkpoint restores into another cached module and
    into an HBM-direct one with identical lookups."""
    with patched_module():
        m = _cached_module(CPU)
        m.train()
        ids = torch.arange(1, 1501, dtype=torch.int64)                   # more keys than the cache holds
        for lo in (0, 500, 1000):
            x = ids[lo:lo + 500]
            out = m(x, torch.arange(0, 501, dtype=torch.int64))
            out.backward(torch.ones_like(out))
        m.dump(str(tmp_path), optim=True)
        want = ids.to(torch.float32) - LR
        for fresh in (_cached_module(CPU), _module({"device": CPU, "fused_prefetch": False}, None, cap=8192, dim=D)):
            fresh.train()
            stale = fresh(ids[:10], torch.arange(0, 11, dtype=torch.int64))       # something in the cache / table before the load
            stale.backward(torch.zeros_like(stale))
            fresh.load(str(tmp_path), optim=True)
            fresh.eval()
            out = fresh(ids, torch.arange(0, 1501, dtype=torch.int64))
            assert torch.equal(out[:, 0], want)


# ---------------------------------------------------------------------------------------------------------------------------------
# hybrid storage: 0 < local_hbm_for_values < table bytes WITHOUT caching = two disjoint tiers (reference HybridStorage,
# key_value_table.py:2107-2404): new keys enter the HBM tier, its evictions move to the host tier, host-tier keys stay there
def _hybrid_module(dev, threshold=None, T=1, pooling=None, score_strategy=None, cap=8192):
    return _module({"device": dev}, threshold, T=T, pooling=pooling, score_strategy=score_strategy, cap=cap, dim=D, caching=False,
                   local_hbm=1024 * D * 4)


def scenario_hybrid_train(dev, score_strategy=None):
    """Six training steps of 600 distinct keys out of 3000: the HBM tier (1024 rows) overflows into the host tier; outputs follow the
    tier-agnostic dictionary model at every step (rows are trained in whichever tier they sit), the tiers stay disjoint, an export sees
    both, a checkpoint round-trips."""
    rng = np.random.default_rng(5)
    m = _hybrid_module(dev, score_strategy=score_strategy)
    assert m._hybrid and m.cache is None and m._cache.capacity() == 1024 and m.tables.capacity() == 7168
    m.train()
    model = {}
    for step in range(6):
        uniq = rng.choice(np.arange(1, 3001), size=600, replace=False)
        ids = np.concatenate([uniq, rng.choice(uniq, size=200)]).astype(np.int64)
        rng.shuffle(ids)
        out = m(torch.from_numpy(ids).to(dev), torch.arange(0, ids.size + 1, dtype=torch.int64, device=dev))
        want = torch.tensor([model.get(int(k), float(k % 100000)) for k in ids], dtype=torch.float32)
        assert torch.equal(out[:, 0].cpu(), want), f"step {step}: forward differs from the model"
        out.backward(torch.ones_like(out))
        for k in ids.tolist():
            model[k] = model.get(k, float(k % 100000)) - LR
    hot = set(next(m._cache.export(0))[0].tolist())
    cold = set()
    for keys, _, _ in m.tables.export(0):
        cold |= set(keys.tolist())
    assert hot and cold and not (hot & cold) and (hot | cold) == set(model), "two disjoint tiers that together hold every key"
    assert m._cache._ref_counter.sum().item() == 0 and m.tables._ref_counter.sum().item() == 0, "every pin released"
    snap = _snapshot(m)
    assert set(snap) == set(model) and all(float(snap[k][0]) == model[k] for k in model)
    m.eval()
    some = list(model)[:40] + [5000]
    out = m(torch.tensor(some, dtype=torch.int64, device=dev), torch.arange(0, len(some) + 1, dtype=torch.int64, device=dev))
    assert torch.equal(out[:, 0].cpu(), torch.tensor([model.get(k, 0.0) for k in some], dtype=torch.float32))
    return m, model


def scenario_hybrid_pooled_checkpoint(dev, tmp_path):
    from dynamicemb import DynamicEmbPoolingMode
    rng = np.random.default_rng(6)
    T, B = 2, 64
    m = _hybrid_module(dev, T=T, pooling=DynamicEmbPoolingMode.SUM)
    m.train()
    model = [{}, {}]
    for step in range(4):
        lens = rng.integers(0, 12, size=T * B)
        ids = rng.integers(1, 2500, size=int(lens.sum())).astype(np.int64)
        off_np = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
        out = m(torch.from_numpy(ids).to(dev), torch.from_numpy(off_np).to(dev))
        want = torch.zeros(B, T * D, dtype=torch.float64)
        for f in range(T):
            for b in range(B):
                for k in ids[off_np[f * B + b]:off_np[f * B + b + 1]].tolist():
                    want[b, f * D:(f + 1) * D] += model[f].get(k, float(k % 100000))
        assert torch.allclose(out.cpu().double(), want, rtol=1e-5, atol=1e-2), f"step {step}"
        out.backward(torch.ones_like(out))
        for f in range(T):
            for b in range(B):
                for k in ids[off_np[f * B + b]:off_np[f * B + b + 1]].tolist():
                    model[f][k] = model[f].get(k, float(k % 100000)) - LR
    m.dump(str(tmp_path), optim=True)
    m2 = _hybrid_module(dev, T=T, pooling=DynamicEmbPoolingMode.SUM)
    m2.load(str(tmp_path), optim=True)
    for f in range(T):
        snap = _snapshot(m2, f)
        assert set(snap) == set(model[f])
        assert all(abs(float(snap[k][0]) - model[f][k]) <= 1e-3 * max(1.0, abs(model[f][k])) for k in model[f])


@pytest.mark.parametrize("strategy", ["step", "timestamp"])
def test_hybrid_train_cpu_shim(strategy):
    from dynamicemb import DynamicEmbScoreStrategy as S
    with patched_module():
        scenario_hybrid_train(CPU, {"step": S.STEP, "timestamp": S.TIMESTAMP}[strategy])


def test_hybrid_pooled_checkpoint_cpu_shim(tmp_path):
    with patched_module():
        scenario_hybrid_pooled_checkpoint(CPU, tmp_path)


@pytest.mark.parametrize("tier", ["cache", "hybrid"])
def test_tiered_eval_absent_constant_cpu_shim(tier):
    """Eval through both tiers with a non-zero eval constant: stored ids read their row (wherever it sits), absent ids the constant —
    sequence and SUM-pooled, two tables."""
    from dynamicemb import (BatchedDynamicEmbeddingTablesV2, DynamicEmbInitializerArgs, DynamicEmbInitializerMode as IM, DynamicEmbPoolingMode as P,
                            DynamicEmbScoreStrategy, DynamicEmbTableOptions, EmbOptimType)
    with patched_module():
        for pooling in (P.NONE, P.SUM):
            opts = [DynamicEmbTableOptions(dim=D, max_capacity=8192, bucket_capacity=128, score_strategy=DynamicEmbScoreStrategy.STEP,
                                           caching=(tier == "cache"), local_hbm_for_values=1024 * D * 4,
                                           initializer_args=DynamicEmbInitializerArgs(mode=IM.DEBUG),
                                           eval_initializer_args=DynamicEmbInitializerArgs(mode=IM.CONSTANT, value=-3.0)) for _ in range(2)]
            m = BatchedDynamicEmbeddingTablesV2(opts, table_names=["a", "b"], feature_table_map=[0, 1], pooling_mode=pooling, optimizer=EmbOptimType.SGD,
                                                learning_rate=LR, device=CPU)
            assert (m._caching, m._hybrid) == ((True, False) if tier == "cache" else (False, True))
            m.train()
            # fill past the 1024-row hot tier so that stored keys sit in both tiers
            for lo in range(1, 2401, 600):
                ids = torch.arange(lo, lo + 600, dtype=torch.int64).repeat(2)             # the same keys in table a and table b
                off = (torch.arange(0, 1201, dtype=torch.int64) if pooling == P.NONE else torch.arange(0, 1201, 75, dtype=torch.int64))
                out = m(ids, off)
                out.backward(torch.zeros_like(out))
            m.eval()
            q = torch.tensor([5, 2399, 9000, 700, 9001, 1500, 9002, 42], dtype=torch.int64)     # per table: 2 stored, 1 absent, 1 stored
            off = torch.arange(0, 9, dtype=torch.int64) if pooling == P.NONE else torch.tensor([0, 2, 4, 6, 8], dtype=torch.int64)
            out = m(q, off)
            val = lambda k: float(k) if k < 2401 else -3.0                               # noqa: E731
            if pooling == P.NONE:
                assert out[:, 0].tolist() == [val(int(k)) for k in q]
            else:                                                                        # [B=2, 2 tables]: bags of 2 ids
                want = [[val(5) + val(2399), val(9001) + val(1500)], [val(9000) + val(700), val(9002) + val(42)]]
                assert [[float(out[b, f * D]) for f in range(2)] for b in range(2)] == want
