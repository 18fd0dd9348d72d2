"""GPU: module.dump / module.load in the reference's checkpoint layout (SURVEY 8(f) row 3;
corelib/dynamicemb/dynamicemb/batched_dynamicemb_tables.py:1262-1410, key_value_table.py:1134-1520).
The reference's own check (test/unit_tests/test_embedding_dump_load.py) is: train, dump, load into a fresh model, same lookups.
Here additionally the FILE BYTES are checked against what the table holds (they are what the reference would read: the CPU suite pins the
byte layout to the reference's reader), and the fresh module has another capacity, so every key lands in another slot."""
import json
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _module(cuda, opt, strategy, D=32, cap=4096, names=("ta", "tb"), lr=0.1, **kw):
    from dynamicemb import (BatchedDynamicEmbeddingTablesV2, DynamicEmbInitializerArgs, DynamicEmbInitializerMode, DynamicEmbPoolingMode,
                            DynamicEmbTableOptions)
    opts = [DynamicEmbTableOptions(dim=D, max_capacity=cap, local_hbm_for_values=1 << 40, score_strategy=strategy,
                                   initializer_args=DynamicEmbInitializerArgs(mode=DynamicEmbInitializerMode.NORMAL)) for _ in names]
    return BatchedDynamicEmbeddingTablesV2(opts, table_names=list(names), pooling_mode=DynamicEmbPoolingMode.NONE, optimizer=opt, device=cuda,
                                           learning_rate=lr, initial_accumulator_value=0.5, **kw)


def _train(m, cuda, steps=4, n=3000, seed=0):
    g = torch.Generator().manual_seed(seed)
    T = len(m.table_names)
    m.train()
    for _ in range(steps):
        ids = torch.randint(0, 2500, (T * n,), generator=g).to(cuda)     # 2500 keys in 32 buckets of 128: no eviction
        off = torch.arange(0, T * n + 1, dtype=torch.int64, device=cuda)          # T features x n samples x 1 id
        out = m(ids, off)
        out.backward(torch.randn(out.shape, generator=g).to(cuda))
    torch.cuda.synchronize()


def _content(m, t):
    """{key: (value row, score words)} of one table."""
    tb = m.tables
    out = {}
    for keys, dense, scores in m._export_batches(t):
        sc = scores if scores.dim() == 2 else scores[:, None]
        for k, v, s in zip(keys.tolist(), dense.cpu(), sc.tolist()):
            out[k] = (v, s)
    assert len(out) == tb.size(t)
    return out


def _read(path, dtype, width=None):
    a = np.fromfile(path, dtype=dtype)
    return a if width is None else a.reshape(-1, width)


@pytest.mark.parametrize("opt_name", ["adagrad", "rowwise", "adam", "sgd"])
def test_dump_load_round_trip_and_file_bytes(cuda, tmp_path, opt_name):
    from dynamicemb import DynamicEmbScoreStrategy, EmbOptimType
    from dynamicemb import checkpoint as ck
    ot = {"adagrad": EmbOptimType.EXACT_ADAGRAD, "rowwise": EmbOptimType.EXACT_ROWWISE_ADAGRAD, "adam": EmbOptimType.ADAM, "sgd": EmbOptimType.SGD}[opt_name]
    D = 32
    m = _module(cuda, ot, DynamicEmbScoreStrategy.STEP, D=D)
    _train(m, cuda)
    d = str(tmp_path)
    m.dump(d, optim=True)
    n_tb = m.tables.size(1)
    sdim, cdim = m.optimizer.get_state_dim(D), m.optimizer.get_ckpt_state_dim(D)
    for t, name in enumerate(m.table_names):
        meta = json.load(open(ck.encode_meta_json_file_path(d, name)))
        assert meta["opt_type"] == m.optimizer.get_opt_args()["opt_type"] and meta["evict_strategy"] == "EvictStrategy.KCustomized"
        assert meta["dist_type"] == "roundrobin" and meta["step_score"] == m.get_score()[name]
        keys = _read(ck.encode_checkpoint_file_path(d, name, 0, 1, "keys"), np.int64)
        vals = _read(ck.encode_checkpoint_file_path(d, name, 0, 1, "values"), np.float32, D)
        scs = _read(ck.encode_checkpoint_file_path(d, name, 0, 1, "scores"), np.int64)
        have = _content(m, t)
        assert len(keys) == len(have) == len(set(keys.tolist())) and len(have) > 2000
        if cdim:
            opts = _read(ck.encode_checkpoint_file_path(d, name, 0, 1, "opt_values"), np.float32, cdim)
        for i, k in enumerate(keys.tolist()):
            row, sc = have[k]
            assert np.array_equal(vals[i], row[:D].numpy()) and scs[i] == sc[0]
            if cdim:
                assert np.array_equal(opts[i], row[D:D + cdim].numpy())
    # load into a module of another capacity (other buckets, other slots), optimizer arguments from the meta file
    m2 = _module(cuda, ot, DynamicEmbScoreStrategy.STEP, D=D, cap=8192, lr=0.7)
    m2.load(d, optim=True)
    assert m2.get_score() == m.get_score()
    if sdim:
        assert m2.optimizer.get_opt_args() == m.optimizer.get_opt_args()
    for t in range(2):
        a, b = _content(m, t), _content(m2, t)
        assert set(a) == set(b)
        for k, (row, sc) in a.items():
            assert sc == b[k][1]
            assert torch.equal(row[:D], b[k][0][:D])
            if opt_name == "rowwise":
                assert row[D] == b[k][0][D] and bool((b[k][0][D + 1:D + sdim] == 0.5).all())       # padded with the initial accumulator value
            else:
                assert torch.equal(row[D:D + sdim], b[k][0][D:D + sdim])
    # same lookups, and training continues identically from both
    ids = torch.cat([torch.arange(0, 3000, dtype=torch.int64, device=cuda)] * 2)            # both tables: all keys + 500 absent ones
    off = torch.arange(0, ids.numel() + 1, dtype=torch.int64, device=cuda)
    m.eval(); m2.eval()
    assert torch.equal(m(ids, off), m2(ids, off))
    if opt_name != "rowwise":
        m2.set_learning_rate(m.optimizer.args.learning_rate)
        _train(m, cuda, steps=1, seed=5); _train(m2, cuda, steps=1, seed=5)
        m.eval(); m2.eval()
        assert torch.equal(m(ids, off), m2(ids, off))
    # without optim=True the states start from the initial value
    m3 = _module(cuda, ot, DynamicEmbScoreStrategy.STEP, D=D)
    m3.load(d, optim=False, table_names=["tb"])
    assert m3.tables.size(0) == 0 and m3.tables.size(1) == n_tb
    if sdim and opt_name in ("adagrad", "rowwise"):
        _, vals3 = m3.export_keys_values(1)
        assert bool((vals3[:, D:D + sdim] == 0.5).all())


def test_lru_scores_are_stored_as_age(cuda, tmp_path):
    """TIMESTAMP (LRU) tables store `dump time - last access`; a later load restores `load time - age`: order and gaps survive."""
    from dynamicemb import DynamicEmbScoreStrategy, EmbOptimType
    from dynamicemb import checkpoint as ck
    m = _module(cuda, EmbOptimType.SGD, DynamicEmbScoreStrategy.TIMESTAMP, names=("t",))
    _train(m, cuda, steps=3)
    d = str(tmp_path)
    m.dump(d)
    meta = json.load(open(ck.encode_meta_json_file_path(d, "t")))
    assert meta["evict_strategy"] == "EvictStrategy.KLru"
    ages = _read(ck.encode_checkpoint_file_path(d, "t", 0, 1, "scores"), np.int64)
    keys = _read(ck.encode_checkpoint_file_path(d, "t", 0, 1, "keys"), np.int64)
    have = _content(m, 0)
    assert (ages >= 0).all() and len(set(ages.tolist())) >= 3                       # three steps, three timestamps
    stamp = {k: have[k][1][0] for k in keys.tolist()}
    base = ages[0] + stamp[int(keys[0])]
    assert all(a + stamp[k] == base for a, k in zip(ages.tolist(), keys.tolist()))   # one dump timestamp
    m2 = _module(cuda, EmbOptimType.SGD, DynamicEmbScoreStrategy.TIMESTAMP, names=("t",))
    m2.load(d)
    got = _content(m2, 0)
    shift = got[int(keys[0])][1][0] - stamp[int(keys[0])]
    assert shift > 0 and all(got[k][1][0] - stamp[k] == shift for k in stamp)
    with pytest.raises(ValueError, match="Evict strategy mismatch"):
        _module(cuda, EmbOptimType.SGD, DynamicEmbScoreStrategy.LFU, names=("t",)).load(d)


@pytest.mark.parametrize("order", ["ts_lfu", "lfu_ts"])
def test_compound_scores_keep_both_words_in_logical_order(cuda, tmp_path, order):
    from dynamicemb import DynamicEmbScoreStrategy as S, EmbOptimType
    from dynamicemb import checkpoint as ck
    st = (S.TIMESTAMP, S.LFU) if order == "ts_lfu" else (S.LFU, S.TIMESTAMP)
    m = _module(cuda, EmbOptimType.SGD, st, names=("t",))
    _train(m, cuda, steps=3)
    d = str(tmp_path)
    m.dump(d)
    keys = _read(ck.encode_checkpoint_file_path(d, "t", 0, 1, "keys"), np.int64)
    blocks = _read(ck.encode_checkpoint_file_path(d, "t", 0, 1, "scores"), np.int64, 2)
    have = _content(m, 0)
    fcol = 1 if order == "ts_lfu" else 0
    for k, b in zip(keys.tolist(), blocks.tolist()):
        ts_word, freq_word = have[k][1]                                            # device order: (timestamp, frequency)
        assert b[fcol] == freq_word and b[1 - fcol] == ts_word
    assert blocks[:, fcol].min() >= 1 and blocks[:, fcol].max() > 1                # frequencies: some key was seen more than once
    m2 = _module(cuda, EmbOptimType.SGD, st, names=("t",), cap=8192)
    m2.load(d)
    got = _content(m2, 0)
    assert set(got) == set(have)
    for k in have:
        assert got[k][1] == have[k][1] and torch.equal(got[k][0], have[k][0])


def test_resharded_load_keeps_the_keys_each_rank_owns(cuda, tmp_path):
    """A checkpoint written by 2 ranks: one rank loads all of it; 3 ranks partition it by the owner rule of the table's dist_type."""
    from dynamicemb import DynamicEmbScoreStrategy, EmbOptimType
    from dynamicemb import checkpoint as ck
    D = 32
    src = _module(cuda, EmbOptimType.EXACT_ADAGRAD, DynamicEmbScoreStrategy.STEP, D=D, names=("t",))
    _train(src, cuda, steps=2)
    have = _content(src, 0)
    d = str(tmp_path)
    src.dump(d, optim=True)                                                        # meta file + the single-rank files ...
    one = {item: ck.encode_checkpoint_file_path(d, "t", 0, 1, item) for item in ("keys", "values", "scores", "opt_values")}
    keys = torch.from_numpy(_read(one["keys"], np.int64))
    vals = torch.from_numpy(_read(one["values"], np.float32, D))
    scs = torch.from_numpy(_read(one["scores"], np.int64))
    opts = torch.from_numpy(_read(one["opt_values"], np.float32, D))
    for p in one.values():
        os.remove(p)
    for r in range(2):                                                             # ... rewritten as the two files of a 2-rank job
        sel = keys % 2 == r
        with ck.TableFileWriter(*(ck.encode_checkpoint_file_path(d, "t", r, 2, item) for item in ("keys", "values", "scores", "opt_values"))) as w:
            w.write(keys[sel], vals[sel], scs[sel], opts[sel])
    m1 = _module(cuda, EmbOptimType.EXACT_ADAGRAD, DynamicEmbScoreStrategy.STEP, D=D, names=("t",))
    m1.load(d, optim=True)                                                         # world 1 reads both files
    got = _content(m1, 0)
    assert set(got) == set(have) and all(torch.equal(got[k][0], have[k][0]) for k in have)
    meta = ck.encode_meta_json_file_path(d, "t")
    for dist_type in ("roundrobin", "hash_roundrobin"):
        js = json.load(open(meta)); js["dist_type"] = dist_type
        json.dump(js, open(meta, "w"))
        seen = {}
        for r in range(3):
            mr = _module(cuda, EmbOptimType.EXACT_ADAGRAD, DynamicEmbScoreStrategy.STEP, D=D, names=("t",))
            mr._dynamicemb_options[0].dist_type = dist_type
            for f in range(2):
                mr._load_table_files(0, meta, *(ck.encode_checkpoint_file_path(d, "t", f, 2, item) for item in ("keys", "values", "scores", "opt_values")),
                                     include_optim=True, timestamp=0, filter_rank=(r, 3))
            c = _content(mr, 0)
            ks = torch.tensor(sorted(c), dtype=torch.int64)
            assert bool((ck.owner_rank(ks, 3, dist_type) == r).all())
            assert not (set(c) & set(seen))
            seen.update(c)
        assert set(seen) == set(have) and all(torch.equal(seen[k][0], have[k][0]) for k in have)


def test_export_keys_values_reference_signature(cuda):
    from dynamicemb import DynamicEmbScoreStrategy, EmbOptimType
    m = _module(cuda, EmbOptimType.EXACT_ADAGRAD, DynamicEmbScoreStrategy.STEP, D=32)
    _train(m, cuda, steps=1)
    keys, emb = m.export_keys_values("tb", torch.device("cpu"))
    assert keys.device.type == "cpu" and emb.shape == (keys.numel(), 32) and keys.numel() == m.tables.size(1)
    k2, rows = m.export_keys_values(1)
    o1, o2 = torch.argsort(keys), torch.argsort(k2.cpu())                    # the table scan compacts with an atomic counter: no fixed order
    assert torch.equal(k2.cpu()[o2], keys[o1]) and rows.shape[1] == m.value_dim and torch.equal(rows[:, :32].cpu()[o2], emb[o1])
