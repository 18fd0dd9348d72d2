"""Checkpoint wire format, CPU side: file names, discovery, re-sharding file selection, score column order,
optimizer state width and the byte layout of the four per-rank files — against `tests/golden/checkpoint.json`, produced by executing the
reference's own functions (tests/golden/gen_golden_ckpt.py).  The device half (module.dump / load) is tests/test_checkpoint_gpu.py."""
import json
import os

import numpy as np
import pytest
import torch

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def gold():
    with open(os.path.join(G, "checkpoint.json")) as f:
        return json.load(f)


def test_file_names_match_reference(gold):
    from dynamicemb import checkpoint as ck
    assert ck.encode_meta_json_file_path("ROOT", "t_user") == gold["names"]["meta"]
    for item, want in gold["names"]["ckpt"].items():
        assert ck.encode_checkpoint_file_path("ROOT", "t_user", 3, 8, item) == want
    for item, want in gold["names"]["counter"].items():
        assert ck.encode_counter_checkpoint_file_path("ROOT", "t_user", 1, 2, item) == want


def test_loading_file_selection_matches_reference(gold, tmp_path):
    from dynamicemb import checkpoint as ck
    d = str(tmp_path)
    for r in range(2):
        for item in ("keys", "values", "scores", "opt_values"):
            open(ck.encode_checkpoint_file_path(d, "tab", r, 2, item), "wb").close()
        for item in ("keys", "values"):
            open(ck.encode_checkpoint_file_path(d, "bare", r, 2, item), "wb").close()
    rel = lambda lists: [[os.path.relpath(p, d) for p in l] for l in lists]      # noqa: E731
    L = gold["loading"]
    assert rel(ck.get_loading_files(d, "tab", rank=1, world_size=2)) == L["tab_rank1_world2"]          # own files
    assert rel(ck.get_loading_files(d, "tab", rank=0, world_size=1)) == L["tab_rank0_world1"]          # fewer ranks: all files
    assert rel(ck.get_loading_files(d, "tab", rank=2, world_size=3)) == L["tab_rank2_world3"]          # more ranks: all files
    assert rel(ck.get_loading_files(d, "bare", rank=0, world_size=2)) == L["bare_rank0_world2"]        # no score / optimizer files
    assert rel(ck.get_loading_files(d, "nothing", rank=0, world_size=1)) == L["missing_rank0_world1"]
    os.remove(ck.encode_checkpoint_file_path(d, "tab", 1, 2, "keys"))
    with pytest.raises(RuntimeError) as e:
        ck.get_loading_files(d, "tab", rank=0, world_size=2)
    assert str(e.value).replace(d, "ROOT") == L["corrupt_error"]
    with pytest.raises(RuntimeError):
        ck.get_loading_files(os.path.join(d, "absent"), "tab", rank=0, world_size=1)


def test_score_column_order_matches_reference(gold):
    from dynamicemb import checkpoint as ck
    from dynamicemb.types import DynamicEmbScoreStrategy as S
    cases = {"timestamp": S.TIMESTAMP, "lfu": S.LFU, "ts_lfu": (S.TIMESTAMP, S.LFU), "lfu_ts": (S.LFU, S.TIMESTAMP)}
    for name, st in cases.items():
        want = gold["perms"][name]
        assert [int(x) for x in ck.get_physical_score_order(st)] == want["physical"]
        assert ck.score_dump_permutation(st) == want["dump"]
        assert ck.score_load_permutation(st) == want["load"]
    with pytest.raises(NotImplementedError):
        ck.get_physical_score_order((S.STEP, S.LFU))


def test_optimizer_state_width_matches_reference(gold):
    from dynamicemb import checkpoint as ck
    from dynamicemb.optimizer import OptimizerArgs, SparseOptimizer
    from dynamicemb.types import EmbOptimType
    opt = SparseOptimizer(EmbOptimType.EXACT_ROWWISE_ADAGRAD, OptimizerArgs())
    assert (opt.get_state_dim(8), opt.get_ckpt_state_dim(8)) == (4, 1)
    rt = torch.tensor(gold["opt"]["runtime"])
    t = ck.truncate_optimizer_states_for_checkpoint(opt, 8, rt)
    assert t.tolist() == gold["opt"]["truncated"]
    assert ck.pad_optimizer_states_from_checkpoint(opt, 8, t, 0.5, torch.float32, "cpu").tolist() == gold["opt"]["padded_init_0.5"]
    for ty, dim, want in ((EmbOptimType.SGD, 8, 0), (EmbOptimType.EXACT_ADAGRAD, 8, 8), (EmbOptimType.ADAM, 8, 16)):
        assert SparseOptimizer(ty, OptimizerArgs()).get_ckpt_state_dim(dim) == want


def test_optimizer_meta_round_trip():
    """get_opt_args / set_opt_args: the keys the reference's optimizers put into `<table>_opt_args.json` (optimizer.py:248-500)."""
    from dynamicemb.optimizer import OptimizerArgs, SparseOptimizer
    from dynamicemb.types import EmbOptimType
    want_keys = {EmbOptimType.SGD: {"opt_type", "lr"},
                 EmbOptimType.ADAM: {"opt_type", "lr", "iters", "beta1", "beta2", "eps", "weight_decay"},
                 EmbOptimType.EXACT_ADAGRAD: {"opt_type", "lr", "eps", "initial_accumulator_value"},
                 EmbOptimType.EXACT_ROWWISE_ADAGRAD: {"opt_type", "lr", "eps", "initial_accumulator_value"}}
    names = {EmbOptimType.SGD: "sgd", EmbOptimType.ADAM: "adam", EmbOptimType.EXACT_ADAGRAD: "exact_adagrad",
             EmbOptimType.EXACT_ROWWISE_ADAGRAD: "exact_row_wise_adagrad"}
    for ty, keys in want_keys.items():
        a = SparseOptimizer(ty, OptimizerArgs(learning_rate=0.3, eps=1e-6, initial_accumulator_value=0.25, beta1=0.8, beta2=0.9, weight_decay=0.01))
        a.iter = 17
        meta = a.get_opt_args()
        assert set(meta) == keys and meta["opt_type"] == names[ty]
        b = SparseOptimizer(ty, OptimizerArgs())
        b.set_opt_args(json.loads(json.dumps(meta)))
        assert b.get_opt_args() == meta
        with pytest.raises(ValueError):
            b.set_opt_args({"opt_type": names[ty]})


def test_reader_and_writer_bytes_match_reference(gold, tmp_path):
    """Our writer produces the reference's bytes, and our reader returns what the reference's `_iter_batches_from_files` returned."""
    from dynamicemb import checkpoint as ck
    R = gold["reader"]
    keys = torch.tensor(R["keys"], dtype=torch.int64)
    emb = torch.tensor(R["emb"], dtype=torch.float32)
    sc1 = torch.tensor(R["scores1"], dtype=torch.int64)
    sc2 = torch.tensor(R["scores2"], dtype=torch.int64)
    opt = torch.tensor(R["opt"], dtype=torch.float32)
    p = {k: str(tmp_path / k) for k in ("keys", "emb", "scores1", "scores2", "opt")}
    with ck.TableFileWriter(p["keys"], p["emb"], p["scores1"], p["opt"]) as w:       # two appends: the batching must not show in the bytes
        w.write(keys[:5], emb[:5], sc1[:5], opt[:5])
        w.write(keys[5:], emb[5:], sc1[5:], opt[5:])
    with ck.TableFileWriter(str(tmp_path / "k2"), str(tmp_path / "e2"), p["scores2"], None) as w:
        w.write(keys, emb, sc2, None)
    for name in ("keys", "emb", "scores1", "scores2", "opt"):
        assert open(p[name], "rb").read().hex() == R["bytes"][name], name
    assert not os.path.exists(str(tmp_path / "opt2"))

    def as_lists(batches):
        return [[x.tolist() if x is not None else None for x in b] for b in batches]
    got1 = list(ck.iter_batches_from_files(p["keys"], p["emb"], p["scores1"], p["opt"], 4, 4, "cpu", batch_size=3, num_scores=1))
    assert as_lists(got1) == R["read1"]
    got2 = list(ck.iter_batches_from_files(p["keys"], p["emb"], p["scores2"], p["opt"], 4, 4, "cpu", batch_size=3, num_scores=2))
    assert as_lists(got2) == R["read2"]
    got3 = list(ck.iter_batches_from_files(p["keys"], p["emb"], None, None, 4, 0, "cpu", batch_size=5))
    assert as_lists(got3) == R["read_no_scores_no_opt"]
    # re-sharded read: every key goes to exactly one of W ranks, the one its dist_type names
    for dist_type in ("roundrobin", "hash_roundrobin"):
        seen = []
        for r in range(3):
            for k, e, s, o in ck.iter_batches_from_files(p["keys"], p["emb"], p["scores1"], p["opt"], 4, 4, "cpu", batch_size=3, rank=r, world_size=3,
                                                         dist_type=dist_type):
                assert (ck.owner_rank(k, 3, dist_type) == r).all() and e.shape == (k.numel(), 4) and s.shape == k.shape and o.shape == (k.numel(), 4)
                seen += k.tolist()
        assert sorted(seen) == sorted(R["keys"])


def test_owner_rank_matches_bucketize_oracle():
    """The re-sharding owner rule equals the routing rule of the input dist (oracle restatement of sparse_block_bucketize_features.cu)."""
    from dynamicemb import checkpoint as ck
    from dynamicemb.scored_hashtable import murmur3_hash_64bits
    rng = np.random.default_rng(3)
    keys = np.concatenate([rng.integers(-(1 << 63), (1 << 63) - 1, size=2000, dtype=np.int64), np.array([0, 1, -1, -(1 << 63), (1 << 63) - 1], dtype=np.int64)])
    kt = torch.from_numpy(keys)
    for W in (2, 3, 5, 8):
        u = keys.view(np.uint64)
        assert ck.owner_rank(kt, W, "roundrobin").tolist() == (u % np.uint64(W)).astype(np.int64).tolist()
        assert ck.owner_rank(kt, W, "hash_roundrobin").tolist() == [murmur3_hash_64bits(int(k)) % W for k in u.tolist()]
    with pytest.raises(NotImplementedError):
        ck.owner_rank(kt, 2, "continuous")


def test_validate_load_meta_errors(tmp_path):
    from dynamicemb import checkpoint as ck
    from dynamicemb.optimizer import OptimizerArgs, SparseOptimizer
    from dynamicemb.types import EmbOptimType
    n, dim = 5, 4
    kp, vp, sp, op_ = (str(tmp_path / x) for x in ("k", "v", "s", "o"))
    open(kp, "wb").write(np.arange(n, dtype=np.int64).tobytes())
    open(vp, "wb").write(np.zeros((n, dim), np.float32).tobytes())
    open(sp, "wb").write(np.zeros(n, np.int64).tobytes())
    open(op_, "wb").write(np.zeros((n, dim), np.float32).tobytes())
    opt = SparseOptimizer(EmbOptimType.EXACT_ADAGRAD, OptimizerArgs())
    meta = dict(opt.get_opt_args(), evict_strategy="EvictStrategy.KLru", dist_type="roundrobin")
    assert ck.validate_load_meta(meta, opt, "EvictStrategy.KLru", "roundrobin", dim, 1, kp, vp, sp, op_, True) == (True, dim, n)
    assert ck.validate_load_meta(meta, opt, "EvictStrategy.KLru", "roundrobin", dim, 1, kp, vp, sp, None, True)[0] is False
    sgd = SparseOptimizer(EmbOptimType.SGD, OptimizerArgs())
    assert ck.validate_load_meta(meta, sgd, "EvictStrategy.KLru", "roundrobin", dim, 1, kp, vp, sp, op_, True)[0] is False   # optimizer type differs
    with pytest.raises(ValueError, match="Evict strategy mismatch"):
        ck.validate_load_meta(meta, opt, "EvictStrategy.KLfu", "roundrobin", dim, 1, kp, vp, sp, op_, True)
    with pytest.raises(ValueError, match="dist_type mismatch"):
        ck.validate_load_meta(meta, opt, "EvictStrategy.KLru", "hash_roundrobin", dim, 1, kp, vp, sp, op_, True)
    with pytest.raises(ValueError, match="number of scores"):
        ck.validate_load_meta(meta, opt, "EvictStrategy.KLru", "roundrobin", dim, 2, kp, vp, sp, op_, True)
    with pytest.raises(ValueError, match="number of embeddings"):
        ck.validate_load_meta(meta, opt, "EvictStrategy.KLru", "roundrobin", dim + 1, 1, kp, vp, sp, op_, True)
    open(op_, "wb").write(np.zeros((n, 1), np.float32).tobytes())
    with pytest.raises(ValueError, match="Optimizer state width"):
        ck.validate_load_meta(meta, opt, "EvictStrategy.KLru", "roundrobin", dim, 1, kp, vp, sp, op_, True)


def test_score_block_view_follows_table_layout():
    """gather / scatter of whole score blocks address `scores[C][ns]` behind keys and digests of each bucket (csrc/demb_common.cuh)."""
    from dynamicemb.scored_hashtable import LinearBucketTable
    C, ns, nb = 16, 2, 5
    bucket_bytes = (9 + 8 * ns) * C
    img = np.zeros(nb * bucket_bytes, dtype=np.uint8)
    want = np.zeros((nb, C, ns), dtype=np.int64)
    for b in range(nb):
        sc = img[b * bucket_bytes + 9 * C:(b + 1) * bucket_bytes].view(np.int64).reshape(C, ns)
        sc[:] = (np.arange(C * ns).reshape(C, ns) + 1000 * b)
        want[b] = sc
    tb = object.__new__(LinearBucketTable)
    tb.bucket_capacity_, tb.num_scores_, tb.num_buckets_ = C, ns, nb
    tb.table_storage_ = torch.from_numpy(img)
    tb.table_bucket_offsets_cpu_ = torch.tensor([0, 2, 5])
    assert torch.equal(tb._score_words(), torch.from_numpy(want))
    idx = torch.tensor([0, 17, 47, 5])                      # table 1 starts at bucket 2
    got = tb.gather_score_blocks(1, idx)
    assert got.tolist() == [want[2 + i // C, i % C].tolist() for i in idx.tolist()]
    new = torch.tensor([[7, 8], [9, 10], [11, 12], [13, 14]])
    tb.scatter_score_blocks(1, torch.tensor([3, -1, 20, 47]), new)          # -1 = failed insert: skipped
    after = tb._score_words()
    assert after[2, 3].tolist() == [7, 8] and after[3, 4].tolist() == [11, 12] and after[4, 15].tolist() == [13, 14]
    keys_untouched = torch.from_numpy(img).view(torch.uint8)[:8 * C]
    assert int(keys_untouched.sum()) == 0
    with pytest.raises(ValueError):
        tb.scatter_score_blocks(0, torch.tensor([1, 2]), torch.zeros(2, dtype=torch.int64))
