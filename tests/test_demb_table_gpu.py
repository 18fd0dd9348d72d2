"""GPU parity: hash table ops (C ABI through the op layer) vs the CPU oracle — bit-exact table image,
slot indices, InsertResult codes, scores.  Mirrors the reference's
corelib/dynamicemb/test/unit_tests/table_operation/test_table_operation.py:277-655 (insert / re-insert /
lookup / erase / reclaim / evict invariants) and test_batched_dynamic_embedding_tables_v2.py:1954-2057
(test_deterministic_insert: byte-identical `keys_` image)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _mk(cuda, caps, C=128, policy=1, ns=1):
    from dynamicemb.scored_hashtable import LinearBucketTable, ScoreSpec
    from dynamicemb.dynamicemb_extensions import ScorePolicy
    from oracle.dynamicemb import OracleTable
    t = LinearBucketTable(caps, [ScoreSpec(name="s", policy=ScorePolicy(policy))], bucket_capacity=C, device=cuda)
    o = OracleTable(caps, bucket_capacity=C, num_scores=ns)
    return t, o


def _img(t):
    return t.table_storage_.cpu().numpy()


def _rand_keys(rng, n, lo=-(1 << 62), hi=(1 << 62)):
    k = rng.integers(lo, hi, size=3 * n, dtype=np.int64)
    k = np.unique(k)
    rng.shuffle(k)
    return k[:n]


def _same_image(t, o, what):
    a, b = _img(t), o.storage
    if not np.array_equal(a, b):
        bad = np.nonzero(a != b)[0]
        raise AssertionError(f"{what}: table image differs at {bad.size} bytes, first {bad[:8]}")
    assert np.array_equal(t.bucket_sizes.cpu().numpy(), o.bucket_sizes), what + ": bucket_sizes"


@pytest.mark.parametrize("C,nb,n", [(128, 64, 3000), (128, 1, 100), (16, 7, 100), (1024, 4, 3500), (128, 512, 20000)])
def test_insert_lookup_bit_exact(cuda, C, nb, n):
    from dynamicemb.scored_hashtable import ScoreArg
    from dynamicemb.dynamicemb_extensions import ScorePolicy
    rng = np.random.default_rng(C * 1000 + nb)
    t, o = _mk(cuda, [C * nb], C=C)
    keys = _rand_keys(rng, n)
    scores = rng.integers(1, 1 << 40, size=n, dtype=np.int64)
    kt = torch.from_numpy(keys).to(cuda)
    tid = torch.zeros(n, dtype=torch.int64, device=cuda)
    res = torch.empty(n, dtype=torch.uint8, device=cuda)
    so = torch.empty(n, dtype=torch.int64, device=cuda)
    idx = t.insert(kt, tid, ScoreArg("s", torch.from_numpy(scores).to(cuda), ScorePolicy.ASSIGN), insert_results=res, score_out=so)
    oidx, ores, oso, _ = o.insert(keys, None, policy=1, score_in=scores)
    assert np.array_equal(idx.cpu().numpy(), oidx)
    assert np.array_equal(res.cpu().numpy(), ores)
    assert np.array_equal(so.cpu().numpy(), oso)
    _same_image(t, o, "after insert")
    # lookup (CONST) returns the same slots and stored scores; unknown keys miss
    probe = np.concatenate([keys[: n // 2], _rand_keys(rng, 50, 1 << 62, (1 << 63) - 8)])
    s, f, i = t.lookup(torch.from_numpy(probe).to(cuda), torch.zeros(probe.size, dtype=torch.int64, device=cuda), ScoreArg("s", None, ScorePolicy.CONST))
    os_, of, oi = o.lookup(probe, None, policy=0)
    assert np.array_equal(i.cpu().numpy(), oi) and np.array_equal(f.cpu().numpy(), of) and np.array_equal(s.cpu().numpy(), os_)
    # re-insert => ASSIGN with identical indices (test_table_operation.py:277-526)
    idx2 = t.insert(kt, tid, ScoreArg("s", torch.from_numpy(scores + 1).to(cuda), ScorePolicy.ASSIGN), insert_results=res)
    oidx2, ores2, _, _ = o.insert(keys, None, policy=1, score_in=scores + 1)
    ok = oidx >= 0
    assert np.array_equal(idx2.cpu().numpy(), oidx2) and np.array_equal(res.cpu().numpy(), ores2)
    assert set(ores2[ok & (oidx2 == oidx)].tolist()) <= {0, 2, 3}   # ASSIGN for resident keys; keys evicted by a full bucket re-enter as INSERT/EVICT
    _same_image(t, o, "after re-insert")


def test_evict_erase_reclaim_sequence(cuda):
    """Fill past capacity (EVICT / min-score-first-wins), erase (RECLAIM sentinel), reinsert (RECLAIM result)."""
    from dynamicemb.scored_hashtable import ScoreArg
    from dynamicemb.dynamicemb_extensions import ScorePolicy
    rng = np.random.default_rng(7)
    C, nb = 64, 16
    t, o = _mk(cuda, [C * nb], C=C)
    tid = lambda n: torch.zeros(n, dtype=torch.int64, device=cuda)
    for step in range(12):
        n = 400
        keys = _rand_keys(rng, n)
        sc = np.full(n, step + 1, dtype=np.int64) if step % 3 else rng.integers(1, 50, size=n, dtype=np.int64)
        res = torch.empty(n, dtype=torch.uint8, device=cuda)
        idx, nev, ek, ei, es, et = t.insert_and_evict(torch.from_numpy(keys).to(cuda), tid(n), ScoreArg("s", torch.from_numpy(sc).to(cuda), ScorePolicy.ASSIGN), insert_results=res)
        oidx, ores, _, (oek, oei, oes, oet) = o.insert(keys, None, policy=1, score_in=sc)
        assert np.array_equal(idx.cpu().numpy(), oidx), f"step {step}"
        assert np.array_equal(res.cpu().numpy(), ores), f"step {step}"
        assert nev == oek.size
        g = sorted(zip(ek.cpu().numpy().view(np.uint64).tolist(), ei.cpu().numpy().tolist(), es.cpu().numpy().tolist()))
        w = sorted(zip(oek.tolist(), oei.tolist(), oes.tolist()))
        assert g == w, f"evicted records differ at step {step}"
        _same_image(t, o, f"step {step}")
        if step % 4 == 3:   # erase a subset of what is in the table, then it must be reclaimed later
            present = keys[oidx >= 0][:100]
            t.erase(torch.from_numpy(present).to(cuda), tid(present.size))
            o.erase(present)
            _same_image(t, o, f"erase {step}")
    assert (np.asarray(_img(t)) == o.storage).all()


def test_pinned_rows_are_not_evicted(cuda):
    from dynamicemb.scored_hashtable import ScoreArg
    from dynamicemb.dynamicemb_extensions import ScorePolicy
    rng = np.random.default_rng(11)
    C = 32
    t, o = _mk(cuda, [C], C=C)
    k0 = _rand_keys(rng, C)
    z = torch.zeros(C, dtype=torch.int64, device=cuda)
    idx = t.insert(torch.from_numpy(k0).to(cuda), z, ScoreArg("s", torch.arange(1, C + 1, device=cuda), ScorePolicy.ASSIGN))
    oidx, _, _, _ = o.insert(k0, None, policy=1, score_in=np.arange(1, C + 1))
    # pin every slot except 3
    pin = idx[idx != idx[3]]
    t.increment_counter(pin, torch.zeros_like(pin))
    o.counter[oidx[oidx != oidx[3]]] += 1
    k1 = _rand_keys(rng, 5)
    res = torch.empty(5, dtype=torch.uint8, device=cuda)
    i1 = t.insert(torch.from_numpy(k1).to(cuda), z[:5], ScoreArg("s", torch.full((5,), 99, device=cuda), ScorePolicy.ASSIGN), insert_results=res)
    oi1, ores, _, _ = o.insert(k1, None, policy=1, score_in=np.full(5, 99))
    assert np.array_equal(i1.cpu().numpy(), oi1) and np.array_equal(res.cpu().numpy(), ores)
    assert set(ores.tolist()) <= {3, 5}  # only the one unpinned slot can be taken (EVICT) or the insert is BUSY
    _same_image(t, o, "pinned")


def test_multi_table_and_policies(cuda):
    from dynamicemb.scored_hashtable import ScoreArg
    from dynamicemb.dynamicemb_extensions import ScorePolicy
    rng = np.random.default_rng(3)
    caps = [128 * 8, 128 * 3, 128 * 20]
    for pol in (1, 2):
        t, o = _mk(cuda, caps, C=128, policy=pol)
        for step in range(4):
            n = 1500
            keys = _rand_keys(rng, n)
            tids = rng.integers(0, 3, size=n).astype(np.int64)
            sc = rng.integers(1, 9, size=n, dtype=np.int64)
            idx = t.insert(torch.from_numpy(keys).to(cuda), torch.from_numpy(tids).to(cuda), ScoreArg("s", torch.from_numpy(sc).to(cuda), ScorePolicy(pol)))
            oidx, _, _, _ = o.insert(keys, tids, policy=pol, score_in=sc)
            assert np.array_equal(idx.cpu().numpy(), oidx)
            # lookups with the table's policy mutate scores of hits (key_value_table.py:811-864)
            sub = rng.permutation(n)[:700]
            s, f, i = t.lookup(torch.from_numpy(keys[sub]).to(cuda), torch.from_numpy(tids[sub]).to(cuda), ScoreArg("s", torch.from_numpy(sc[sub]).to(cuda), ScorePolicy(pol)))
            os_, of, oi = o.lookup(keys[sub], tids[sub], policy=pol, score_in=sc[sub])
            assert np.array_equal(i.cpu().numpy(), oi) and np.array_equal(s.cpu().numpy(), os_)
            _same_image(t, o, f"policy {pol} step {step}")


def test_illegal_and_empty(cuda):
    from dynamicemb.scored_hashtable import ScoreArg
    from dynamicemb.dynamicemb_extensions import ScorePolicy
    t, o = _mk(cuda, [256], C=128)
    bad = np.array([-1, -2, -3, -4, 5], dtype=np.int64)     # top 62 bits set => reserved (types.cuh:144-146)
    res = torch.empty(5, dtype=torch.uint8, device=cuda)
    idx = t.insert(torch.from_numpy(bad).to(cuda), torch.zeros(5, dtype=torch.int64, device=cuda), ScoreArg("s", torch.ones(5, dtype=torch.int64, device=cuda), ScorePolicy.ASSIGN), insert_results=res)
    oidx, ores, _, _ = o.insert(bad, None, policy=1, score_in=np.ones(5, dtype=np.int64))
    assert np.array_equal(idx.cpu().numpy(), oidx) and np.array_equal(res.cpu().numpy(), ores)
    assert list(ores[:4]) == [6, 6, 6, 6] and ores[4] == 0
    e = torch.empty(0, dtype=torch.int64, device=cuda)
    assert t.insert(e, e, ScoreArg("s", e, ScorePolicy.ASSIGN)).numel() == 0
    s, f, i = t.lookup(e, e, ScoreArg("s", None, ScorePolicy.CONST))
    assert i.numel() == 0
    _same_image(t, o, "illegal")


def test_deterministic_insert_is_order_independent(cuda):
    """test_deterministic_insert (reference :1954-2057): same key set, different presentation order => identical keys_ image."""
    from dynamicemb.scored_hashtable import ScoreArg
    from dynamicemb.dynamicemb_extensions import ScorePolicy
    rng = np.random.default_rng(5)
    ta, _ = _mk(cuda, [128 * 32], C=128)
    tb, _ = _mk(cuda, [128 * 32], C=128)
    for it in range(6):
        keys = _rand_keys(rng, 2048)
        perm = rng.permutation(keys.size)
        sc = torch.full((keys.size,), it + 1, dtype=torch.int64, device=cuda)
        z = torch.zeros(keys.size, dtype=torch.int64, device=cuda)
        ia = ta.insert(torch.from_numpy(keys).to(cuda), z, ScoreArg("s", sc, ScorePolicy.ASSIGN))
        ib = tb.insert(torch.from_numpy(keys[perm]).to(cuda), z, ScoreArg("s", sc, ScorePolicy.ASSIGN))
        assert torch.equal(ta.keys_, tb.keys_)
        assert torch.equal(ia[torch.from_numpy(perm).to(cuda)], ib)
