"""Golden vectors produced by the reference's OWN CUDA kernels on a B200 (tests/golden/gen_golden_gpu.py drives the unmodified
`dynamicemb_extensions` build the way scored_hashtable.py does).  The CPU half pins the oracle (oracle/) to the reference; the GPU
half replays the same inputs through the product's C ABI.  Integer work is compared bit-exactly (table image bytes, slot indices,
evicted sets, bucketized ids); fp32 row math at 1e-5 (the reference build uses --use_fast_math, corelib/dynamicemb/setup.py:121)."""
import os

import numpy as np
import pytest

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TABLES = ["table_single_c128", "table_multi_c64", "table_one_bucket_c16", "table_c1024"]


def _load(name):
    return np.load(os.path.join(G, name + ".npz"))


def _evset(keys, idx, scores):
    return sorted(zip(np.asarray(keys).view(np.uint64).tolist(), np.asarray(idx).tolist(), np.asarray(scores).tolist()))


@pytest.mark.parametrize("name", TABLES)
def test_oracle_table_matches_reference_kernels(name):
    from oracle.dynamicemb import OracleTable
    z = _load(name)
    o = OracleTable([int(c) for c in z["caps"]], bucket_capacity=int(z["C"]))
    for s in range(int(z["nsteps"])):
        keys, tids, scores = z[f"s{s}_keys"], z[f"s{s}_tids"], z[f"s{s}_scores"]
        _, _, _, (ek, ei, es, _) = o.insert(keys, tids, policy=1, score_in=scores)
        _, _, idx = o.lookup(keys, tids, policy=0)                       # final indices, as _deterministic_insert returns them
        assert np.array_equal(idx, z[f"s{s}_indices"]), f"{name} step {s}: indices"
        assert np.array_equal(o.storage, z[f"s{s}_image"]), f"{name} step {s}: table image"
        assert np.array_equal(o.bucket_sizes, z[f"s{s}_bucket_sizes"]), f"{name} step {s}: bucket sizes"
        if f"s{s}_ev_keys" in z:
            assert _evset(ek, ei, es) == _evset(z[f"s{s}_ev_keys"], z[f"s{s}_ev_idx"], z[f"s{s}_ev_scores"]), f"{name} step {s}: evicted"
        else:
            assert ek.size == 0
        if f"s{s}_erase_keys" in z:
            o.erase(z[f"s{s}_erase_keys"], z[f"s{s}_erase_tids"])
            assert np.array_equal(o.storage, z[f"s{s}_image_after_erase"]), f"{name} step {s}: erase"
        so, fo, io = o.lookup(z[f"s{s}_q_keys"], z[f"s{s}_q_tids"], policy=1, score_in=z[f"s{s}_q_scores"])
        assert np.array_equal(fo, z[f"s{s}_q_founds"]) and np.array_equal(io, z[f"s{s}_q_indices"]) and np.array_equal(so, z[f"s{s}_q_score_out"])
        assert np.array_equal(o.storage, z[f"s{s}_image_after_lookup"]), f"{name} step {s}: lookup score update"


def test_oracle_rows_match_reference_kernels():
    from oracle import dynamicemb as orc
    z = _load("rows")
    D, B, F = int(z["D"]), int(z["B"]), int(z["F"])
    inv, offs, ue = z["inverse"], z["offsets"], z["unique_embs"]
    assert np.array_equal(orc.gather_rows(ue, D, inv), z["seq"])
    n = inv.size
    bag = np.searchsorted(offs, np.arange(n), side="right") - 1
    f, b = bag // B, bag % B
    grow = b * F + f
    blen = (offs[1:] - offs[:-1])[bag]
    for comb in (0, 1):
        np.testing.assert_allclose(orc.pool_rows(ue, D, offs, inv, comb, B, F), z[f"pooled_{comb}"], rtol=1e-5, atol=1e-5)
        scale = (np.float32(1) / blen.astype(np.float32)) if comb == 1 else np.ones(n, dtype=np.float32)
        ug = orc.reduce_grads(inv, z[f"grads_{comb}"].reshape(B * F, D)[grow], scale, ue.shape[0], D)
        np.testing.assert_allclose(ug, z[f"ugrads_{comb}"], rtol=1e-4, atol=1e-4)          # fp32 sums of up to ~1e3 terms, order differs
    rows = np.arange(z["opt_rows"].size)
    for nm, opt in [("sgd", "sgd"), ("adagrad", "adagrad"), ("adam", "adam"), ("rowwise", "rowwise_adagrad")]:
        vals = z[f"opt_{nm}_before"].copy()
        orc.optimizer_update(vals, D, rows, z["opt_grads"], opt, lr=0.05, eps=1e-8, beta1=0.9, beta2=0.999, weight_decay=0.01, step=3)
        np.testing.assert_allclose(vals, z[f"opt_{nm}_after"], rtol=2e-5, atol=2e-6, err_msg=nm)


def test_oracle_unique_and_bucketize_match_reference_kernels():
    from oracle import dynamicemb as orc
    z = _load("unique_bucketize")
    keys, trange = z["keys"], z["trange"]
    uk, inv, offs = orc.segmented_unique(keys, trange)
    assert uk.size == int(z["num_unique"]) and np.array_equal(offs, z["table_offsets"])
    ruk, rrev = z["unique_keys"], z["reverse"]
    assert np.array_equal(ruk[rrev], keys)                                   # reference inverse reconstructs the ids
    for t in range(trange.size - 1):                                         # same per-table key sets (reference order is racy)
        assert set(uk[offs[t]:offs[t + 1]].tolist()) == set(ruk[offs[t]:offs[t + 1]].tolist())
    for tag in ("cont", "rr", "hash", "mixed"):
        nl, ni, perm = orc.block_bucketize(z["bk_lengths"], z["bk_ids"], int(z["bk_B"]), int(z["bk_W"]), z["bk_blk"], z[f"bk_{tag}_dts"])
        assert np.array_equal(nl, z[f"bk_{tag}_new_lengths"]), tag
        assert np.array_equal(ni, z[f"bk_{tag}_new_ids"]), tag
        assert np.array_equal(perm, z[f"bk_{tag}_perm"]), tag


# --------------------------------------------------------------------------------------------------- GPU half
@pytest.mark.gpu
@pytest.mark.parametrize("name", TABLES)
def test_cuda_table_matches_reference_kernels(cuda, name):
    import torch
    from dynamicemb.scored_hashtable import LinearBucketTable, ScoreArg, ScoreSpec
    from dynamicemb.dynamicemb_extensions import ScorePolicy
    z = _load(name)
    t = LinearBucketTable([int(c) for c in z["caps"]], [ScoreSpec("s", ScorePolicy.ASSIGN)], bucket_capacity=int(z["C"]), device=cuda)
    dv = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(cuda)
    for s in range(int(z["nsteps"])):
        keys, tids, scores = dv(z[f"s{s}_keys"]), dv(z[f"s{s}_tids"]), dv(z[f"s{s}_scores"])
        _, nev, ek, ei, es, _ = t.insert_and_evict(keys, tids, ScoreArg("s", scores, ScorePolicy.ASSIGN))
        _, _, idx = t.lookup(keys, tids, ScoreArg("s", None, ScorePolicy.CONST))
        assert np.array_equal(idx.cpu().numpy(), z[f"s{s}_indices"]), f"{name} step {s}: indices"
        assert np.array_equal(t.table_storage_.cpu().numpy(), z[f"s{s}_image"]), f"{name} step {s}: table image"
        assert np.array_equal(t.bucket_sizes.cpu().numpy(), z[f"s{s}_bucket_sizes"])
        if f"s{s}_ev_keys" in z:
            assert _evset(ek.cpu().numpy(), ei.cpu().numpy(), es.cpu().numpy()) == _evset(z[f"s{s}_ev_keys"], z[f"s{s}_ev_idx"], z[f"s{s}_ev_scores"])
        else:
            assert nev == 0
        if f"s{s}_erase_keys" in z:
            t.erase(dv(z[f"s{s}_erase_keys"]), dv(z[f"s{s}_erase_tids"]))
            assert np.array_equal(t.table_storage_.cpu().numpy(), z[f"s{s}_image_after_erase"])
        so, fo, io = t.lookup(dv(z[f"s{s}_q_keys"]), dv(z[f"s{s}_q_tids"]), ScoreArg("s", dv(z[f"s{s}_q_scores"]), ScorePolicy.ASSIGN))
        assert np.array_equal(io.cpu().numpy(), z[f"s{s}_q_indices"]) and np.array_equal(so.cpu().numpy(), z[f"s{s}_q_score_out"])
        assert np.array_equal(t.table_storage_.cpu().numpy(), z[f"s{s}_image_after_lookup"])


@pytest.mark.gpu
def test_cuda_rows_match_reference_kernels(cuda):
    import torch
    from dynamicemb import dynamicemb_extensions as ext
    z = _load("rows")
    D, B, F = int(z["D"]), int(z["B"]), int(z["F"])
    dv = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(cuda)
    inv, offs, ue = dv(z["inverse"]), dv(z["offsets"]), dv(z["unique_embs"])
    out = torch.empty(inv.numel(), D, device=cuda)
    ext.gather_embedding(ue, out, inv)
    assert np.array_equal(out.cpu().numpy(), z["seq"])
    for comb in (0, 1):
        out = torch.empty(B, F * D, device=cuda)
        ext.gather_embedding_pooled(ue, out, inv, offs, comb, F * D, B)
        np.testing.assert_allclose(out.cpu().numpy(), z[f"pooled_{comb}"], rtol=1e-5, atol=1e-5)
        ug = ext.reduce_grads(inv, dv(z[f"grads_{comb}"]), ue.shape[0], B, D, offs, None, comb, F * D)
        np.testing.assert_allclose(ug.cpu().numpy(), z[f"ugrads_{comb}"], rtol=1e-4, atol=1e-4)
    rows = torch.arange(z["opt_rows"].size, dtype=torch.int64, device=cuda)
    for nm, code in [("sgd", 1), ("adagrad", 3), ("adam", 2), ("rowwise", 4)]:
        vals = dv(z[f"opt_{nm}_before"])
        ext.update_rows(vals, D, rows, dv(z["opt_grads"]), code, lr=0.05, eps=1e-8, beta1=0.9, beta2=0.999, weight_decay=0.01,
                        bc1=1 - 0.9 ** 3, bc2=1 - 0.999 ** 3)
        np.testing.assert_allclose(vals.cpu().numpy(), z[f"opt_{nm}_after"], rtol=2e-5, atol=2e-6, err_msg=nm)


@pytest.mark.gpu
def test_cuda_bucketize_matches_reference_kernels(cuda):
    import torch
    from dynamicemb import dynamicemb_extensions as ext
    z = _load("unique_bucketize")
    dv = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(cuda)
    for tag in ("cont", "rr", "hash", "mixed"):
        nl, ni, _, perm = ext.block_bucketize_sparse_features(dv(z["bk_lengths"]), dv(z["bk_ids"]), int(z["bk_B"]), int(z["bk_W"]), dv(z["bk_blk"]),
                                                              dv(z[f"bk_{tag}_dts"]).to(torch.int32))
        assert np.array_equal(nl.cpu().numpy(), z[f"bk_{tag}_new_lengths"]), tag
        assert np.array_equal(ni.cpu().numpy(), z[f"bk_{tag}_new_ids"]), tag
        assert np.array_equal(perm.cpu().numpy(), z[f"bk_{tag}_perm"]), tag
