"""GPU parity: dedup, fused forward (sequence + pooled), row init, fused backward + optimizers, bucketize —
against the CPU oracle.  Copies and id/index work are bit-exact; SUM pooling and gradient reduction are
bit-exact too because the kernels fix the summation order the oracle restates; optimizer math is
compared at 1e-6 (the reference's own tolerance vs FBGEMM TBE, test_batched_dynamic_embedding_tables_v2.py:1594-1745)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _zipf_ids(rng, n, space, alpha=1.05):
    r = rng.zipf(alpha, size=n).astype(np.int64)
    return (r % space)


def test_segmented_unique_first_occurrence(cuda):
    from dynamicemb import dynamicemb_extensions as ext
    from oracle import dynamicemb as orc
    rng = np.random.default_rng(0)
    for n, T in [(1, 1), (1000, 1), (50000, 3), (4096, 4)]:
        keys = _zipf_ids(rng, n, 5000) - 7
        if n > 10:
            keys[5] = -1   # the ~0 key takes the reserved scratch slot
            keys[n - 1] = -1
        cuts = np.sort(rng.integers(0, n + 1, size=T - 1))
        trange = np.concatenate([[0], cuts, [n]]).astype(np.int64)
        nu, uk, rev, toffs, freq, utids = ext.segmented_unique_cuda(torch.from_numpy(keys).to(cuda), torch.from_numpy(trange).to(cuda), T,
                                                                     torch.empty(0, dtype=torch.int64, device=cuda), want_table_ids=True)
        ouk, oinv, ooffs = orc.segmented_unique(keys, trange)
        k = int(nu.item())
        assert k == ouk.size
        assert np.array_equal(uk[:k].cpu().numpy(), ouk)
        assert np.array_equal(rev.cpu().numpy(), oinv)
        assert np.array_equal(toffs.cpu().numpy(), ooffs)
        cnt = np.bincount(oinv, minlength=k)
        assert np.array_equal(freq[:k].cpu().numpy(), cnt)
        exp_tid = np.searchsorted(ooffs, np.arange(k), side="right") - 1
        assert np.array_equal(utids[:k].cpu().numpy(), exp_tid)
        assert np.array_equal(ext.expand_table_ids_cuda(toffs, k).cpu().numpy(), exp_tid)


def _filled_table(cuda, rng, caps, D, state, n_keys, C=128):
    from dynamicemb.scored_hashtable import LinearBucketTable, ScoreSpec, ScoreArg
    from dynamicemb.dynamicemb_extensions import ScorePolicy
    t = LinearBucketTable(caps, [ScoreSpec("s", ScorePolicy.ASSIGN)], bucket_capacity=C, device=cuda)
    vdim = D + state
    values = torch.randn(t.capacity_, vdim, device=cuda)
    T = len(caps)
    keys = np.unique(rng.integers(0, 1 << 40, size=2 * n_keys, dtype=np.int64))[:n_keys]
    tids = rng.integers(0, T, size=keys.size).astype(np.int64)
    kt, tt = torch.from_numpy(keys).to(cuda), torch.from_numpy(tids).to(cuda)
    t.insert(kt, tt, ScoreArg("s", torch.ones(keys.size, dtype=torch.int64, device=cuda)))
    _, _, idx = t.lookup(kt, tt, ScoreArg("s", None, ScorePolicy.CONST))     # final slots (a full bucket may have evicted early keys)
    return t, values, keys, tids, idx.cpu().numpy()


@pytest.mark.parametrize("D,out_dtype", [(128, torch.float32), (64, torch.float32), (256, torch.float32), (32, torch.bfloat16)])
def test_fused_lookup_forward_sequence(cuda, D, out_dtype):
    from dynamicemb import dynamicemb_extensions as ext
    rng = np.random.default_rng(D)
    caps = [128 * 40, 128 * 24]
    t, values, keys, tids, slots = _filled_table(cuda, rng, caps, D, 4, 5000)
    # ids grouped by table (feature-major KJT), with ~10% unknown ids
    ids, id_t = [], []
    for tb in range(2):
        mine = keys[(tids == tb) & (slots >= 0)]
        pick = mine[rng.integers(0, mine.size, size=3000)]
        unk = rng.integers(1 << 41, 1 << 42, size=300, dtype=np.int64)
        a = np.concatenate([pick, unk]); rng.shuffle(a)
        ids.append(a); id_t.append(np.full(a.size, tb))
    trange = np.array([0, ids[0].size, ids[0].size + ids[1].size], dtype=np.int64)
    ids = np.concatenate(ids); id_t = np.concatenate(id_t)
    out, founds, got_slots = ext.lookup_forward(t.table_storage_, t.table_bucket_offsets_, t.bucket_capacity_, values, D, torch.from_numpy(ids).to(cuda),
                                                row_base=t.row_base_, table_range=torch.from_numpy(trange).to(cuda), num_tables=2,
                                                out_dtype=out_dtype, absent_value=0.0, want_founds=True)
    slot_of = {(int(k), int(tt)): int(s) for k, tt, s in zip(keys, tids, slots) if s >= 0}
    exp_slot = np.array([slot_of.get((int(k), int(tt)), -1) for k, tt in zip(ids, id_t)])
    assert np.array_equal(got_slots.cpu().numpy(), exp_slot)
    assert np.array_equal(founds.cpu().numpy(), exp_slot >= 0)
    base = t.row_base_.cpu().numpy()
    rows = np.where(exp_slot >= 0, base[id_t] + exp_slot, -1)
    v = values.cpu()
    exp = torch.zeros(ids.size, D)
    m = torch.from_numpy(rows >= 0)
    exp[m] = v[torch.from_numpy(rows[rows >= 0])][:, :D]
    assert torch.equal(out.cpu(), exp.to(out_dtype))       # pure copy (+ one rounding for bf16): bit-exact


@pytest.mark.parametrize("n_per_table,n_keys", [(3300, 5000), (17, 5000), (600000, 5000), (3300, 8150)])   # 600000: > 2 ring revolutions per CTA in the
def test_fused_lookup_kernel_variants_agree(cuda, n_per_table, n_keys):                                        # specialised layouts; 8150 keys: buckets (nearly) full
    """demb_set_option(0, v): every fused-lookup kernel (round-1 thread-per-key probe, per-warp pipeline = default, the specialised
    probe / copy layouts) must return the same rows, founds and slots; a ragged tail (n % 32 != 0), two tables, absent ids."""
    from dynamicemb import dynamicemb_extensions as ext
    from dynamicemb import _native as N
    D = 128
    rng = np.random.default_rng(n_per_table)
    t, values, keys, tids, slots = _filled_table(cuda, rng, [128 * 40, 128 * 24], D, 4, n_keys)
    ids = []
    for tb in range(2):
        mine = keys[(tids == tb) & (slots >= 0)]
        a = np.concatenate([mine[rng.integers(0, mine.size, size=n_per_table)], rng.integers(1 << 41, 1 << 42, size=n_per_table // 10 + 1, dtype=np.int64)])
        rng.shuffle(a); ids.append(a)
    trange = torch.tensor([0, ids[0].size, ids[0].size + ids[1].size], dtype=torch.int64, device=cuda)
    idt = torch.from_numpy(np.concatenate(ids)).to(cuda)
    res = {}
    try:
        for v in (1, 0, 2, 3, 4, 5):
            N.lib.demb_set_option(0, v)
            out, founds, sl = ext.lookup_forward(t.table_storage_, t.table_bucket_offsets_, t.bucket_capacity_, values, D, idt, row_base=t.row_base_,
                                                 table_range=trange, num_tables=2, out_dtype=torch.float32, absent_value=0.25, want_founds=True)
            torch.cuda.synchronize()
            res[v] = (out.clone(), founds.clone(), sl.clone())
    finally:
        N.lib.demb_set_option(0, 1)
    assert int(res[1][1].sum()) >= 2 * n_per_table
    for v in (0, 2, 3, 4, 5):
        assert torch.equal(res[v][2], res[1][2]), v
        assert torch.equal(res[v][1], res[1][1]), v
        assert torch.equal(res[v][0], res[1][0]), v


@pytest.mark.parametrize("combiner", [0, 1])
@pytest.mark.parametrize("D", [128, 64])
def test_fused_lookup_forward_pooled(cuda, combiner, D):
    from dynamicemb import dynamicemb_extensions as ext
    from oracle import dynamicemb as orc
    rng = np.random.default_rng(17 + D + combiner)
    t, values, keys, tids, slots = _filled_table(cuda, rng, [128 * 64], D, 0, 6000)
    ok = keys[slots >= 0]
    B, F = 37, 3
    lens = rng.integers(0, 45, size=F * B)          # empty bags and bags longer than a warp
    offsets = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    ids = ok[rng.integers(0, ok.size, size=int(offsets[-1]))]
    ids[rng.integers(0, ids.size, size=20)] = 1 << 45   # a few absent ids
    out = ext.lookup_forward(t.table_storage_, t.table_bucket_offsets_, t.bucket_capacity_, values, D, torch.from_numpy(ids).to(cuda),
                             row_base=t.row_base_, offsets=torch.from_numpy(offsets).to(cuda), batch_size=B, num_features=F, combiner=combiner)
    slot_of = dict(zip(keys.tolist(), slots.tolist()))
    sl = np.array([slot_of.get(int(k), -1) for k in ids])
    exp = orc.pool_rows(values.cpu().numpy(), D, offsets, sl, combiner, B, F)
    assert np.array_equal(out.cpu().numpy(), exp)          # same accumulation order => bit-exact


def test_init_rows_modes(cuda):
    from dynamicemb import dynamicemb_extensions as ext
    D, vdim, n = 128, 256, 4096
    values = torch.full((n, vdim), 7.0, device=cuda)
    keys = torch.arange(1000, 1000 + n, dtype=torch.int64, device=cuda) * 7919
    rows = torch.arange(n, dtype=torch.int64, device=cuda).flip(0).contiguous()
    ext.init_rows(values, D, rows, keys, ext.InitializerMode.DEBUG, state_init=0.5)
    exp = (keys % 100000).to(torch.float32)
    assert torch.equal(values[rows][:, :D], exp[:, None].expand(n, D))      # initializer.cuh:142-156 debug = key % 100000
    assert torch.equal(values[:, D:], torch.full((n, vdim - D), 0.5, device=cuda))
    ext.init_rows(values, D, rows, keys, ext.InitializerMode.UNIFORM, -0.25, 0.75, seed=3)
    e = values[:, :D]
    assert e.min() >= -0.25 and e.max() <= 0.75 and abs(e.mean().item() - 0.25) < 5e-3 and abs(e.std().item() - (1 / 12) ** 0.5) < 5e-3
    first = e.clone()
    ext.init_rows(values, D, rows.flip(0).contiguous(), keys.flip(0).contiguous(), ext.InitializerMode.UNIFORM, -0.25, 0.75, seed=3)
    assert torch.equal(values[:, :D], first)     # value depends only on (seed, key, column)
    ext.init_rows(values, D, rows, keys, ext.InitializerMode.NORMAL, 1.0, 2.0, seed=9)
    e = values[:, :D]
    assert abs(e.mean().item() - 1.0) < 2e-2 and abs(e.std().item() - 2.0) < 2e-2
    ext.init_rows(values, D, rows, keys, ext.InitializerMode.TRUNCATED_NORMAL, 0.0, 1.0, -0.5, 0.5, seed=9)
    e = values[:, :D]
    assert e.min() >= -0.5 and e.max() <= 0.5 and e.std().item() > 0.2
    ext.init_rows(values, D, rows, keys, ext.InitializerMode.CONSTANT, 3.25)
    assert torch.equal(values[:, :D], torch.full((n, D), 3.25, device=cuda))


OPTS = [("sgd", 1, 0), ("adagrad", 3, 1), ("adam", 2, 2), ("rowwise_adagrad", 4, None)]


@pytest.mark.parametrize("opt,code,state_mult", OPTS)
@pytest.mark.parametrize("mode", ["seq", "sum", "mean"])
def test_backward_reduce_and_update(cuda, opt, code, state_mult, mode):
    from dynamicemb import dynamicemb_extensions as ext
    from oracle import dynamicemb as orc
    D = 128
    rng = np.random.default_rng(code * 10 + len(mode))
    state = 4 if state_mult is None else state_mult * D
    vdim = D + state
    R = 3000
    values = torch.rand(R, vdim, device=cuda) + 0.1
    if mode == "seq":
        n = 20000
        B = F = 0
        offsets = None
    else:
        B, F = 64, 4
        lens = rng.integers(0, 60, size=B * F)
        offsets = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
        n = int(offsets[-1])
    # Zipf-hot unique ids: one unique id spans many 32-row tiles (exercises the span kernel)
    nu = 700
    inverse = np.minimum(rng.zipf(1.3, size=n) - 1, nu - 1).astype(np.int64)
    inverse[:nu] = np.arange(nu)            # every unique idx appears
    rows = rng.permutation(R)[:nu].astype(np.int64)
    rows[rng.integers(0, nu, size=10)] = -1  # failed inserts are skipped (optimizer_kernel.cuh:420-422)
    if mode == "seq":
        grads = torch.randn(n, D, device=cuda)
        grow = np.arange(n); scale = np.ones(n, dtype=np.float32)
        g_np = grads.cpu().numpy()
    else:
        grads = torch.randn(B, F * D, device=cuda)
        bag = np.searchsorted(offsets, np.arange(n), side="right") - 1     # bag = f*B+b
        f, b = bag // B, bag % B
        grow = b * F + f
        blen = (offsets[1:] - offsets[:-1])[bag]
        scale = (np.float32(1.0) / blen.astype(np.float32)) if mode == "mean" else np.ones(n, dtype=np.float32)
        g_np = grads.cpu().numpy().reshape(B * F, D)
    comb = {"seq": -1, "sum": 0, "mean": 1}[mode]
    kw = dict(lr=0.05, eps=1e-8, beta1=0.9, beta2=0.999, weight_decay=0.01)
    step = 3
    vals_before = values.cpu().numpy().copy()
    ug = ext.backward(values, D, torch.from_numpy(inverse).to(cuda), nu, torch.from_numpy(rows).to(cuda), grads,
                      offsets=None if offsets is None else torch.from_numpy(offsets).to(cuda), batch_size=B, num_features=F, combiner=comb,
                      opt_type=code, bc1=1 - 0.9 ** step, bc2=1 - 0.999 ** step, want_unique_grads=True, **kw)
    exp_ug = orc.reduce_grads(inverse, g_np[grow], scale, nu, D)
    assert np.array_equal(ug.cpu().numpy(), exp_ug), "reduced gradients must be bit-exact (fixed summation order)"
    exp_vals = orc.optimizer_update(vals_before.copy(), D, rows, exp_ug, opt, step=step, **kw)
    np.testing.assert_allclose(values.cpu().numpy(), exp_vals, rtol=1e-6, atol=1e-6)
    # standalone op pair: reduce_grads + update_rows gives the same rows
    v2 = torch.from_numpy(vals_before).to(cuda)
    ug2 = ext.reduce_grads(torch.from_numpy(inverse).to(cuda), grads, nu, B, D, offsets=None if offsets is None else torch.from_numpy(offsets).to(cuda),
                           combiner=comb, total_D=F * D)
    assert torch.equal(ug2, ug)
    ext.update_rows(v2, D, torch.from_numpy(rows).to(cuda), ug2, code, bc1=1 - 0.9 ** step, bc2=1 - 0.999 ** step, **kw)
    assert torch.equal(v2, values)


def test_block_bucketize(cuda):
    from dynamicemb import dynamicemb_extensions as ext
    from oracle import dynamicemb as orc
    rng = np.random.default_rng(23)
    B, F, W = 9, 3, 8
    lens = rng.integers(0, 150, size=F * B)
    n = int(lens.sum())
    ids = rng.integers(0, 1 << 50, size=n, dtype=np.int64)
    ids[:50] = rng.integers(0, 4000, size=50)
    blk = np.array([1000, 1 << 47, 77], dtype=np.int64)
    for dts in ([0, 0, 0], [1, 1, 1], [2, 2, 2], [0, 1, 2]):
        nl, ni, _, perm = ext.block_bucketize_sparse_features(torch.from_numpy(lens).to(cuda), torch.from_numpy(ids).to(cuda), B, W,
                                                              torch.from_numpy(blk).to(cuda), torch.tensor(dts, dtype=torch.int32, device=cuda))
        enl, eni, eperm = orc.block_bucketize(lens, ids, B, W, blk, dts)
        assert np.array_equal(nl.cpu().numpy(), enl)
        assert np.array_equal(ni.cpu().numpy(), eni)
        assert np.array_equal(perm.cpu().numpy(), eperm)


def test_block_bucketize_short_slots(cuda):
    """<= 4 ids per slot on average (what index dedup produces): the thread-per-slot kernels must give the warp kernels' exact output."""
    from dynamicemb import dynamicemb_extensions as ext
    from oracle import dynamicemb as orc
    rng = np.random.default_rng(29)
    B, F, W = 1000, 2, 3
    lens = rng.integers(0, 4, size=F * B)
    lens[rng.integers(0, F * B, size=5)] = 9            # a few longer slots inside the short regime
    n = int(lens.sum())
    assert n <= 4 * F * B
    ids = rng.integers(0, 1 << 50, size=n, dtype=np.int64)
    blk = np.array([1 << 40, 12345], dtype=np.int64)
    for dts in ([0, 0], [1, 2], [2, 2]):
        nl, ni, _, perm = ext.block_bucketize_sparse_features(torch.from_numpy(lens).to(cuda), torch.from_numpy(ids).to(cuda), B, W,
                                                              torch.from_numpy(blk).to(cuda), torch.tensor(dts, dtype=torch.int32, device=cuda))
        enl, eni, eperm = orc.block_bucketize(lens, ids, B, W, blk, dts)
        assert np.array_equal(nl.cpu().numpy(), enl)
        assert np.array_equal(ni.cpu().numpy(), eni)
        assert np.array_equal(perm.cpu().numpy(), eperm)
