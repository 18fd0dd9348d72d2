"""GPU: the kernels every bench step runs — the fused training prefetch (csrc/demb_train.cu: train_lookup / train_insert_thread /
train_insert / train_init_rows), the forward gather and the fused backward — replayed DIRECTLY against the CPU oracle
(oracle/dynamicemb_oracle.c, the restatement of the reference table pinned by the reference-kernel goldens), at the bench's own shape:
2^20 power-law ids per step, a table small enough that eviction starts inside the run, and a tiny table where every step evicts.

Per step the oracle does what the reference's _prefetch_hbm_direct_path does (batched_dynamicemb_function.py:559-699): unique ->
lookup (ASSIGN score = step) -> pin the hits -> deterministic insert of the misses (victim = min score, unpinned) -> pin.  Bit-exact:
table image bytes, slot of every unique key, bucket_sizes.  Value rows: debug initializer (key % 100000) on every (re)inserted row."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
D = 128


def _power_law(n, rng, hi=1e9, alpha=1.05):
    y = rng.random(n)
    g = 1.0 - alpha
    return ((hi ** g - 1.0) * y + 1.0) ** (1.0 / g)


def _module(cuda, cap, D=D, opt=None, pooling=None, lr=0.1):
    from dynamicemb import (BatchedDynamicEmbeddingTablesV2, DynamicEmbInitializerArgs, DynamicEmbInitializerMode, DynamicEmbPoolingMode,
                            DynamicEmbScoreStrategy, DynamicEmbTableOptions, EmbOptimType)
    o = DynamicEmbTableOptions(dim=D, max_capacity=cap, local_hbm_for_values=1 << 50, score_strategy=DynamicEmbScoreStrategy.STEP,
                               initializer_args=DynamicEmbInitializerArgs(mode=DynamicEmbInitializerMode.DEBUG))
    return BatchedDynamicEmbeddingTablesV2([o], table_names=["t"], pooling_mode=DynamicEmbPoolingMode.NONE if pooling is None else pooling,
                                           optimizer=EmbOptimType.EXACT_ADAGRAD if opt is None else opt, learning_rate=lr, eps=1e-8, device=cuda)


def _first_occurrence_unique(ids):
    uk, first = np.unique(ids, return_index=True)
    order = np.argsort(first, kind="stable")
    return uk[order]


def _oracle_prefetch(tab, ids, step):
    """One reference prefetch on the oracle table; returns (unique keys in first-occurrence order, slot per unique key, found mask)."""
    uk = _first_occurrence_unique(ids)
    score = np.full(uk.size, step, dtype=np.int64)
    _, found, slots = tab.lookup(uk, None, policy=1, score_in=score)
    np.add.at(tab.counter, slots[found], 1)                       # single table: global slot == table-local slot
    miss = ~found
    if miss.any():
        ns, res, _, _ = tab.insert(uk[miss], None, policy=1, score_in=score[miss], deterministic=True, use_counter=True)
        slots[miss] = ns
        np.add.at(tab.counter, ns[ns >= 0], 1)
    return uk, slots, found


def _segment_sum(inv, rows, nu):
    """sum of rows[i] over ids with inv[i] == u, float64 (sort + reduceat)"""
    order = np.argsort(inv, kind="stable")
    si = inv[order]
    starts = np.flatnonzero(np.concatenate([[True], si[1:] != si[:-1]]))
    out = np.zeros((nu, rows.shape[1]), dtype=np.float64)
    out[si[starts]] = np.add.reduceat(rows[order], starts, axis=0)
    return out


def _replay(cuda, cap, n_ids, steps, key_hi, seed, expect_eviction):
    from oracle.dynamicemb import OracleTable
    m = _module(cuda, cap)
    m.train()
    tab = OracleTable([cap], 128)
    rng = np.random.default_rng(seed)
    offsets = torch.arange(0, n_ids + 1, dtype=torch.int64, device=cuda)
    for step in range(1, steps + 1):
        ids_np = _power_law(n_ids, rng, hi=key_hi).astype(np.int64)
        ids = torch.from_numpy(ids_np).to(cuda)
        uk, slots, _ = _oracle_prefetch(tab, ids_np, step)
        # ---- product: the fused prefetch alone, then the forward gather through the public forward()
        m.prefetch(ids, offsets)
        st = m._prefetch_states[0]
        nu = int(st.num_unique_dev.item())
        assert nu == uk.size, f"step {step}: unique count"
        assert np.array_equal(st.unique_keys[:nu].cpu().numpy(), uk), f"step {step}: unique keys / first-occurrence order"
        got_slots = st.slot_indices[:nu].cpu().numpy()
        assert np.array_equal(got_slots, slots), f"step {step}: slot assignment differs from the oracle in {(got_slots != slots).sum()} keys"
        assert np.array_equal(m.tables.table_storage_.cpu().numpy(), tab.storage), f"step {step}: table image differs"
        assert np.array_equal(m.tables.bucket_sizes.cpu().numpy(), tab.bucket_sizes), f"step {step}: bucket_sizes"
        assert np.array_equal(m.tables._ref_counter.cpu().numpy(), tab.counter), f"step {step}: pins"
        with torch.no_grad():
            out = m(ids, offsets)                                   # consumes the queued prefetch state, unpins
        # every id's row holds its key's debug value (rows of evicted keys were re-initialised for their new owner); failed inserts
        # (slot -1: every candidate pinned) read zeros
        want = torch.from_numpy((ids_np % 100000).astype(np.float32)).to(cuda)
        ok_slot = torch.from_numpy(slots >= 0).to(cuda)[st.reverse_indices]
        assert torch.equal(out[:, 0], torch.where(ok_slot, want, torch.zeros_like(want))), f"step {step}: forward rows"
        assert torch.equal(out, out[:, :1].expand(-1, D)), f"step {step}: row contents"
        tab.counter[:] = 0                                          # unpin (decrement_counter after the step)
        assert int(m.tables._ref_counter.abs().sum().item()) == 0
    load = tab.bucket_sizes.sum() / cap
    if expect_eviction:
        assert load > 0.97, f"table never filled up (load {load:.3f}): the run did not reach eviction"
    return load


def test_fused_prefetch_vs_oracle_bench_shape_with_eviction(cuda):
    """2^20 power-law ids / step on a 1 Mi-row table: ~0.39 M unique and ~0.2-0.3 M new keys per step, full after ~4 steps,
    then every step evicts ~0.2 M keys (min-score victim among the unpinned slots of the bucket)."""
    load = _replay(cuda, cap=1 << 20, n_ids=1 << 20, steps=10, key_hi=1e9, seed=11, expect_eviction=True)
    assert load > 0.97


def test_fused_prefetch_vs_oracle_every_step_evicts(cuda):
    """16-bucket table (2048 slots), 1500 ids per step from a 2^30 key space: from the second step on every insert evicts, buckets see
    ~50 new keys per step each (long per-bucket lists -> the warp-cooperative insert kernel and its selection path)."""
    load = _replay(cuda, cap=2048, n_ids=1500, steps=25, key_hi=2.0 ** 30, seed=5, expect_eviction=True)
    assert load > 0.97


@pytest.mark.parametrize("pooling", ["sum", "mean"])
def test_full_size_adagrad_pooled_step_vs_oracle(cuda, pooling):
    """BASELINE-size pooled training steps (2 features x 52429 bags x hotness 10 = 1 048 580 ids, D=128, fused Adagrad) against the
    oracle: rows from the oracle table, pooled forward bit-exact against oracle.pool_rows (same accumulation order), per-key gradient
    sums + Adagrad row update against the oracle's optimizer restatement (float64 sums; fp32 tolerance stated below)."""
    from dynamicemb import DynamicEmbPoolingMode
    from oracle.dynamicemb import OracleTable, pool_rows
    F, B, H = 2, 52429, 10
    n = F * B * H
    cap = 4 << 20
    lr = 0.1
    pm = DynamicEmbPoolingMode.SUM if pooling == "sum" else DynamicEmbPoolingMode.MEAN
    # one TABLE, two features: feature_table_map [0, 0]
    from dynamicemb import (BatchedDynamicEmbeddingTablesV2, DynamicEmbInitializerArgs, DynamicEmbInitializerMode, DynamicEmbScoreStrategy,
                            DynamicEmbTableOptions, EmbOptimType)
    o = DynamicEmbTableOptions(dim=D, max_capacity=cap, local_hbm_for_values=1 << 50, score_strategy=DynamicEmbScoreStrategy.STEP,
                               initializer_args=DynamicEmbInitializerArgs(mode=DynamicEmbInitializerMode.DEBUG))
    m = BatchedDynamicEmbeddingTablesV2([o], table_names=["t"], feature_table_map=[0, 0], pooling_mode=pm, optimizer=EmbOptimType.EXACT_ADAGRAD,
                                        learning_rate=lr, eps=1e-8, device=cuda)
    m.train()
    tab = OracleTable([cap], 128)
    values = np.zeros((cap, 2 * D), dtype=np.float32)               # oracle value rows [emb | Adagrad state]
    rng = np.random.default_rng(21)
    lens = np.full(F * B, H, dtype=np.int64)
    lens[rng.integers(0, F * B, size=2000)] = 0                     # some empty bags
    lens[rng.integers(0, F * B, size=2000)] = 2 * H
    offs_np = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    n = int(offs_np[-1])
    offsets = torch.from_numpy(offs_np).to(cuda)
    for step in range(1, 3):
        ids_np = _power_law(n, rng).astype(np.int64)
        ids = torch.from_numpy(ids_np).to(cuda)
        uk, slots, found = _oracle_prefetch(tab, ids_np, step)
        assert (slots >= 0).all()
        inv = np.searchsorted(np.sort(uk), ids_np)
        inv = np.argsort(uk, kind="stable")[inv]                    # id -> index into uk
        # new rows: debug initializer, zero state
        fresh = ~found
        values[slots[fresh], D:] = 0.0
        values[slots[fresh], :D] = (uk[fresh] % 100000).astype(np.float32)[:, None]
        out = m(ids, offsets)
        exp = pool_rows(values, D, offs_np, slots[inv], 0 if pooling == "sum" else 1, B, F)
        assert np.array_equal(out.detach().cpu().numpy(), exp), f"step {step}: pooled forward differs from the oracle"
        g = torch.randn(B, F * D, device=cuda, generator=torch.Generator(device=cuda).manual_seed(step)) * 1e-2
        out.backward(g)
        # ---- oracle backward: gradient row of id i = grad[b, f*D:(f+1)*D] (x 1/len for MEAN), summed per key in float64
        g_np = g.cpu().numpy().reshape(B, F, D)
        bag = np.repeat(np.arange(F * B), lens)                     # bag index f*B + b of every id
        f_idx, b_idx = bag // B, bag % B
        rows_g = g_np[b_idx, f_idx].astype(np.float64)
        if pooling == "mean":
            rows_g = (g_np[b_idx, f_idx] * (1.0 / lens[bag].astype(np.float32))[:, None]).astype(np.float64)   # kernel: fmul by fp32 1/len
        gsum = _segment_sum(inv, rows_g, uk.size)
        absum = _segment_sum(inv, np.abs(rows_g), uk.size)
        w = values[slots, :D].astype(np.float64)
        s = values[slots, D:].astype(np.float64)
        s_new = s + gsum * gsum
        w_new = w - lr * gsum / (np.sqrt(s_new) + 1e-8)
        got = m._values[torch.from_numpy(slots).to(cuda)].cpu().numpy().astype(np.float64)
        # fp32 tolerance: the kernel sums the rows of a key in fp32 (fixed tile order); error <= 1e-5 x sum|g| on the gradient sum,
        # propagated through d(state) = 2 g dg and dw <= lr * dg / sqrt(state) (+ fp32 rounding of the row itself)
        dg = 1e-5 * absum + 1e-12
        tol_s = 2 * np.abs(gsum) * dg + 1e-6 * np.abs(s_new) + 1e-12
        tol_w = lr * dg / (np.sqrt(s_new) + 1e-8) * 2 + 2e-6 * np.abs(w_new) + 1e-7
        assert (np.abs(got[:, D:] - s_new) <= tol_s).all(), f"step {step}: Adagrad state, max excess {(np.abs(got[:, D:] - s_new) - tol_s).max()}"
        assert (np.abs(got[:, :D] - w_new) <= tol_w).all(), f"step {step}: weights, max excess {(np.abs(got[:, :D] - w_new) - tol_w).max()}"
        values[slots] = got.astype(np.float32)                      # continue from the product's rows (errors do not compound in the check)
        tab.counter[:] = 0
    assert int(m.tables._ref_counter.abs().sum().item()) == 0
