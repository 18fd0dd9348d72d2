"""GPU: BatchedDynamicEmbeddingTablesV2 end to end (prefetch -> forward -> fused backward) against a dense
torch reference, the way the reference checks itself against FBGEMM TBE
(corelib/dynamicemb/test/test_batched_dynamic_embedding_tables_v2.py:1594-1745 test_backward: same fixed
indices/offsets, 10 iterations, assert_close 1e-6) plus eval/empty-batch/eviction cases (:1436, :2078)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

IDX = [0, 1, 12, 64, 8, 12, 15, 2, 7, 105, 0]
OFF = [0, 2, 3, 5, 6, 9, 10, 10, 11]


def _module(cuda, opt, pooling, D=128, n_tables=2, cap=4096, **kw):
    from dynamicemb import (BatchedDynamicEmbeddingTablesV2, DynamicEmbInitializerArgs, DynamicEmbInitializerMode, DynamicEmbScoreStrategy,
                            DynamicEmbTableOptions)
    opts = [DynamicEmbTableOptions(dim=D, max_capacity=cap, local_hbm_for_values=1 << 40, score_strategy=DynamicEmbScoreStrategy.STEP,
                                   initializer_args=DynamicEmbInitializerArgs(mode=DynamicEmbInitializerMode.DEBUG)) for _ in range(n_tables)]
    return BatchedDynamicEmbeddingTablesV2(opts, table_names=[f"t{i}" for i in range(n_tables)], pooling_mode=pooling, optimizer=opt, device=cuda, **kw)


class DenseRef:
    """Per-table dict of fp32 rows with the same debug initializer + optimizer math (fp32 torch ops)."""

    def __init__(self, D, n_tables, opt, lr, eps=1e-8, beta1=0.9, beta2=0.999, wd=0.0, acc0=0.0):
        self.D, self.opt, self.lr, self.eps, self.b1, self.b2, self.wd, self.acc0 = D, opt, lr, eps, beta1, beta2, wd, acc0
        self.w = [dict() for _ in range(n_tables)]
        self.s = [dict() for _ in range(n_tables)]
        self.it = 0

    def row(self, t, k):
        if k not in self.w[t]:
            self.w[t][k] = torch.full((self.D,), float(k % 100000))
            self.s[t][k] = [torch.full((self.D,), self.acc0), torch.zeros(self.D)]
        return self.w[t][k]

    def step(self, grads_by_key):
        self.it += 1
        for (t, k), g in grads_by_key.items():
            w, st = self.w[t][k], self.s[t][k]
            if self.opt == "sgd":
                w -= self.lr * g
            elif self.opt == "adagrad":
                st[0] += g * g
                w -= self.lr * g / (st[0].sqrt() + self.eps)
            elif self.opt == "adam":
                f32 = lambda x: torch.tensor(x, dtype=torch.float32)
                omb1, omb2 = f32(1.0) - f32(self.b1), f32(1.0) - f32(self.b2)     # fp32 (1 - beta), as optimizer_kernel.cuh:115-131
                st[0].mul_(self.b1).add_(omb1 * g)
                st[1].mul_(self.b2).add_(omb2 * g * g)
                mh, vh = st[0] / (1 - self.b1 ** self.it), st[1] / (1 - self.b2 ** self.it)
                w -= self.lr * (mh / (vh.sqrt() + self.eps) + self.wd * w)
            elif self.opt == "rowwise":
                st[0][0] += (g * g).mean()
                w -= self.lr * g / (st[0][0].sqrt() + self.eps)


@pytest.mark.parametrize("opt_name", ["sgd", "adagrad", "adam", "rowwise"])
@pytest.mark.parametrize("pooling", ["none", "sum", "mean"])
def test_train_matches_dense_reference(cuda, opt_name, pooling):
    from dynamicemb import DynamicEmbPoolingMode, EmbOptimType
    D, T = 128, 2
    pm = {"none": DynamicEmbPoolingMode.NONE, "sum": DynamicEmbPoolingMode.SUM, "mean": DynamicEmbPoolingMode.MEAN}[pooling]
    ot = {"sgd": EmbOptimType.SGD, "adagrad": EmbOptimType.EXACT_ADAGRAD, "adam": EmbOptimType.ADAM, "rowwise": EmbOptimType.EXACT_ROWWISE_ADAGRAD}[opt_name]
    lr = 0.3
    m = _module(cuda, ot, pm, D=D, n_tables=T, learning_rate=lr, weight_decay=0.0, initial_accumulator_value=0.0)
    m.train()
    ref = DenseRef(D, T, opt_name, lr)
    indices = torch.tensor(IDX, dtype=torch.int64, device=cuda)
    offsets = torch.tensor(OFF, dtype=torch.int64, device=cuda)
    F, B = T, (len(OFF) - 1) // T          # 2 features (one per table) x 4 samples
    for it in range(10):
        out = m(indices, offsets)
        # reference forward
        rows, owner = [], []
        for g in range(F * B):
            t = g // B
            for i in range(OFF[g], OFF[g + 1]):
                rows.append(ref.row(t, IDX[i]).clone()); owner.append((t, IDX[i], g))
        if pooling == "none":
            exp = torch.stack(rows)
        else:
            exp = torch.zeros(B, F * D)
            for r, (t, k, g) in zip(rows, owner):
                f, b = g // B, g % B
                L = OFF[g + 1] - OFF[g]
                exp[b, f * D:(f + 1) * D] += r / L if pooling == "mean" else r
        torch.testing.assert_close(out.cpu(), exp, rtol=1e-6, atol=1e-6)
        gen = torch.Generator().manual_seed(it)
        gout = torch.randn(out.shape, generator=gen)
        out.backward(gout.to(cuda))
        gk = {}
        for j, (t, k, g) in enumerate(owner):
            if pooling == "none":
                gr = gout[j]
            else:
                f, b = g // B, g % B
                L = OFF[g + 1] - OFF[g]
                gr = gout[b, f * D:(f + 1) * D] / (L if pooling == "mean" else 1)
            gk[(t, k)] = gk.get((t, k), torch.zeros(D)) + gr
        ref.step(gk)
    # final weights
    for t in range(T):
        keys, vals = m.export_keys_values(t)
        got = {int(k): v[:D].cpu() for k, v in zip(keys.tolist(), vals)}
        assert set(got) == set(ref.w[t])
        for k, w in ref.w[t].items():
            torch.testing.assert_close(got[k], w, rtol=2e-5, atol=2e-5)


def test_eval_forward_and_missing(cuda):
    from dynamicemb import DynamicEmbPoolingMode, EmbOptimType
    m = _module(cuda, EmbOptimType.SGD, DynamicEmbPoolingMode.NONE, n_tables=1)
    m.train()
    idx = torch.arange(10, 500, dtype=torch.int64, device=cuda)
    off = torch.arange(0, idx.numel() + 1, dtype=torch.int64, device=cuda)
    with torch.no_grad():
        tr = m(idx, off)
    assert torch.equal(tr, (idx % 100000).to(torch.float32)[:, None].expand(-1, 128))
    m.eval()
    q = torch.cat([idx[:100], torch.arange(10_000, 10_050, device=cuda)])
    out = m(q, torch.arange(0, q.numel() + 1, dtype=torch.int64, device=cuda))
    assert torch.equal(out[:100], tr[:100]) and torch.count_nonzero(out[100:]) == 0     # absent => eval initializer zeros (DE/README.md:108)
    assert m.tables.size() == idx.numel()     # eval does not insert
    assert int(m.tables._ref_counter.sum().item()) == 0   # all pins released


def test_eviction_steady_state_keeps_training(cuda):
    """Tiny table, STEP scores: old keys get evicted, current batch always present, pins balanced."""
    from dynamicemb import DynamicEmbPoolingMode, EmbOptimType
    m = _module(cuda, EmbOptimType.EXACT_ADAGRAD, DynamicEmbPoolingMode.SUM, D=64, n_tables=1, cap=1024, learning_rate=0.1)
    m.train()
    rng = np.random.default_rng(0)
    for it in range(30):
        ids = torch.from_numpy(rng.integers(0, 1 << 30, size=600, dtype=np.int64)).to(cuda)
        off = torch.arange(0, 601, 6, dtype=torch.int64, device=cuda)
        out = m(ids, off)
        out.sum().backward()
        assert torch.isfinite(out).all()
    assert m.tables.size() <= 1024
    assert int(m.tables._ref_counter.sum().item()) == 0


@pytest.mark.parametrize("pooling,T", [("none", 1), ("sum", 2)])
def test_fused_prefetch_matches_op_by_op(cuda, pooling, T):
    """The fused no-sync prefetch (csrc/demb_train.cu) must leave byte-identical table images, value rows and outputs as the
    op-by-op path (lookup -> insert -> init_rows), including under eviction pressure (small table)."""
    from dynamicemb import DynamicEmbPoolingMode, EmbOptimType
    pm = {"none": DynamicEmbPoolingMode.NONE, "sum": DynamicEmbPoolingMode.SUM}[pooling]
    ma = _module(cuda, EmbOptimType.EXACT_ADAGRAD, pm, D=64, n_tables=T, cap=2048, learning_rate=0.05, fused_prefetch=True)
    mb = _module(cuda, EmbOptimType.EXACT_ADAGRAD, pm, D=64, n_tables=T, cap=2048, learning_rate=0.05, fused_prefetch=False)
    ma.train(); mb.train()
    rng = np.random.default_rng(42)
    B = 32
    for it in range(40):
        lens = rng.integers(0, 24, size=T * B)
        offsets = torch.from_numpy(np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)).to(cuda)
        ids = torch.from_numpy((rng.zipf(1.1, size=int(lens.sum())) % 20000).astype(np.int64) * 7919).to(cuda)
        oa, ob = ma(ids, offsets), mb(ids, offsets)
        assert torch.equal(oa, ob), f"iter {it}: outputs differ"
        g = torch.randn_like(oa)
        oa.backward(g); ob.backward(g)
        assert torch.equal(ma.tables.table_storage_, mb.tables.table_storage_), f"iter {it}: table image differs"
        assert torch.equal(ma._values, mb._values), f"iter {it}: value rows differ"
        assert torch.equal(ma.tables.bucket_sizes, mb.tables.bucket_sizes)
        assert int(ma.tables._ref_counter.sum().item()) == 0 and int(mb.tables._ref_counter.sum().item()) == 0
        assert int((ma.tables._bucket_heads != -1).sum().item()) == 0
    assert ma.tables.size() > 1000


def test_fused_prefetch_with_erasures_and_negative_keys(cuda):
    """Same equivalence with the cases the thread-per-bucket insert adds: erased (reclaimable) slots inside buckets that are not
    full, and negative int64 keys (the deterministic order inside a bucket is by SIGNED key)."""
    from dynamicemb import DynamicEmbPoolingMode, EmbOptimType
    ma = _module(cuda, EmbOptimType.EXACT_ADAGRAD, DynamicEmbPoolingMode.NONE, D=64, n_tables=1, cap=4096, learning_rate=0.05, fused_prefetch=True)
    mb = _module(cuda, EmbOptimType.EXACT_ADAGRAD, DynamicEmbPoolingMode.NONE, D=64, n_tables=1, cap=4096, learning_rate=0.05, fused_prefetch=False)
    ma.train(); mb.train()
    rng = np.random.default_rng(77)
    n = 600
    offsets = torch.arange(0, n + 1, dtype=torch.int64, device=cuda)
    seen = set()
    for it in range(30):
        raw = (rng.zipf(1.15, size=n) % 5000).astype(np.int64) * 104729
        raw[rng.random(n) < 0.4] *= -1                                   # negative keys are legal (only the top four values are reserved)
        ids = torch.from_numpy(raw).to(cuda)
        oa, ob = ma(ids, offsets), mb(ids, offsets)
        assert torch.equal(oa, ob), f"iter {it}: outputs differ"
        g = torch.randn_like(oa)
        oa.backward(g); ob.backward(g)
        seen.update(raw.tolist())
        if it % 3 == 2:                                                  # erase a third of what is there: leaves reclaimable slots behind
            pool = np.array(sorted(seen), dtype=np.int64)
            drop = rng.choice(pool, size=max(1, pool.size // 3), replace=False)
            dk = torch.from_numpy(drop).to(cuda)
            z = torch.zeros(drop.size, dtype=torch.int64, device=cuda)
            ma.tables.erase(dk, z); mb.tables.erase(dk, z)
            seen.difference_update(drop.tolist())
        assert torch.equal(ma.tables.table_storage_, mb.tables.table_storage_), f"iter {it}: table image differs"
        assert torch.equal(ma._values, mb._values), f"iter {it}: value rows differ"
        assert torch.equal(ma.tables.bucket_sizes, mb.tables.bucket_sizes)
        assert int(ma.tables._ref_counter.sum().item()) == 0 and int((ma.tables._bucket_heads != -1).sum().item()) == 0
    assert ma.tables.size() > 200


def test_graphed_step_matches_eager(cuda):
    """make_graphed_step (one CUDA graph per training step) leaves the same table / rows / outputs as the eager step."""
    from dynamicemb import DynamicEmbPoolingMode, EmbOptimType
    ma = _module(cuda, EmbOptimType.EXACT_ADAGRAD, DynamicEmbPoolingMode.NONE, D=64, n_tables=1, cap=4096, learning_rate=0.05)
    mb = _module(cuda, EmbOptimType.EXACT_ADAGRAD, DynamicEmbPoolingMode.NONE, D=64, n_tables=1, cap=4096, learning_rate=0.05)
    ma.train(); mb.train()
    n = 3000
    rng = np.random.default_rng(3)
    ids_static = torch.zeros(n, dtype=torch.int64, device=cuda)
    offsets = torch.arange(0, n + 1, dtype=torch.int64, device=cuda)
    grad = torch.randn(n, 64, device=cuda)
    batches = [torch.from_numpy((rng.zipf(1.1, size=n) % 30000).astype(np.int64) * 31).to(cuda) for _ in range(12)]
    ids_static.copy_(batches[0])
    graph, out, loss = ma.make_graphed_step(ids_static, offsets, grad)      # warm-up (3) + capture (not executed) on batches[0]
    for _ in range(3):
        o = mb(batches[0], offsets); o.backward(grad)
    for b in batches[1:]:
        ids_static.copy_(b)
        graph.replay()
        o = mb(b, offsets)
        l = o.detach().sum()
        o.backward(grad)
        assert torch.equal(out, o) and torch.equal(loss, l)
        assert torch.equal(ma.tables.table_storage_, mb.tables.table_storage_)
        assert torch.equal(ma._values, mb._values)
    assert int(ma.tables._ref_counter.sum().item()) == 0
