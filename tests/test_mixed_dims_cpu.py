"""Mixed embedding dims inside one pooled module (reference: `dims` / `max_D` / `D_offsets` of gather_embedding_pooled,
corelib/dynamicemb/src/lookup_kernel.cuh:901-962, DynamicEmbeddingFunction.forward batched_dynamicemb_function.py:1066-1068): output
`[B, total_D]` with feature f at columns D_offsets[f] : D_offsets[f] + dims[table(f)].  Host logic on the op-layer shim
(tests/cpu_ext_shim.py); tests/test_zz_mixed_dims_gpu.py runs the same scenarios on the kernels."""
import numpy as np
import pytest
import torch

from tests.cpu_ext_shim import patched_module

LR = 0.5


def _mixed_module(dev, dims, fmap, optimizer, mean=False, **kw):
    from dynamicemb import (BatchedDynamicEmbeddingTablesV2, DynamicEmbInitializerArgs, DynamicEmbInitializerMode, DynamicEmbPoolingMode,
                            DynamicEmbScoreStrategy, DynamicEmbTableOptions)
    opts = [DynamicEmbTableOptions(dim=d, max_capacity=1024, bucket_capacity=128, score_strategy=DynamicEmbScoreStrategy.STEP,
                                   initializer_args=DynamicEmbInitializerArgs(mode=DynamicEmbInitializerMode.DEBUG)) for d in dims]
    return BatchedDynamicEmbeddingTablesV2(opts, table_names=[f"t{i}" for i in range(len(dims))], feature_table_map=fmap,
                                           pooling_mode=DynamicEmbPoolingMode.MEAN if mean else DynamicEmbPoolingMode.SUM,
                                           optimizer=optimizer, learning_rate=LR, eps=1e-8, device=dev, fused_prefetch=False, **kw)


def scenario_mixed_dims(dev, optimizer_name, mean, tmp_path=None):
    from dynamicemb import EmbOptimType
    opt = {"sgd": EmbOptimType.SGD, "adagrad": EmbOptimType.EXACT_ADAGRAD, "adam": EmbOptimType.ADAM}[optimizer_name]
    dims, fmap, B = [32, 64, 128], [0, 1, 1, 2], 5       # table 1 serves two features; widths the GPU suite validates
    F = len(fmap)
    doff = np.concatenate([[0], np.cumsum([dims[t] for t in fmap])])
    m = _mixed_module(dev, dims, fmap, opt, mean)
    TD = int(doff[-1])
    assert m.total_D == TD == 288 and m.max_D == 128
    m.train()
    rng = np.random.default_rng(3)
    # model per table: key -> (weight vector, optimizer state) in float64; DEBUG init = key % 100000 in every column
    W = [dict() for _ in dims]
    S = [dict() for _ in dims]
    for step in range(1, 5):
        lens = rng.integers(0, 4, size=F * B)
        ids = rng.integers(1, 30, size=int(lens.sum())).astype(np.int64)
        off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
        out = m(torch.from_numpy(ids).to(dev), torch.from_numpy(off).to(dev))
        assert tuple(out.shape) == (B, TD)
        want = np.zeros((B, TD))
        for f, t in enumerate(fmap):
            for b in range(B):
                s, e = off[f * B + b], off[f * B + b + 1]
                for k in ids[s:e].tolist():
                    w = W[t].get(k, np.full(dims[t], float(k % 100000)))
                    want[b, doff[f]:doff[f + 1]] += w / ((e - s) if mean else 1)
        assert np.allclose(out.detach().cpu().numpy(), want, rtol=1e-5, atol=1e-3), f"step {step} forward"
        g = rng.standard_normal((B, TD)).astype(np.float32)
        out.backward(torch.from_numpy(g).to(dev))
        # per-key gradient = sum over its occurrences of the bag's gradient slice (divided by the bag length for MEAN)
        G = [dict() for _ in dims]
        for f, t in enumerate(fmap):
            for b in range(B):
                s, e = off[f * B + b], off[f * B + b + 1]
                for k in ids[s:e].tolist():
                    G[t][k] = G[t].get(k, np.zeros(dims[t])) + g[b, doff[f]:doff[f + 1]].astype(np.float64) / ((e - s) if mean else 1)
        for t in range(len(dims)):
            for k, gk in G[t].items():
                w = W[t].get(k, np.full(dims[t], float(k % 100000)))
                if optimizer_name == "sgd":
                    w = w - LR * gk
                elif optimizer_name == "adagrad":
                    acc = S[t].get(k, np.zeros(dims[t])) + gk * gk
                    S[t][k] = acc
                    w = w - LR * gk / (np.sqrt(acc) + 1e-8)
                else:
                    mm, vv = S[t].get(k, (np.zeros(dims[t]), np.zeros(dims[t])))
                    mm, vv = 0.9 * mm + 0.1 * gk, 0.999 * vv + 0.001 * gk * gk
                    S[t][k] = (mm, vv)
                    w = w - LR * (mm / (1 - 0.9 ** step)) / (np.sqrt(vv / (1 - 0.999 ** step)) + 1e-8)
                W[t][k] = w
    for t in range(len(dims)):
        keys, emb = m.export_keys_values(f"t{t}", device=torch.device("cpu"))
        assert emb.shape[1] == dims[t] and set(keys.tolist()) == set(W[t])
        for k, row in zip(keys.tolist(), emb.numpy()):
            assert np.allclose(row, W[t][k], rtol=2e-4, atol=2e-3), (t, k)
    if tmp_path is not None:            # checkpoint round trip incl. optimizer state of the narrower tables, then one more identical step
        m.dump(str(tmp_path), optim=True)
        m2 = _mixed_module(dev, dims, fmap, opt, mean)
        m2.load(str(tmp_path), optim=True)
        m2._optimizer.iter = m._optimizer.iter
        m.train(), m2.train()
        lens = rng.integers(1, 4, size=F * B)
        ids = rng.integers(1, 30, size=int(lens.sum())).astype(np.int64)
        off = torch.from_numpy(np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)).to(dev)
        g = torch.from_numpy(rng.standard_normal((B, TD)).astype(np.float32)).to(dev)
        outs = []
        for mod in (m, m2):
            o = mod(torch.from_numpy(ids).to(dev), off)
            o.backward(g)
            outs.append(o.detach())
        assert torch.allclose(outs[0], outs[1])
        for t in range(len(dims)):
            a, b = m.export_keys_values(t), m2.export_keys_values(t)
            ia, ib = torch.argsort(a[0]), torch.argsort(b[0])
            assert torch.equal(a[0][ia], b[0][ib]) and torch.allclose(a[1][ia][:, :dims[t]], b[1][ib][:, :dims[t]], rtol=1e-5, atol=1e-5)


CPU = torch.device("cpu")


@pytest.mark.parametrize("optimizer_name,mean", [("sgd", False), ("sgd", True), ("adagrad", False), ("adam", True)])
def test_mixed_dims_pooled_cpu_shim(optimizer_name, mean, tmp_path):
    with patched_module():
        scenario_mixed_dims(CPU, optimizer_name, mean, tmp_path)


def test_mixed_dims_rules_cpu_shim():
    from dynamicemb import DynamicEmbPoolingMode, EmbOptimType
    with patched_module():
        with pytest.raises(NotImplementedError):
            _mixed_module(CPU, [32, 64], [0, 1], EmbOptimType.EXACT_ROWWISE_ADAGRAD)
        from dynamicemb import BatchedDynamicEmbeddingTablesV2, DynamicEmbTableOptions
        with pytest.raises(NotImplementedError):
            BatchedDynamicEmbeddingTablesV2([DynamicEmbTableOptions(dim=32, max_capacity=256), DynamicEmbTableOptions(dim=64, max_capacity=256)],
                                            pooling_mode=DynamicEmbPoolingMode.NONE, device=CPU)
