"""GPU: the reference-named ops of the pybind module `dynamicemb_extensions` that the reference's own Python calls (module_bind.cu:22-43),
exported with the reference's argument order over the C ABI: flagged_compact, load/store_*_flat_table_*, *_update_for_flat_table,
bucketize_keys, the initializer ops.  Each alias is checked against plain torch / numpy restatements of the reference kernels
(dynamic_emb_op.cu:295-490, optimizer_kernel.cuh, index_calculation.cu:130, bucketize.cu:38-58,111, initializer.cuh)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _tables(cuda):
    g = torch.Generator(device=cuda).manual_seed(0)
    dims = [(8, 16), (16, 48), (4, 4)]               # (emb, value) per table: Adagrad-like, Adam-like, SGD (no state)
    tabs = [torch.randn(40 + 10 * t, v, device=cuda, generator=g) for t, (e, v) in enumerate(dims)]
    ptrs = torch.tensor([t.data_ptr() for t in tabs], dtype=torch.int64, device=cuda)
    vd = torch.tensor([v for _, v in dims], dtype=torch.int64, device=cuda)
    ed = torch.tensor([e for e, _ in dims], dtype=torch.int64, device=cuda)
    return tabs, ptrs, vd, ed, dims


def test_flagged_compact(cuda):
    from dynamicemb import dynamicemb_extensions as ext
    g = torch.Generator(device=cuda).manual_seed(1)
    for n in (0, 1, 31, 2048, 2049, 100_003):
        flags = torch.rand(n, device=cuda, generator=g) < 0.37
        a = torch.randint(-1 << 60, 1 << 60, (n,), device=cuda, generator=g)
        b = torch.arange(n, device=cuda) * 3
        cnt, idx, (oa, none, ob) = ext.flagged_compact(flags, [a, None, b])
        want = flags.nonzero().flatten()
        assert cnt == want.numel() and none is None
        assert torch.equal(idx, want) and torch.equal(oa, a[want]) and torch.equal(ob, b[want])
    c, i, _ = ext.flagged_compact(torch.ones(5000, dtype=torch.bool, device=cuda), [])
    assert c == 5000 and torch.equal(i, torch.arange(5000, device=cuda))


def test_load_store_flat_table_mixed_dims(cuda):
    from dynamicemb import dynamicemb_extensions as ext
    tabs, ptrs, vd, ed, dims = _tables(cuda)
    g = torch.Generator(device=cuda).manual_seed(2)
    n, max_e = 64, 16
    tids = torch.randint(0, 3, (n,), device=cuda, generator=g)
    idx = torch.stack([torch.randint(0, tabs[int(t)].shape[0], (1,), device=cuda, generator=g)[0] for t in tids])
    idx[5] = -1
    # emb
    out = torch.full((n, max_e), 7.0, device=cuda)
    ext.load_from_flat_table_emb(ptrs, idx, tids, out, vd, ed, max_e, True)
    for i in range(n):
        e = dims[int(tids[i])][0]
        if int(idx[i]) < 0:
            assert bool((out[i] == 7).all())
        else:
            assert torch.equal(out[i, :e], tabs[int(tids[i])][int(idx[i]), :e]) and bool((out[i, e:] == 7).all())
    # value: [emb | pad to max_e | state]
    width = max_e + max(v - e for e, v in dims)
    out = torch.zeros(n, width, device=cuda)
    ext.load_from_flat_table_value(ptrs, idx, tids, out, vd, ed, max_e, True)
    for i in range(n):
        if int(idx[i]) < 0:
            continue
        e, v = dims[int(tids[i])]
        row = tabs[int(tids[i])][int(idx[i])]
        assert torch.equal(out[i, :e], row[:e]) and torch.equal(out[i, max_e:max_e + v - e], row[e:])
    # store value back after a change, then contiguous load of one table
    out2 = out * 2
    uniq = {}
    for i in range(n):                                   # keep one writer per row
        uniq[(int(tids[i]), int(idx[i]))] = i
    keep = torch.tensor(sorted(uniq.values()), device=cuda)
    ext.store_to_flat_table_value(ptrs, idx[keep], tids[keep], out2[keep].contiguous(), vd, ed, max_e, True)
    for i in keep.tolist():
        if int(idx[i]) < 0:
            continue
        e, v = dims[int(tids[i])]
        row = tabs[int(tids[i])][int(idx[i])]
        assert torch.equal(row[:e], out2[i, :e]) and torch.equal(row[e:], out2[i, max_e:max_e + v - e])
    t1 = torch.arange(0, 20, device=cuda)
    dense = torch.zeros(20, dims[1][1], device=cuda)
    ext.load_from_flat_table_contiguous(ptrs, t1, 1, dense, vd, ed, max_e, True)
    assert torch.equal(dense, tabs[1][:20])
    ext.store_to_flat_table_contiguous(ptrs, t1, 1, (dense + 1).contiguous(), vd, ed, max_e, True)
    assert torch.equal(tabs[1][:20], dense + 1)


@pytest.mark.parametrize("opt", ["sgd", "adagrad", "adam", "rowwise"])
def test_update_for_flat_table(cuda, opt):
    from dynamicemb import dynamicemb_extensions as ext
    g = torch.Generator(device=cuda).manual_seed(3)
    e = [8, 16]
    state = {"sgd": [0, 0], "adagrad": e, "adam": [2 * x for x in e], "rowwise": [4, 4]}[opt]
    tabs = [torch.rand(30, e[t] + state[t], device=cuda, generator=g) + 0.1 for t in range(2)]
    ptrs = torch.tensor([t.data_ptr() for t in tabs], dtype=torch.int64, device=cuda)
    vd = torch.tensor([e[t] + state[t] for t in range(2)], dtype=torch.int64, device=cuda)
    ed = torch.tensor(e, dtype=torch.int64, device=cuda)
    n, max_e = 20, 16
    tids = torch.tensor([i % 2 for i in range(n)], device=cuda)
    idx = torch.arange(n, device=cuda)                   # distinct rows
    idx[3] = -1
    grads = torch.randn(n, max_e, device=cuda, generator=g)
    before = [t.clone() for t in tabs]
    lr, eps, b1, b2, wd, it = 0.1, 1e-8, 0.9, 0.999, 0.01, 3
    if opt == "sgd":
        ext.sgd_update_for_flat_table(grads, idx, ptrs, tids, vd, ed, max_e, True, lr, 0)
    elif opt == "adagrad":
        ext.adagrad_update_for_flat_table(grads, idx, ptrs, tids, vd, ed, lr, eps, max_e, True, 0)
    elif opt == "adam":
        ext.adam_update_for_flat_table(grads, idx, ptrs, tids, vd, ed, lr, b1, b2, eps, wd, it, max_e, True, 0)
    else:
        ext.rowwise_adagrad_for_flat_table(grads, idx, ptrs, tids, vd, ed, lr, eps, max_e, True, 0)
    for i in range(n):
        t, r = int(tids[i]), int(idx[i])
        if r < 0:
            continue
        D = e[t]
        gi = grads[i, :D]
        w, s = before[t][r, :D].clone(), before[t][r, D:].clone()
        if opt == "sgd":
            w -= lr * gi
        elif opt == "adagrad":
            s = s + gi * gi
            w -= lr * gi / (s.sqrt() + eps)
        elif opt == "adam":
            m, v = s[:D] * b1 + (1 - b1) * gi, s[D:] * b2 + (1 - b2) * gi * gi
            w -= lr * ((m / (1 - b1 ** it)) / ((v / (1 - b2 ** it)).sqrt() + eps) + wd * w)
            s = torch.cat([m, v])
        else:
            acc = s[0] + (gi * gi).mean()
            w -= lr * gi / (acc.sqrt() + eps)
            s = s.clone(); s[0] = acc
        torch.testing.assert_close(tabs[t][r, :D], w, rtol=2e-5, atol=2e-6)
        if opt == "rowwise":
            torch.testing.assert_close(tabs[t][r, D], s[0], rtol=2e-5, atol=2e-6)
        elif opt != "sgd":
            torch.testing.assert_close(tabs[t][r, D:], s, rtol=2e-5, atol=2e-6)
    untouched = torch.ones(30, dtype=torch.bool, device=cuda)
    untouched[idx[idx >= 0][tids[idx >= 0] == 0]] = False
    assert torch.equal(tabs[0][untouched], before[0][untouched])


def test_bucketize_keys(cuda):
    from dynamicemb import dynamicemb_extensions as ext
    from oracle.dynamicemb import hash63
    rng = np.random.default_rng(4)
    C = 128
    bkt_off = np.array([0, 5, 5, 12], dtype=np.int64)                # table 1 is empty
    keys = rng.integers(-(1 << 62), 1 << 62, size=700, dtype=np.int64)
    tids = rng.choice([0, 2], size=700).astype(np.int64)
    ko, offsets, inv = ext.bucketize_keys(torch.from_numpy(keys).to(cuda), torch.from_numpy(tids).to(cuda), torch.from_numpy(bkt_off).to(cuda), 12, C)
    bucket = np.array([bkt_off[t] + (hash63(int(k)) % ((bkt_off[t + 1] - bkt_off[t]) * C)) // C for k, t in zip(keys, tids)])
    order = np.lexsort((keys, bucket))                               # by bucket, then signed key (scored_hashtable.py:1451-1557)
    assert np.array_equal(inv.cpu().numpy(), order) and np.array_equal(ko.cpu().numpy(), keys[order])
    active, counts = np.unique(bucket, return_counts=True)
    assert np.array_equal(offsets.cpu().numpy(), np.concatenate([[0], np.cumsum(counts)]))


def test_initializer_ops(cuda):
    from dynamicemb import dynamicemb_extensions as ext
    n, D = 4096, 64
    idx = torch.arange(n, device=cuda)
    idx[7] = -1
    buf = torch.full((n, D), -5.0, device=cuda)
    ext.const_init(buf, idx, 0.25)
    assert bool((buf[7] == -5).all()) and bool((buf[idx >= 0] == 0.25).all())
    keys = torch.arange(n, device=cuda) * 100_003 + 17
    ext.debug_init(buf, idx, keys)
    assert torch.equal(buf[idx >= 0][:, 0], (keys[idx >= 0] % 100000).float()) and bool((buf[7] == -5).all())
    ctx = ext.CurandStateContext(11)
    ext.uniform_init(buf, idx, ctx, -0.5, 1.5)
    x = buf[idx >= 0]
    assert float(x.min()) >= -0.5 and float(x.max()) <= 1.5 and abs(float(x.mean()) - 0.5) < 0.01 and abs(float(x.std()) - 2 / 12 ** 0.5) < 0.01
    ext.normal_init(buf, idx, ctx, 1.0, 2.0)
    x = buf[idx >= 0]
    assert abs(float(x.mean()) - 1.0) < 0.02 and abs(float(x.std()) - 2.0) < 0.02
    ext.truncated_normal_init(buf, idx, ctx, 0.0, 1.0, -1.0, 1.0)
    x = buf[idx >= 0]
    assert float(x.min()) >= -1.0 and float(x.max()) <= 1.0 and abs(float(x.mean())) < 0.02
    buf2 = torch.empty_like(buf)
    ext.truncated_normal_init(buf2, idx, ctx, 0.0, 1.0, -1.0, 1.0)                # counter-based: same seed, same rows
    assert torch.equal(buf2[idx >= 0], x)
