"""GPU, world_size 2 (NCCL over NVLink): RowWiseShardedDynamicEmbedding against the UNSHARDED module on the same ids — the reference's
own multi-GPU check (corelib/dynamicemb/test/unit_tests/test_sequence_embedding_fw.py, test_pooled_embedding_fw.py: sharded lookup vs
a plain dict of rows).  Each rank feeds its own batch; forward rows must equal what one unsharded table holding every key returns, and
after a training step the rows of every key (wherever they live) must equal the unsharded module's rows after the same step with the
union of all ranks' gradients.  Skipped on boxes with fewer than 2 GPUs (run with `gpurun --gpus 2`; evidence: profiles/)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
D = 128


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _mk(dev, cap, pooling, opt, lr):
    from dynamicemb import (BatchedDynamicEmbeddingTablesV2, DynamicEmbInitializerArgs, DynamicEmbInitializerMode, DynamicEmbScoreStrategy,
                            DynamicEmbTableOptions)
    o = DynamicEmbTableOptions(dim=D, max_capacity=cap, local_hbm_for_values=1 << 50, score_strategy=DynamicEmbScoreStrategy.STEP,
                               initializer_args=DynamicEmbInitializerArgs(mode=DynamicEmbInitializerMode.DEBUG))
    return BatchedDynamicEmbeddingTablesV2([o, o], table_names=["a", "b"], feature_table_map=[0, 1], pooling_mode=pooling, optimizer=opt,
                                           learning_rate=lr, eps=1e-8, device=dev)


def _rows_by_key(m):
    out = {}
    for t in range(2):
        keys, vals = m.export_keys_values(t)
        for k, v in zip(keys.tolist(), vals.cpu()):
            out[(t, k)] = v
    return out


def _worker(rank, world, port, pooled, dedup, dist_type, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "recsys-examples_b200"))
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from dynamicemb import DynamicEmbPoolingMode, EmbOptimType
        from dynamicemb.shard import RowWiseShardedDynamicEmbedding
        pm = DynamicEmbPoolingMode.SUM if pooled else DynamicEmbPoolingMode.NONE
        lr = 0.05
        F, B = 2, 257
        local = _mk(dev, 1 << 16, pm, EmbOptimType.EXACT_ADAGRAD, lr)
        local.train()
        model = RowWiseShardedDynamicEmbedding(local, None, dist_type=dist_type, use_index_dedup=dedup, num_embeddings_per_feature=[1 << 20] * F)
        ref = _mk(dev, 1 << 17, pm, EmbOptimType.EXACT_ADAGRAD, lr)       # unsharded: every rank replays the union of all batches
        ref.train()
        for step in range(4):
            rng = np.random.default_rng(1000 * step + rank)
            lengths_np = rng.integers(0, 12, size=F * B).astype(np.int64)
            ids_np = (rng.zipf(1.2, size=int(lengths_np.sum())) % 5000).astype(np.int64) * 7 + 3
            ids, lengths = torch.from_numpy(ids_np).to(dev), torch.from_numpy(lengths_np).to(dev)
            out = model(ids, lengths)
            # ---- the unsharded run of the union: KJT of world*B samples per feature (rank-major inside a feature)
            everyone = [None] * world
            dist.all_gather_object(everyone, (ids_np, lengths_np))
            u_ids, u_len = [], []
            for f in range(F):
                for i_r, l_r in everyone:
                    o_r = np.concatenate([[0], np.cumsum(l_r)])
                    u_ids.append(i_r[o_r[f * B]: o_r[(f + 1) * B]])
                    u_len.append(l_r[f * B:(f + 1) * B])
            u_ids, u_len = np.concatenate(u_ids), np.concatenate(u_len)
            u_off = torch.from_numpy(np.concatenate([[0], np.cumsum(u_len)]).astype(np.int64)).to(dev)
            ref_out = ref(torch.from_numpy(u_ids).to(dev), u_off)
            # my slice of the unsharded output
            if pooled:
                mine = ref_out[rank * B:(rank + 1) * B]                       # [W*B, F*D] -> my B samples
            else:
                pos, cur = [], 0
                for f in range(F):
                    for r, (i_r, l_r) in enumerate(everyone):
                        cnt = int(l_r[f * B:(f + 1) * B].sum())
                        if r == rank:
                            pos.append(np.arange(cur, cur + cnt))
                        cur += cnt
                mine = ref_out[torch.from_numpy(np.concatenate(pos)).to(dev)]
            assert out.shape == mine.shape, f"step {step}: {out.shape} vs {mine.shape}"
            if step == 0:       # nothing trained yet: pure copies (sequence) / identical accumulation order (pooled) -> bit-exact
                assert torch.equal(out.detach(), mine.detach()), f"step {step}: sharded forward differs from the unsharded module"
            else:               # trained rows: the sharded path sums a key's gradient per rank first, then across ranks (fp32 order differs)
                torch.testing.assert_close(out.detach(), mine.detach(), rtol=2e-5, atol=2e-5, msg=f"step {step}: sharded forward")
            # ---- one training step with per-rank gradients; the unsharded module gets the union
            g = torch.randn(out.shape, device=dev, generator=torch.Generator(device=dev).manual_seed(77 * step + rank))
            out.backward(g)
            if pooled:
                allg = [torch.empty(B, F * D, device=dev) for _ in range(world)]
                dist.all_gather(allg, g.contiguous())
                ref_out.backward(torch.cat(allg, 0))
            else:
                sizes = [int(l_r.sum()) for _, l_r in everyone]
                allg = [torch.empty(s, D, device=dev) for s in sizes]
                dist.all_gather(allg, g.contiguous())
                ug, cur_r = [], [0] * world
                for f in range(F):
                    for r, (i_r, l_r) in enumerate(everyone):
                        cnt = int(l_r[f * B:(f + 1) * B].sum())
                        ug.append(allg[r][cur_r[r]: cur_r[r] + cnt])
                        cur_r[r] += cnt
                ref_out.backward(torch.cat(ug, 0))
        # ---- rows after training: every key this rank owns == the unsharded module's row (fp32 sums in a different order: tolerance)
        mine_rows = _rows_by_key(local)
        ref_rows = _rows_by_key(ref)
        owned = [None] * world
        dist.all_gather_object(owned, sorted(mine_rows.keys()))
        allk = sorted(k for o in owned for k in o)
        assert allk == sorted(ref_rows.keys()), "the shards together do not hold exactly the keys of the unsharded table"
        assert len(set(allk)) == len(allk), "a key lives on two ranks"
        for k, v in mine_rows.items():
            torch.testing.assert_close(v, ref_rows[k], rtol=2e-5, atol=2e-5)
        model.check()
        q.put((rank, "ok"))
    except Exception:  # pragma: no cover
        import traceback
        q.put((rank, traceback.format_exc()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("pooled,dedup,dist_type", [(False, True, "hash_roundrobin"), (False, False, "roundrobin"), (True, False, "hash_roundrobin")])
def test_sharded_matches_unsharded_nccl(pooled, dedup, dist_type):
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, pooled, dedup, dist_type, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=600) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for r, msg in res:
        assert msg == "ok", f"rank {r}: {msg}"


def test_sharded_world1_matches_unsharded(cuda):
    """The same kernels with a process group of ONE rank (every id routes to the local shard through the symmetric buffer): exchange
    bookkeeping, device-side counts, owner prefetch with a device count, peer-address gather / gradient stores, CUDA-graph step — on a
    single GPU, so it runs wherever the GPU suite runs."""
    from dynamicemb import DynamicEmbPoolingMode, EmbOptimType
    from dynamicemb.shard import RowWiseShardedDynamicEmbedding
    created = False
    if not dist.is_initialized():
        dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{_free_port()}", rank=0, world_size=1, device_id=cuda)
        created = True
    try:
        for pm in (DynamicEmbPoolingMode.NONE, DynamicEmbPoolingMode.MEAN):
            lr, F, B = 0.05, 2, 300
            local = _mk(cuda, 1 << 16, pm, EmbOptimType.EXACT_ADAGRAD, lr)
            ref = _mk(cuda, 1 << 16, pm, EmbOptimType.EXACT_ADAGRAD, lr)
            local.train(); ref.train()
            model = RowWiseShardedDynamicEmbedding(local, None, dist_type="hash_roundrobin", num_embeddings_per_feature=[1 << 20] * F, max_ids_per_step=8192)
            for step in range(5):
                rng = np.random.default_rng(step)
                lengths_np = rng.integers(0, 14, size=F * B).astype(np.int64)
                ids_np = (rng.zipf(1.2, size=int(lengths_np.sum())) % 5000).astype(np.int64) * 7 + 3
                ids, lengths = torch.from_numpy(ids_np).to(cuda), torch.from_numpy(lengths_np).to(cuda)
                off = torch.from_numpy(np.concatenate([[0], np.cumsum(lengths_np)]).astype(np.int64)).to(cuda)
                out, rout = model(ids, lengths), ref(ids, off)
                assert torch.equal(out.detach(), rout.detach()), f"{pm} step {step}: forward"     # one rank: same reduction order everywhere
                g = torch.randn(out.shape, device=cuda, generator=torch.Generator(device=cuda).manual_seed(step))
                out.backward(g); rout.backward(g)
                model.check()
                assert torch.equal(local.tables.table_storage_, ref.tables.table_storage_), f"{pm} step {step}: table image"
                torch.testing.assert_close(local._values, ref._values, rtol=1e-6, atol=1e-6)
            assert int(local.tables._ref_counter.abs().sum().item()) == 0
        # CUDA-graph step of the sharded wrapper == its eager step
        la, lb = (_mk(cuda, 1 << 16, DynamicEmbPoolingMode.NONE, EmbOptimType.EXACT_ADAGRAD, 0.05) for _ in range(2))
        la.train(); lb.train()
        ma = RowWiseShardedDynamicEmbedding(la, None, max_ids_per_step=4096)
        mb = RowWiseShardedDynamicEmbedding(lb, None, max_ids_per_step=4096)
        n, F = 3000, 2
        lengths = torch.full((F * 300,), n // (F * 300), dtype=torch.int64, device=cuda)
        ids_static = torch.zeros(n, dtype=torch.int64, device=cuda)
        grad = torch.randn(n, D, device=cuda)
        rng = np.random.default_rng(3)
        batches = [torch.from_numpy((rng.zipf(1.1, size=n) % 30000).astype(np.int64) * 31).to(cuda) for _ in range(8)]
        ids_static.copy_(batches[0])
        graph, out, loss = ma.make_graphed_step(ids_static, lengths, grad)
        for _ in range(3):
            o = mb(batches[0], lengths); o.backward(grad)
        for b in batches[1:]:
            ids_static.copy_(b)
            graph.replay()
            o = mb(b, lengths)
            l = o.detach().sum()
            o.backward(grad)
            assert torch.equal(out, o) and torch.equal(loss, l)
            assert torch.equal(la.tables.table_storage_, lb.tables.table_storage_) and torch.equal(la._values, lb._values)
        ma.check(); mb.check()
        # ---- checkpoint through the wrapper (dynamicemb/checkpoint.py): every rank dumps its own shard under the process group's rank /
        # world size, the meta file carries the WRAPPER's dist_type; a fresh wrapper loads it back to identical lookups
        import json
        import tempfile
        from dynamicemb import checkpoint as ck
        with tempfile.TemporaryDirectory() as ckdir:
            mb.dump(ckdir, optim=True)
            name = lb.table_names[0]
            assert json.load(open(ck.encode_meta_json_file_path(ckdir, name)))["dist_type"] == "hash_roundrobin"
            for item in ("keys", "values", "scores", "opt_values"):
                assert os.path.getsize(ck.encode_checkpoint_file_path(ckdir, name, 0, 1, item)) > 0
            lc = _mk(cuda, 1 << 17, DynamicEmbPoolingMode.NONE, EmbOptimType.EXACT_ADAGRAD, 0.05)      # another capacity: other slots
            mc = RowWiseShardedDynamicEmbedding(lc, None, max_ids_per_step=4096)
            mc.load(ckdir, optim=True)
            lc.train()
            with torch.no_grad():
                for b in batches[:3]:
                    assert torch.equal(mb(b, lengths), mc(b, lengths))
            ka, va = lb.export_keys_values(0)
            kc, vc = lc.export_keys_values(0)
            oa, oc = torch.argsort(ka), torch.argsort(kc)
            assert torch.equal(ka[oa], kc[oc]) and torch.equal(va[oa], vc[oc])            # embeddings AND Adagrad accumulators
            lo = _mk(cuda, 1 << 16, DynamicEmbPoolingMode.NONE, EmbOptimType.EXACT_ADAGRAD, 0.05)
            with pytest.raises(ValueError, match="dist_type mismatch"):                      # an unsharded module defaults to roundrobin
                lo.load(ckdir)
        # ---- planner + sharder classes (reference names: planner/planner.py:213, shard/embedding.py:343, shard/embeddingbag.py:79): a row-wise
        # plan for two tables, sharder.shard() builds the rank's module from the (duck-typed) TorchRec configs and returns the wrapper
        from dynamicemb import DynamicEmbInitializerArgs, DynamicEmbInitializerMode, DynamicEmbScoreStrategy, DynamicEmbTableOptions
        from dynamicemb.shard import (DynamicEmbeddingBagCollectionSharder, DynamicEmbeddingCollectionSharder, DynamicEmbeddingShardingPlanner,
                                      DynamicEmbParameterConstraints)

        class _Cfg:
            def __init__(self, name, dim, num, feats, pooling="SUM"):
                self.name, self.embedding_dim, self.num_embeddings, self.feature_names, self.pooling = name, dim, num, feats, pooling

        class _FakeCollection:
            def __init__(self, cfgs):
                self._c = cfgs

            def embedding_configs(self):
                return self._c

            def embedding_bag_configs(self):
                return self._c

        def mk_opt():
            return DynamicEmbTableOptions(score_strategy=DynamicEmbScoreStrategy.STEP, dist_type="hash_roundrobin",
                                          initializer_args=DynamicEmbInitializerArgs(mode=DynamicEmbInitializerMode.DEBUG))
        cons = {"user": DynamicEmbParameterConstraints(use_dynamicemb=True, dynamicemb_options=mk_opt()),
                "item": DynamicEmbParameterConstraints(use_dynamicemb=True, dynamicemb_options=mk_opt())}
        plan = DynamicEmbeddingShardingPlanner(cons, world_size=1).plan({"user": 50_000, "item": 200_000})
        assert plan["user"]["local_capacity"] == 50_048 and plan["item"]["sharding_type"] == "row_wise"      # round_up(ceil(N / W), 128)
        for nm in plan:
            plan[nm]["dynamicemb_options"] = cons[nm].dynamicemb_options
        cfgs = [_Cfg("user", D, 50_000, ["f_user"]), _Cfg("item", D, 200_000, ["f_item_a", "f_item_b"])]
        for sharder, pooled in ((DynamicEmbeddingCollectionSharder(use_index_dedup=True, fused_params={"optimizer": EmbOptimType.SGD, "learning_rate": 0.5}), False),
                                (DynamicEmbeddingBagCollectionSharder(fused_params={"optimizer": EmbOptimType.SGD, "learning_rate": 0.5}), True)):
            sh = sharder.shard(_FakeCollection(cfgs), plan, env=None, device=cuda)
            sh.local.train()
            Bq = 16
            lengths = torch.full((3 * Bq,), 2, dtype=torch.int64, device=cuda)
            idsq = torch.arange(3 * Bq * 2, device=cuda, dtype=torch.int64) % 37 + 5
            o = sh(idsq, lengths)
            assert o.shape == ((Bq, 3 * D) if pooled else (idsq.numel(), D))
            want = (idsq % 100000).float()
            if pooled:
                exp = want.view(3, Bq, 2).sum(-1).t().contiguous()           # [B, F] bag sums of the debug-initialised rows
                assert torch.equal(o[:, ::D], exp)
            else:
                assert torch.equal(o[:, 0], want)
            o.sum().backward()
            sh.check()
    finally:
        if created:
            dist.destroy_process_group()
