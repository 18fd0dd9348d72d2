"""world_size-2 gloo test of the SHARDED MODULE host path: `RowWiseShardedDynamicEmbeddingA2A` around a real
`BatchedDynamicEmbeddingTablesV2` per rank (bucketize -> all_to_all -> the module's op-by-op prefetch / forward / fused backward ->
all_to_all of rows and of row gradients), with the op layer replaced by the oracle-backed CPU shim (tests/cpu_ext_shim.py).

The check is the reference's sharded-vs-unsharded check (corelib/dynamicemb/test/unit_tests/test_sequence_embedding_fw.py) in closed
form: with the DEBUG initializer a row starts at key % 100000 and SGD moves it by -lr * (number of occurrences of the key over ALL ranks),
whatever rank owns it — so outputs of step 2 prove that ids, rows and gradients crossed the ranks correctly.  Further variants put an
admission strategy on the shards (a key is stored by its owner only at its second presentation) and / or make every shard a cached
module (HBM cache over a host-resident table).
"""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests.test_dist_cpu import ROOT, _free_port

LR = 0.5


CONFIGS = [(True, None, False, "none"), (False, None, False, "none"), (True, 2, False, "none"), (True, None, True, "none"), (True, 2, True, "none"),
           (True, None, False, "sum"), (True, 2, True, "mean")]      # (dedup, admission threshold, cached shard, pooling at the wrapper)


def _worker(rank, world, port, q):
    """One process group, every configuration in turn (a spawn costs ~10 s of imports)."""
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "recsys-examples_b200"))
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        for cfg in CONFIGS:
            _run_config(rank, world, *cfg)
        q.put((rank, "ok"))
    except Exception as e:      # noqa: BLE001
        import traceback
        q.put((rank, f"{type(e).__name__}: {e}\n{traceback.format_exc()}"))
    finally:
        dist.destroy_process_group()


def _run_config(rank, world, dedup, threshold, cached, pooling):
    from tests.cpu_ext_shim import patched_module
    from tests.test_admission_cpu import _module
    from dynamicemb import DynamicEmbPoolingMode as P
    from dynamicemb.shard import RowWiseShardedDynamicEmbeddingA2A
    tag = f"dedup={dedup} threshold={threshold} cached={cached} pooling={pooling}: "
    T, B, D = 2, 4, 8
    with patched_module():
        local = _module({"fused_prefetch": False}, threshold, T=T, dim=D, caching=cached, local_hbm=1024 * D * 4 if cached else 0)
        assert (local.cache is not None) == cached
        local.train()
        sharded = RowWiseShardedDynamicEmbeddingA2A(local, dist.group.WORLD, dist_type="roundrobin", use_index_dedup=dedup,
                                                    pooling_mode={"none": P.NONE, "sum": P.SUM, "mean": P.MEAN}[pooling])
        # every rank can compute every rank's batch (same generator), so the global per-key gradient weights are known everywhere
        batches = []
        for r in range(world):
            g = np.random.default_rng(7 + r)
            lengths = g.integers(0, 5, size=T * B).astype(np.int64)
            ids = g.integers(1, 40, size=int(lengths.sum())).astype(np.int64)
            batches.append((lengths, ids))

        def per_id(lengths):        # (feature, bag, pooling weight) of every id: the weight is also its gradient under d(out) = ones
            bag = np.repeat(np.arange(T * B), lengths)
            w = np.ones(bag.size) if pooling != "mean" else 1.0 / np.repeat(lengths, lengths)
            return bag // B, bag, w

        moved = {}
        for ln_, idv in batches:
            f_, _, w_ = per_id(ln_)
            for f, k, w in zip(f_.tolist(), idv.tolist(), w_.tolist()):
                moved[(f, k)] = moved.get((f, k), 0.0) + w
        lengths, ids = batches[rank]
        feat, bag, w = per_id(lengths)
        x, ln = torch.from_numpy(ids), torch.from_numpy(lengths)

        def expected(trained_steps):
            val = np.array([(k % 100000) - LR * trained_steps * moved[(f, k)] for f, k in zip(feat.tolist(), ids.tolist())])
            if pooling == "none":
                return torch.from_numpy(np.repeat(val[:, None], D, axis=1)).float()
            out = np.zeros((B, T * D))
            for v, g_, wi in zip(val.tolist(), bag.tolist(), w.tolist()):
                f, b = g_ // B, g_ % B
                out[b, f * D:(f + 1) * D] += v * wi
            return torch.from_numpy(out).float()

        # with a threshold of 2 the first step only counts the keys: nothing is stored or trained before step 2
        trained_before = [0, 1, 2] if threshold is None else [0, 0, 1]
        for step, t_ in enumerate(trained_before, 1):
            out = sharded(x, ln)
            assert torch.allclose(out, expected(t_), rtol=1e-5, atol=1e-4), tag + f"step {step} output"
            out.backward(torch.ones_like(out) if step < 3 else torch.zeros_like(out))
        # ownership: roundrobin => this rank's table holds exactly the keys with key % world == rank
        for t in range(T):
            keys, _ = local.export_keys_values(t)
            allk = {k for (f, k) in moved if f == t}
            assert set(keys.tolist()) == {k for k in allk if k % world == rank}, tag + f"table {t} ownership"


def test_sharded_module_matches_closed_form_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for r, msg in res:
        assert msg == "ok", f"rank {r}: {msg}"


class _Cfg:
    """What the sharder reads of a TorchRec EmbeddingConfig."""

    def __init__(self, name, dim, num):
        self.name, self.embedding_dim, self.num_embeddings, self.feature_names = name, dim, num, [f"f_{name}"]


class _Collection:
    def __init__(self, cfgs):
        self._cfgs = cfgs

    def embedding_configs(self):
        return self._cfgs

    def embedding_bag_configs(self):
        return self._cfgs


class _Env:
    def __init__(self, pg):
        self.process_group = pg


def _worker_planner(rank, world, port, q):
    """The user-level flow of the reference (apply_dmp: constraints -> planner.plan -> sharder.shard -> sharded module) for a table with
    an admission strategy: the sharder hands such a shard to the all_to_all wrapper."""
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "recsys-examples_b200"))
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from tests.cpu_ext_shim import patched_module
        from dynamicemb import (DynamicEmbInitializerArgs, DynamicEmbInitializerMode, DynamicEmbScoreStrategy, DynamicEmbTableOptions, EmbOptimType,
                                FrequencyAdmissionStrategy, KVCounter, get_sharded_table_capacity)
        from dynamicemb.shard import (DynamicEmbeddingCollectionSharder, DynamicEmbeddingShardingPlanner, DynamicEmbParameterConstraints,
                                      RowWiseShardedDynamicEmbeddingA2A)
        D = 8
        with patched_module():
            opt = DynamicEmbTableOptions(score_strategy=DynamicEmbScoreStrategy.STEP, dist_type="roundrobin",
                                         initializer_args=DynamicEmbInitializerArgs(mode=DynamicEmbInitializerMode.DEBUG),
                                         admit_strategy=FrequencyAdmissionStrategy(2), admission_counter=KVCounter(4096, bucket_capacity=128))
            cons = {"item": DynamicEmbParameterConstraints(use_dynamicemb=True, dynamicemb_options=opt)}
            cfgs = [_Cfg("item", D, 10_000)]
            plan = DynamicEmbeddingShardingPlanner(eb_configs=cfgs, constraints=cons, world_size=world).plan()
            assert plan["item"]["local_capacity"] == get_sharded_table_capacity(10_000, world, 128)
            sharder = DynamicEmbeddingCollectionSharder(use_index_dedup=True, fused_params={"optimizer": EmbOptimType.SGD, "learning_rate": LR,
                                                                                            "fused_prefetch": False})
            sharded = sharder.shard(_Collection(cfgs), plan, env=_Env(dist.group.WORLD), device=torch.device("cpu"))
            assert isinstance(sharded, RowWiseShardedDynamicEmbeddingA2A) and sharded.local.tables.capacity() == plan["item"]["local_capacity"]
            sharded.local.train()
            ids = torch.tensor([3 + rank, 10, 11, 10], dtype=torch.int64)             # 10 and 11 are asked by both ranks
            ln = torch.tensor([4], dtype=torch.int64)
            init = ids.to(torch.float32)[:, None].expand(-1, D)
            for step, want in ((1, init), (2, init)):                                  # counted, then stored (untrained) at step 2
                out = sharded(ids, ln)
                assert torch.equal(out, want), f"step {step}"
                out.backward(torch.ones_like(out))
            out = sharded(ids, ln)                                                     # step 2's gradients: both ranks' occurrences
            occ = torch.tensor([1.0, 4.0, 2.0, 4.0])[:, None]
            assert torch.allclose(out, init - LR * occ), "step 3"
            out.backward(torch.zeros_like(out))
            # the bag-collection sharder on a cached table: a pooling all_to_all wrapper around a sequence-mode cached shard
            from dynamicemb import DynamicEmbPoolingMode
            from dynamicemb.shard import DynamicEmbeddingBagCollectionSharder
            opt2 = DynamicEmbTableOptions(score_strategy=DynamicEmbScoreStrategy.STEP, dist_type="roundrobin", caching=True, local_hbm_for_values=1024 * D * 4,
                                          initializer_args=DynamicEmbInitializerArgs(mode=DynamicEmbInitializerMode.DEBUG))
            cons2 = {"item": DynamicEmbParameterConstraints(use_dynamicemb=True, dynamicemb_options=opt2)}
            plan2 = DynamicEmbeddingShardingPlanner(eb_configs=cfgs, constraints=cons2, world_size=world).plan()
            bag = DynamicEmbeddingBagCollectionSharder(fused_params={"optimizer": EmbOptimType.SGD, "learning_rate": LR, "fused_prefetch": False}) \
                .shard(_Collection(cfgs), plan2, env=_Env(dist.group.WORLD), device=torch.device("cpu"))
            assert isinstance(bag, RowWiseShardedDynamicEmbeddingA2A) and bag.pooling_mode == DynamicEmbPoolingMode.SUM and bag.local.cache is not None
            bag.local.train()
            out = bag(ids, torch.tensor([3, 1], dtype=torch.int64))                      # two bags: [3+rank, 10, 11] and [10]
            assert out.shape == (2, D) and out[:, 0].tolist() == [float(3 + rank + 10 + 11), 10.0]
            out.backward(torch.ones_like(out))
        q.put((rank, "ok"))
    except Exception as e:      # noqa: BLE001
        import traceback
        q.put((rank, f"{type(e).__name__}: {e}\n{traceback.format_exc()}"))
    finally:
        dist.destroy_process_group()


def test_planner_sharder_flow_with_admission_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_planner, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for r, msg in res:
        assert msg == "ok", f"rank {r}: {msg}"
