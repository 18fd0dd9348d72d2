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


CONFIGS = [(True, None, False), (False, None, False), (True, 2, False), (True, None, True), (True, 2, True)]      # (dedup, threshold, cached)


def _worker(rank, world, port, q):
    """One process group, every configuration in turn (a spawn costs ~10 s of imports)."""
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "recsys-examples_b200"))
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        for dedup, threshold, cached in CONFIGS:
            _run_config(rank, world, dedup, threshold, cached)
        q.put((rank, "ok"))
    except Exception as e:      # noqa: BLE001
        import traceback
        q.put((rank, f"{type(e).__name__}: {e}\n{traceback.format_exc()}"))
    finally:
        dist.destroy_process_group()


def _run_config(rank, world, dedup, threshold, cached):
    tag = f"dedup={dedup} threshold={threshold} cached={cached}: "
    if True:
        from tests.cpu_ext_shim import patched_module
        from tests.test_admission_cpu import _module
        from dynamicemb.shard import RowWiseShardedDynamicEmbeddingA2A
        T, B, D = 2, 4, 8
        with patched_module():
            local = _module({"fused_prefetch": False}, threshold, T=T, dim=D, caching=cached, local_hbm=1024 * D * 4 if cached else 0)
            assert (local.cache is not None) == cached
            local.train()
            sharded = RowWiseShardedDynamicEmbeddingA2A(local, dist.group.WORLD, dist_type="roundrobin", use_index_dedup=dedup)
            # every rank can compute every rank's batch (same generator), so the global occurrence counts are known everywhere
            batches = []
            for r in range(world):
                g = np.random.default_rng(7 + r)
                lengths = g.integers(0, 5, size=T * B).astype(np.int64)
                ids = g.integers(1, 40, size=int(lengths.sum())).astype(np.int64)
                batches.append((lengths, ids))
            lengths, ids = batches[rank]
            feat = np.repeat(np.arange(T * B) // B, lengths)
            count = {}
            for ln, idv in batches:
                f_ = np.repeat(np.arange(T * B) // B, ln)
                for f, k in zip(f_.tolist(), idv.tolist()):
                    count[(f, k)] = count.get((f, k), 0) + 1
            x, ln = torch.from_numpy(ids), torch.from_numpy(lengths)
            init = torch.from_numpy((ids % 100000).astype(np.float32))[:, None].expand(-1, D)
            moved = torch.tensor([count[(f, k)] for f, k in zip(feat.tolist(), ids.tolist())], dtype=torch.float32)[:, None]
            # step 1: every row is at its initial value (stored or not)
            out = sharded(x, ln)
            assert torch.equal(out, init), tag + "step 1 output"
            out.backward(torch.ones_like(out))
            # step 2, same batch
            out = sharded(x, ln)
            if threshold is None:
                want = init - LR * moved                          # step 1 trained every key with the gradients of all ranks
            else:
                want = init                                       # threshold 2: step 1 only counted; nothing was stored or trained
            assert torch.allclose(out, want), tag + "step 2 output"
            out.backward(torch.ones_like(out))
            # step 3: with admission the keys were stored at step 2 (untrained copy of the initializer) and trained once
            out = sharded(x, ln)
            want = init - LR * moved * (2 if threshold is None else 1)
            assert torch.allclose(out, want), tag + "step 3 output"
            out.backward(torch.zeros_like(out))
            # ownership: roundrobin => this rank's table holds exactly the keys with key % world == rank
            for t in range(T):
                keys, _ = local.export_keys_values(t)
                allk = {k for (f, k) in count if f == t}
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
