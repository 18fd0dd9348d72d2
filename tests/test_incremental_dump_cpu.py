"""world_size-2 gloo test of `module.incremental_dump(..., pg)`: every rank receives all ranks' rows in rank order
(reference key_value_table.py:75-111, :1977-2036).  The module runs on the oracle-backed op-layer shim (tests/cpu_ext_shim.py)."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests.test_dist_cpu import ROOT, _free_port


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "recsys-examples_b200"))
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from tests.cpu_ext_shim import patched_module
        from tests.test_admission_cpu import _module
        with patched_module():
            m = _module({"fused_prefetch": False}, None)
            m.train()
            mine = [10 * (rank + 1) + i for i in range(3 + rank)]          # rank 0: 10,11,12   rank 1: 20,21,22,23
            for step_ids in (mine[:2], mine[1:]):
                x = torch.tensor(step_ids, dtype=torch.int64)
                out = m(x, torch.arange(0, len(step_ids) + 1, dtype=torch.int64))
                out.backward(torch.zeros_like(out))
            tensors, nxt = m.incremental_dump({"t0": 2}, pg=dist.group.WORLD)      # rows touched in step 2
            keys, vals = tensors["t0"]
            want = [11, 12] + [21, 22, 23]
            assert sorted(keys[:2].tolist()) == want[:2] and sorted(keys[2:].tolist()) == want[2:], keys.tolist()
            assert vals.shape == (5, 32) and all(float(v[0]) == float(k) for k, v in zip(keys.tolist(), vals))
            assert nxt == {"t0": 3}
            # an empty rank still takes part in the gather
            tensors, _ = m.incremental_dump({"t0": 3 if rank == 0 else 0}, pg=dist.group.WORLD)
            assert sorted(tensors["t0"][0].tolist()) == [20, 21, 22, 23]
        q.put((rank, "ok"))
    except Exception as e:      # noqa: BLE001
        import traceback
        q.put((rank, f"{type(e).__name__}: {e}\n{traceback.format_exc()}"))
    finally:
        dist.destroy_process_group()


def test_incremental_dump_gathers_all_ranks_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for r, msg in res:
        assert msg == "ok", f"rank {r}: {msg}"
