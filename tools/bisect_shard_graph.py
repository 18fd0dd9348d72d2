"""Development aid: CUDA-graph step of the row-wise sharded wrapper against its eager step (one rank), per prepare_sorts mode and per
development toggle (OPT2/3/4 env) — how the dangling-offsets bug of make_graphed_step was found."""
import os, sys, numpy as np, torch, torch.distributed as dist
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/recsys-examples_b200'); sys.path.insert(0,'/root/repo/tests')
from test_dist_gpu import _mk, _free_port, D
cuda=torch.device('cuda',0)
dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{_free_port()}", rank=0, world_size=1, device_id=cuda)
from dynamicemb import DynamicEmbPoolingMode, EmbOptimType
from dynamicemb.shard import RowWiseShardedDynamicEmbedding
from dynamicemb import _native as N
for opt in (2, 3, 4):
    N.lib.demb_set_option(opt, int(os.environ.get(f"OPT{opt}", "1")))
print("options", [N.lib.demb_get_option(o) for o in (2, 3, 4)], flush=True)
for pm in [int(x) for x in os.environ.get("PMS","0,1,2,3").split(",")]:
    la, lb = (_mk(cuda, 1 << 16, DynamicEmbPoolingMode.NONE, EmbOptimType.EXACT_ADAGRAD, 0.05) for _ in range(2))
    la.train(); lb.train()
    ma = RowWiseShardedDynamicEmbedding(la, None, max_ids_per_step=4096)
    mb = RowWiseShardedDynamicEmbedding(lb, None, max_ids_per_step=4096)
    ma.prepare_sorts = mb.prepare_sorts = pm
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
    torch.cuda.synchronize()
    print('pm',pm,'after warmup: tables equal', torch.equal(la.tables.table_storage_, lb.tables.table_storage_), 'values equal', torch.equal(la._values, lb._values))
    for i,b in enumerate(batches[1:]):
        ids_static.copy_(b)
        graph.replay()
        torch.cuda.synchronize(); print('   graph replay ok', flush=True)
        o = mb(b, lengths)
        torch.cuda.synchronize(); print('   eager fwd ok', flush=True)
        o.backward(grad)
        torch.cuda.synchronize(); print('   eager bwd ok', flush=True)
        print('  replay',i,'out', torch.equal(out,o), 'maxdiff', float((out-o).abs().max()), 'values', torch.equal(la._values, lb._values), 'tables', torch.equal(la.tables.table_storage_, lb.tables.table_storage_))
    # eager vs eager determinism
    lc, ld = (_mk(cuda, 1 << 16, DynamicEmbPoolingMode.NONE, EmbOptimType.EXACT_ADAGRAD, 0.05) for _ in range(2))
    lc.train(); ld.train()
    mc = RowWiseShardedDynamicEmbedding(lc, None, max_ids_per_step=4096); md = RowWiseShardedDynamicEmbedding(ld, None, max_ids_per_step=4096)
    mc.prepare_sorts = md.prepare_sorts = pm
    ok=True
    for b in batches:
        oc=mc(b,lengths); oc.backward(grad); od=md(b,lengths); od.backward(grad)
        ok = ok and torch.equal(oc,od)
    print('  eager-vs-eager equal', ok, torch.equal(lc._values, ld._values))
dist.destroy_process_group()
