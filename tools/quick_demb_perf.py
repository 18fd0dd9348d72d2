"""Quick device-timed look at the embedding kernels (development aid; bench.py is the contract)."""
import os, sys, time
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "recsys-examples_b200"))
from dynamicemb import dynamicemb_extensions as ext
from dynamicemb.scored_hashtable import LinearBucketTable, ScoreArg, ScoreSpec
from dynamicemb.dynamicemb_extensions import ScorePolicy

dev = torch.device("cuda", 0)
D = 128
cap = 32 * 1024 * 1024
t = LinearBucketTable([cap], [ScoreSpec("s", ScorePolicy.ASSIGN)], bucket_capacity=128, device=dev)
values = torch.empty(cap, D, device=dev).normal_()
rng = np.random.default_rng(0)
nkeys = 16 * 1024 * 1024
keys = torch.from_numpy(np.unique(rng.integers(0, 1 << 60, size=nkeys + nkeys // 8, dtype=np.int64))[:nkeys]).to(dev)
def timeit(f, it=10):
    f(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(it): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / it
t0 = time.time()
for i in range(0, nkeys, 1 << 22):
    k = keys[i:i + (1 << 22)]
    t.insert(k, torch.zeros(k.numel(), dtype=torch.int64, device=dev), ScoreArg("s", torch.ones(k.numel(), dtype=torch.int64, device=dev)))
torch.cuda.synchronize()
print(f"insert {nkeys} keys: {time.time()-t0:.2f}s load {t.size()/cap:.3f}")
for n in (1 << 16, 1 << 20, 1 << 22):
    # Zipf-ish ids over present keys
    z = torch.from_numpy((rng.zipf(1.05, size=n) % nkeys).astype(np.int64)).to(dev)
    ids = keys[z]
    nu = torch.unique(ids).numel()
    ms = timeit(lambda: ext.lookup_forward(t.table_storage_, t.table_bucket_offsets_, 128, values, D, ids))
    by = n * (8 + 512) + nu * (24 + 512)
    print(f"fused fwd seq n={n} nu={nu}: {ms*1e3:.1f} us  {by/ms/1e6:.0f} GB/s algorithmic  ({n/ms/1e6:.2f} G lookups/s)")
    uni = keys[torch.randint(0, nkeys, (n,), device=dev)]
    ms = timeit(lambda: ext.lookup_forward(t.table_storage_, t.table_bucket_offsets_, 128, values, D, uni))
    print(f"fused fwd seq uniform n={n}: {ms*1e3:.1f} us  {n*(8+512+24+512)/ms/1e6:.0f} GB/s")
    ms = timeit(lambda: ext.table_lookup(t.table_storage_, t.table_bucket_offsets_, 128, uni, None, None, ScorePolicy.CONST))
    print(f"table_lookup uniform n={n}: {ms*1e3:.1f} us")
    ms = timeit(lambda: ext.segmented_unique_cuda(ids, None, 1, None))
    print(f"segmented_unique n={n}: {ms*1e3:.1f} us")
    _, uk, rev, _, _ = ext.segmented_unique_cuda(ids, None, 1, None)
    _, f, sl = ext.table_lookup(t.table_storage_, t.table_bucket_offsets_, 128, uk[:nu], None, None, ScorePolicy.CONST)
    rows = sl.clone()
    g = torch.randn(n, D, device=dev)
    vals2 = torch.zeros(cap, 2 * D, device=dev) if n == 1 << 20 else None
    ms = timeit(lambda: ext.gather_forward(values, D, rows, rev, n))
    print(f"gather_forward n={n}: {ms*1e3:.1f} us")
    if vals2 is not None:
        ms = timeit(lambda: ext.backward(vals2, D, rev, nu, rows, g, opt_type=3, lr=0.01), it=5)
        by = n * 512 + nu * 2 * 1024
        print(f"backward adagrad n={n} nu={nu}: {ms*1e3:.1f} us  {by/ms/1e6:.0f} GB/s algorithmic")
        del vals2
    # pooled hotness 10
    B = n // 10
    off = torch.arange(0, n + 1, 10, dtype=torch.int64, device=dev)[: B + 1]
    ms = timeit(lambda: ext.lookup_forward(t.table_storage_, t.table_bucket_offsets_, 128, values, D, ids[: B * 10], offsets=off, batch_size=B, num_features=1, combiner=0))
    by = B * 10 * 8 + B * 512 + nu * (24 + 512)
    print(f"fused fwd pooled(10) n={B*10}: {ms*1e3:.1f} us {by/ms/1e6:.0f} GB/s algorithmic")
