"""Which part of the pipelined e2e loop is slow?  (development aid)"""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "recsys-examples_b200"))
import bench
from dynamicemb import (BatchedDynamicEmbeddingTablesV2, DynamicEmbInitializerArgs, DynamicEmbInitializerMode, DynamicEmbPoolingMode, DynamicEmbScoreStrategy, DynamicEmbTableOptions, EmbOptimType)
dev = torch.device("cuda", 0)
n = 1 << 20
opt = DynamicEmbTableOptions(dim=128, max_capacity=8 * 1024 * 1024, local_hbm_for_values=1 << 50, score_strategy=DynamicEmbScoreStrategy.STEP,
                             initializer_args=DynamicEmbInitializerArgs(mode=DynamicEmbInitializerMode.UNIFORM, lower=-0.01, upper=0.01))
m = BatchedDynamicEmbeddingTablesV2([opt], table_names=["t0"], pooling_mode=DynamicEmbPoolingMode.NONE, optimizer=EmbOptimType.EXACT_ADAGRAD, learning_rate=0.1, eps=1e-8, device=dev)
m.train()
gen = torch.Generator(device=dev).manual_seed(1)
offsets = torch.arange(0, n + 1, dtype=torch.int64, device=dev)
grad = torch.randn(n, 128, device=dev)
K = 20
batches = [bench.power_law_ids(n, gen, dev) for _ in range(K)]
host = [b.cpu().pin_memory() for b in batches]
dev_ids = torch.empty(n, dtype=torch.int64, device=dev)
dev_ids.copy_(batches[0])
g = m.make_graphed_step(dev_ids, offsets, grad)
res = torch.zeros(K, dtype=torch.float32).pin_memory()
cur = torch.cuda.current_stream(dev)
print("current stream:", cur, "default:", torch.cuda.default_stream(dev))

def timed(name, fn):
    torch.cuda.synchronize(); t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); print(f"{name:60s} {(time.perf_counter() - t0) / K * 1e3:.3f} ms/step", flush=True)

def a():
    for i in range(K):
        dev_ids.copy_(host[i], non_blocking=True); g[0].replay(); res[i:i+1].copy_(g[2].reshape(1), non_blocking=False)
timed("A serial: H2D, replay, blocking D2H", a)
def b():
    for i in range(K):
        dev_ids.copy_(host[i], non_blocking=True); g[0].replay(); res[i:i+1].copy_(g[2].reshape(1), non_blocking=True)
timed("B serial, async D2H, one sync at the end", b)
cs = torch.cuda.Stream(dev)
stg = [torch.empty_like(dev_ids) for _ in range(2)]
def c():
    h2d = [torch.cuda.Event() for _ in range(2)]; cons = [torch.cuda.Event() for _ in range(2)]
    for e in cons: e.record(cur)
    for i in range(K):
        s = i % 2
        with torch.cuda.stream(cs):
            cs.wait_event(cons[s]); stg[s].copy_(host[i], non_blocking=True); h2d[s].record(cs)
        cur.wait_event(h2d[s]); dev_ids.copy_(stg[s], non_blocking=True); cons[s].record(cur)
        g[0].replay(); res[i:i+1].copy_(g[2].reshape(1), non_blocking=True)
timed("C copy stream + staging, no per-step host sync", c)
def d():
    h2d = [torch.cuda.Event() for _ in range(2)]; cons = [torch.cuda.Event() for _ in range(2)]; done = [torch.cuda.Event() for _ in range(K)]
    for e in cons: e.record(cur)
    for i in range(K):
        s = i % 2
        with torch.cuda.stream(cs):
            cs.wait_event(cons[s]); stg[s].copy_(host[i], non_blocking=True); h2d[s].record(cs)
        cur.wait_event(h2d[s]); dev_ids.copy_(stg[s], non_blocking=True); cons[s].record(cur)
        g[0].replay(); res[i:i+1].copy_(g[2].reshape(1), non_blocking=True); done[i].record(cur)
        if i > 0: done[i - 1].synchronize()
    done[-1].synchronize()
timed("D = C + host consumes result of step i-1", d)
def e():
    for i in range(K):
        g[0].replay()
timed("E replay only", e)
def f():
    for i in range(K):
        stg[0].copy_(host[i], non_blocking=True)
timed("F H2D only (8 MB pinned)", f)
