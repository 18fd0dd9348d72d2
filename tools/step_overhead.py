"""Where does a training step's host time go?  (development aid)"""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "recsys-examples_b200"))
import importlib.util
spec = importlib.util.spec_from_file_location("bench", os.path.join(ROOT, "bench.py")); b = importlib.util.module_from_spec(spec); spec.loader.exec_module(b)
from dynamicemb import (BatchedDynamicEmbeddingTablesV2, DynamicEmbInitializerArgs, DynamicEmbInitializerMode, DynamicEmbPoolingMode, DynamicEmbScoreStrategy, DynamicEmbTableOptions, EmbOptimType)
dev = torch.device("cuda", 0)
D, n = 128, 1 << 20
opt = DynamicEmbTableOptions(dim=D, max_capacity=8 << 20, local_hbm_for_values=1 << 50, score_strategy=DynamicEmbScoreStrategy.STEP,
                             initializer_args=DynamicEmbInitializerArgs(mode=DynamicEmbInitializerMode.UNIFORM, lower=-0.01, upper=0.01))
m = BatchedDynamicEmbeddingTablesV2([opt], table_names=["t"], pooling_mode=DynamicEmbPoolingMode.NONE, optimizer=EmbOptimType.EXACT_ADAGRAD, device=dev)
m.train()
gen = torch.Generator(device=dev).manual_seed(1)
offs = torch.arange(0, n + 1, dtype=torch.int64, device=dev)
grad = torch.randn(n, D, device=dev)
batches = [b.power_law_ids(n, gen, dev) for _ in range(30)]
for i in range(5):
    o = m(batches[i], offs); o.backward(grad)
torch.cuda.synchronize()
def t(label, f, reps=10):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for i in range(reps): f(i)
    t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f"{label}: host {1e3*(t1-t0)/reps:.3f} ms/step, incl. drain {1e3*(t2-t0)/reps:.3f} ms/step")
t("prefetch only", lambda i: (m.prefetch(batches[5 + i], offs), m._prefetch_states.clear(), m._table._ref_counter.zero_()))
t("fwd+bwd", lambda i: m(batches[5 + i], offs).backward(grad))
def synced(i):
    o = m(batches[15 + i], offs); o.backward(grad); return o.sum().item()
t("fwd+bwd+sum.item()", synced)
hb = [x.cpu().pin_memory() for x in batches[:10]]
dv = torch.empty(n, dtype=torch.int64, device=dev)
hr = torch.zeros(1).pin_memory()
def e2e(i):
    dv.copy_(hb[i], non_blocking=True); o = m(dv, offs); o.backward(grad); hr.copy_(o.sum().reshape(1))
t("e2e (pinned H2D + step + D2H)", e2e)
def e2e2(i):
    dv.copy_(hb[i], non_blocking=True); o = m(dv, offs); s = o.sum(); o.backward(grad); hr.copy_(s.reshape(1))
t("e2e, sum before backward", e2e2)
