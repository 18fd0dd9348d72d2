import os, sys, traceback
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "recsys-examples_b200"))
from dynamicemb import (BatchedDynamicEmbeddingTablesV2, DynamicEmbInitializerArgs, DynamicEmbInitializerMode, DynamicEmbPoolingMode, DynamicEmbScoreStrategy, DynamicEmbTableOptions, EmbOptimType)
dev = torch.device("cuda", 0)
opt = DynamicEmbTableOptions(dim=64, max_capacity=4096, local_hbm_for_values=1 << 40, score_strategy=DynamicEmbScoreStrategy.STEP,
                             initializer_args=DynamicEmbInitializerArgs(mode=DynamicEmbInitializerMode.DEBUG))
m = BatchedDynamicEmbeddingTablesV2([opt], table_names=["t"], pooling_mode=DynamicEmbPoolingMode.NONE, optimizer=EmbOptimType.EXACT_ADAGRAD, device=dev)
m.train()
n = 3000
ids = torch.arange(n, dtype=torch.int64, device=dev) * 31
offs = torch.arange(0, n + 1, dtype=torch.int64, device=dev)
grad = torch.randn(n, 64, device=dev)
try:
    g, out, loss = m.make_graphed_step(ids, offs, grad)
    g.replay(); torch.cuda.synchronize()
    print("graph ok", float(loss))
except Exception:
    traceback.print_exc()
