"""ncu target: HSTU fwd + bwd once inside cudaProfilerStart/Stop (ncu --profile-from-start off)."""
import math, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "recsys-examples_b200"))
from hstu import hstu_ops_gpu as ops
B, S, H, D = int(sys.argv[1]) if len(sys.argv) > 1 else 8, 4096, 8, 128
dev = torch.device("cuda", 0)
T = B * S
buf = torch.randn(T, 4 * H * D, device=dev, dtype=torch.bfloat16)
_, v, q, k = (t.view(T, H, D) for t in buf.split(H * D, dim=-1))
do = torch.randn(T, H, D, device=dev, dtype=torch.bfloat16)
cu = torch.arange(0, T + 1, S, dtype=torch.int32, device=dev)
a = 1 / math.sqrt(D)
for _ in range(2):
    ops.hstu_varlen_fwd_100(q, k, v, cu, cu, S, S, None, None, 1, -1, 0, a)
    ops.hstu_varlen_bwd_100(do, q, k, v, cu, cu, S, S, None, None, None, None, None, 1, -1, 0, a)
torch.cuda.synchronize()
torch.cuda.profiler.start()
ops.hstu_varlen_fwd_100(q, k, v, cu, cu, S, S, None, None, 1, -1, 0, a)
ops.hstu_varlen_bwd_100(do, q, k, v, cu, cu, S, S, None, None, None, None, None, 1, -1, 0, a)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
