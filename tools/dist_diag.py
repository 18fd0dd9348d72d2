"""2-GPU diagnostic: NCCL transport / all_to_all bandwidth and a stage-by-stage timing of the row-wise sharded step."""
import os, sys, time
import torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "recsys-examples_b200"))
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
x = torch.randn(48 * 1024 * 1024, device=dev); y = torch.empty_like(x)       # 192 MB, half to each rank
for _ in range(3): dist.all_to_all_single(y, x)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(10): dist.all_to_all_single(y, x)
torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 10
if rank == 0: print(f"all_to_all_single 192 MB/rank: {dt * 1e3:.3f} ms -> {x.numel() * 4 / 2 / dt / 1e9:.1f} GB/s to the peer", flush=True)
from dynamicemb import (BatchedDynamicEmbeddingTablesV2, DynamicEmbInitializerArgs, DynamicEmbInitializerMode, DynamicEmbPoolingMode, DynamicEmbScoreStrategy, DynamicEmbTableOptions, EmbOptimType)
from dynamicemb.shard import RowWiseShardedDynamicEmbedding
from dynamicemb import input_dist as idist
import bench
n = 1 << 20
opt = DynamicEmbTableOptions(dim=128, max_capacity=4 * 1024 * 1024, local_hbm_for_values=1 << 50, score_strategy=DynamicEmbScoreStrategy.STEP,
                             initializer_args=DynamicEmbInitializerArgs(mode=DynamicEmbInitializerMode.UNIFORM, lower=-0.01, upper=0.01))
m = BatchedDynamicEmbeddingTablesV2([opt], table_names=["t0"], pooling_mode=DynamicEmbPoolingMode.NONE, optimizer=EmbOptimType.EXACT_ADAGRAD, learning_rate=0.1, eps=1e-8, device=dev)
m.train()
model = RowWiseShardedDynamicEmbedding(m, None, dist_type="hash_roundrobin", use_index_dedup=True)
gen = torch.Generator(device=dev).manual_seed(rank)
lengths = torch.ones(n, dtype=torch.int64, device=dev)
grad = torch.randn(n, 128, device=dev)
marks = []
def mark(name):
    torch.cuda.synchronize(); marks.append((name, time.perf_counter()))
orig_in, orig_out, orig_local = idist.rw_input_dist, idist.rw_output_dist, m.forward
for it in range(6):
    ids = bench.power_law_ids(n, gen, dev)
    marks.clear(); mark("start")
    out = model(ids, lengths); mark("forward (dedup + input dist + lookup + output dist)")
    out.backward(grad); mark("backward")
    if rank == 0 and it >= 3:
        print(" | ".join(f"{nm}: {(t - marks[i][1]) * 1e3:.2f} ms" for i, (nm, t) in enumerate(marks[1:])), flush=True)
# finer: time pieces of forward by hand
import dynamicemb.shard as sh
ids = bench.power_law_ids(n, gen, dev)
def T(f, *a, **k):
    torch.cuda.synchronize(); t = time.perf_counter(); r = f(*a, **k); torch.cuda.synchronize(); return r, (time.perf_counter() - t) * 1e3
F, B = 1, n
from dynamicemb import dynamicemb_extensions as ext
offsets = torch.zeros(n + 1, dtype=torch.int64, device=dev); torch.cumsum(lengths, 0, out=offsets[1:])
(res, t_u) = T(ext.segmented_unique_cuda, ids, offsets[::B].contiguous(), F, None)
nu = int(res[0].item()); uk = res[1][:nu]
per_f = res[3][1:] - res[3][:-1]; base = per_f // B; rem = per_f - base * B
l2 = (base[:, None] + (torch.arange(B, device=dev)[None, :] < rem[:, None]).to(torch.int64)).reshape(-1)
(r2, t_in) = T(idist.rw_input_dist, uk, l2, B, F, None, model._bucketize)
ids_fm, lengths_fm, ctx = r2
(r3, t_b) = T(model._bucketize, l2, uk)
offs = torch.zeros(lengths_fm.numel() + 1, dtype=torch.int64, device=dev); torch.cumsum(lengths_fm, 0, out=offs[1:])
(rows, t_l) = T(m, ids_fm, offs)
(o, t_o) = T(idist.rw_output_dist, rows, ctx, None)
if rank == 0: print(f"unique {t_u:.2f} ms | input_dist {t_in:.2f} ms (bucketize alone {t_b:.2f}) | local lookup {t_l:.2f} | output_dist {t_o:.2f} | n_unique {nu} recv {ids_fm.numel()}", flush=True)
dist.barrier(); dist.destroy_process_group()
