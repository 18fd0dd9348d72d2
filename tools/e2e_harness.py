"""SURVEY §8(d) cfg 4 — HSTU-large + DynamicEmb end-to-end training step, both arms in ONE harness.

The reference's own e2e benchmark (examples/hstu/training/benchmark, generate_gin_config.py --kernel_backend cutlass ...) needs torchrec +
megatron + gin, none of which is installed, so — as SURVEY §8(d) prescribes — the step is restated as a minimal loop with the SAME two
hot paths and plain torch bf16 ops of the same shapes for everything else:

    ids (one per token, power-law 1.05, key space 1e9)
      -> row-wise sharded DynamicEmb lookup, D=128 fp32 (dedup, route, owner lookup/insert, rows back)      [arm-specific]
      -> bf16 [T,128] @ W_in[128,1024]
      -> 8 x HSTU layer: layer_norm -> @W_uvqk[1024,4096] -> silu -> split u,v,q,k -> jagged HSTU attention (H=8, D=128, causal)
                         -> layer_norm(attn) * u -> @W_o[1024,1024] + residual                            [attention arm-specific]
      -> mean -> backward through all of it -> fused sparse Adagrad update of the embedding rows            [arm-specific]

  arm "ours"      : this repo's kernels (hstu.hstu_attn_varlen_func, BatchedDynamicEmbeddingTablesV2 / RowWiseShardedDynamicEmbedding)
  arm "ours_fused": arm "ours" with every HSTU layer run through hstu.fused_hstu_op (own layer-norm / SiLU / norm-mul-dropout kernels)
  arm "reference_fused": arm "reference" with every layer run through the reference's fused layer restated on its own Triton glue kernels
  arm "reference" : the reference's own GPU kernels, unmodified: hstu_blackwell CuTe-DSL fwd/bwd and the compiled dynamicemb_extensions
                    ops in the reference's HBM-direct op order (batched_dynamicemb_function.py:559-830, :1044-1300), exchanged with
                    torch.distributed all_to_all like TorchRec does (our input_dist host logic with the reference kernels injected).
                    Both are staged under baseline/_ref (never /root/reference).
B = 32 sequences x S = 4096 tokens per GPU (uniform lengths), samples/s = B * world / step time (max over ranks, CUDA events).
"""
import importlib.util
import math
import os
import sys
import time

import torch
import torch.distributed as dist
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "recsys-examples_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

HID, LAYERS, HEADS, DH, DEMB = 1024, 8, 8, 128, 128
KEY_SPACE, ALPHA = 1_000_000_000, 1.05


def power_law_ids(n, gen, device):
    x = torch.rand(n, device=device, dtype=torch.float64, generator=gen)
    g = 1.0 - ALPHA
    return torch.clamp(torch.pow(x * (KEY_SPACE ** g - 1.0) + 1.0, 1.0 / g), max=KEY_SPACE - 1).to(torch.int64)


class DenseStack(torch.nn.Module):
    """The non-hot-path part of HSTU-large as plain torch bf16 ops (identical in both arms)."""

    def __init__(self, dev, attn_fn):
        super().__init__()
        g = torch.Generator(device=dev).manual_seed(42)
        mk = lambda *s: torch.nn.Parameter((torch.randn(*s, device=dev, generator=g) * 0.02).to(torch.bfloat16))
        self.w_in = mk(DEMB, HID)
        self.w_uvqk = torch.nn.ParameterList([mk(HID, 4 * HID) for _ in range(LAYERS)])
        self.w_o = torch.nn.ParameterList([mk(HID, HID) for _ in range(LAYERS)])
        self.attn_fn = attn_fn

    def forward(self, emb, cu, S):
        x = emb.to(torch.bfloat16) @ self.w_in
        T = x.shape[0]
        for l in range(LAYERS):
            n = F.layer_norm(x, (HID,))
            uvqk = F.silu(n @ self.w_uvqk[l])
            u, v, q, k = uvqk.split(HID, dim=-1)
            a = self.attn_fn(q.view(T, HEADS, DH), k.view(T, HEADS, DH), v.view(T, HEADS, DH), cu, S)
            y = F.layer_norm(a.reshape(T, HID), (HID,)) * u
            x = y @ self.w_o[l] + x
        return x.float().mean()


class FusedStack(torch.nn.Module):
    """The same stack through this repo's FusedHSTULayerFunction (hstu/fused_hstu_op.py: layer-norm / SiLU / norm-mul-dropout glue kernels of
    csrc/hstu_glue.cu + the tcgen05 attention + cuBLAS GEMMs) — what the reference runs as its own fused layer (Triton glue + its attention,
    examples/hstu/ops/fused_hstu_op.py).  Same weights / shapes as DenseStack; the layer norms carry affine parameters (ones / zeros) and a
    zero uvqk bias because the fused op's signature requires them, which is slightly MORE work than DenseStack's plain F.layer_norm."""

    def __init__(self, dev, layer_fn=None):
        super().__init__()
        self.layer_fn = layer_fn            # None = this repo's fused_hstu_op; else f(x, cu, S, w_uvqk, b_uvqk, w_proj, in_w, in_b, out_w, out_b)
        g = torch.Generator(device=dev).manual_seed(42)
        mk = lambda *s: torch.nn.Parameter((torch.randn(*s, device=dev, generator=g) * 0.02).to(torch.bfloat16))
        self.w_in = mk(DEMB, HID)
        self.w_uvqk = torch.nn.ParameterList([mk(HID, 4 * HID) for _ in range(LAYERS)])
        self.w_o = torch.nn.ParameterList([mk(HID, HID) for _ in range(LAYERS)])
        ones = lambda: torch.nn.Parameter(torch.ones(HID, device=dev, dtype=torch.bfloat16))
        zeros = lambda n=HID: torch.nn.Parameter(torch.zeros(n, device=dev, dtype=torch.bfloat16))
        self.in_w = torch.nn.ParameterList([ones() for _ in range(LAYERS)])
        self.in_b = torch.nn.ParameterList([zeros() for _ in range(LAYERS)])
        self.out_w = torch.nn.ParameterList([ones() for _ in range(LAYERS)])
        self.out_b = torch.nn.ParameterList([zeros() for _ in range(LAYERS)])
        self.b_uvqk = torch.nn.ParameterList([zeros(4 * HID) for _ in range(LAYERS)])

    def forward(self, emb, cu, S):
        from hstu.fused_hstu_op import fused_hstu_op
        x = emb.to(torch.bfloat16) @ self.w_in
        alpha = 1.0 / math.sqrt(DH)
        for l in range(LAYERS):
            if self.layer_fn is not None:
                x = self.layer_fn(x, cu, S, self.w_uvqk[l], self.b_uvqk[l], self.w_o[l], self.in_w[l], self.in_b[l], self.out_w[l], self.out_b[l])
            else:
                x = fused_hstu_op(x, cu, S, S, self.w_uvqk[l], self.b_uvqk[l], self.w_o[l], HEADS, DH, DH, 1e-5, 0.0, True, self.in_w[l], self.in_b[l],
                                  self.out_w[l], self.out_b[l], None, None, None, 1, alpha, True, None, True)
        return x.float().mean()


def reference_fused_layer():
    """The reference's FusedHSTULayerFunction (examples/hstu/ops/fused_hstu_op.py:75-1103) restated op by op on the reference's OWN kernels,
    staged under baseline/_ref (the function itself cannot be imported: it pulls in the compiled `hstu` package, nvtx and the training configs):
    Triton layer norm / LN*u*dropout kernels (ops/triton_ops, unmodified), torch addmm + silu (what `_get_addmm_silu_fwd_impl` selects on
    sm_100, :43-48), aten silu_backward + cuBLAS (triton_addmm_silu_bwd, triton_addmm.py:278-309), hstu_blackwell attention fwd / bwd.  Saved
    tensors as with recompute_* off: y (proj input) is saved; du / dq / dk / dv are written into slices of one uvqk gradient buffer when the
    reference attention backward accepts strided outputs (else copied)."""
    sys.path.insert(0, os.path.join(ROOT, "baseline", "_ref", "refglue"))
    sys.path.insert(0, os.path.join(ROOT, "baseline", "_ref"))
    from hstu_blackwell import hstu_ops_gpu as refk
    from ops.triton_ops.triton_layer_norm import triton_weighted_layer_norm_bwd, triton_weighted_layer_norm_fwd
    from ops.triton_ops.triton_norm_mul_dropout import triton_layer_norm_mul_dropout_bwd, triton_layer_norm_mul_dropout_fwd
    alpha = 1.0 / math.sqrt(DH)
    state = {"strided_grads_ok": None}

    class RefFusedLayer(torch.autograd.Function):
        @staticmethod
        def forward(ctx, x, cu, S, w_uvqk, b_uvqk, w_proj, in_w, in_b, out_w, out_b):
            T = x.shape[0]
            normed, mean, rstd, bd, nw = triton_weighted_layer_norm_fwd(x=x, weight=in_w, bias=in_b, eps=1e-5)
            pre = torch.addmm(b_uvqk, normed, w_uvqk)
            act = F.silu(pre)
            u, v, q, k = torch.split(act, [HID, HID, HID, HID], dim=-1)
            r = refk.hstu_varlen_fwd_100(q.view(T, HEADS, DH), k.view(T, HEADS, DH), v.view(T, HEADS, DH), cu, cu, S, S, None, None, 1, -1, 0, alpha, None, None)
            attn = (r[0] if isinstance(r, (tuple, list)) else r).reshape(T, HID)
            y, mean2, rstd2, bd2, nw2, seed = triton_layer_norm_mul_dropout_fwd(x=attn, u=u, weight=out_w, bias=out_b, eps=1e-5, dropout_ratio=0.0,
                                                                               training=True, concat_ux=False, seed=None)
            out = torch.addmm(x, y, w_proj)
            ctx.save_for_backward(x, in_w, in_b, mean, rstd, normed, w_uvqk, pre, act, attn, out_w, out_b, mean2, rstd2, y, w_proj, cu)
            ctx.misc = (S, bd, nw, bd2, nw2, seed)
            return out

        @staticmethod
        def backward(ctx, grad):
            x, in_w, in_b, mean, rstd, normed, w_uvqk, pre, act, attn, out_w, out_b, mean2, rstd2, y, w_proj, cu = ctx.saved_tensors
            S, bd, nw, bd2, nw2, seed = ctx.misc
            T = x.shape[0]
            grad = grad.contiguous()
            u, v, q, k = torch.split(act, [HID, HID, HID, HID], dim=-1)
            dy = torch.mm(grad, w_proj.t())
            d_w_proj = torch.mm(y.t(), grad)
            duvqk = torch.empty_like(pre)
            pre_du, pre_dv, pre_dq, pre_dk = duvqk.split([HID, HID, HID, HID], dim=-1)
            dattn, _, d_out_w, d_out_b, _ = triton_layer_norm_mul_dropout_bwd(dy=dy, x=attn, u=u, weight=out_w, bias=out_b, mean=mean2, rstd=rstd2, BLOCK_D=bd2,
                                                                              num_warps=nw2, eps=1e-5, training=True, dropout_ratio=0.0, seed=seed,
                                                                              concat_ux=False, compute_y=False, du=pre_du)
            qc, kc, vc = (t.view(T, HEADS, DH).contiguous() for t in (q, k, v))      # the reference backward rejects strided q / k / v ("stride_order")
            dout = dattn.view(T, HEADS, DH)
            done = False
            if state["strided_grads_ok"] is not False:
                try:
                    refk.hstu_varlen_bwd_100(dout, qc, kc, vc, cu, cu, S, S, pre_dq.view(T, HEADS, DH), pre_dk.view(T, HEADS, DH), pre_dv.view(T, HEADS, DH),
                                             None, None, 1, -1, 0, alpha, None, False, None, False)
                    state["strided_grads_ok"] = done = True
                except Exception:  # noqa: BLE001
                    state["strided_grads_ok"] = False
            if not done:
                g = refk.hstu_varlen_bwd_100(dout, qc, kc, vc, cu, cu, S, S, None, None, None, None, None, 1, -1, 0, alpha, None, False, None, False)
                pre_dq.copy_(g[0].reshape(T, HID)); pre_dk.copy_(g[1].reshape(T, HID)); pre_dv.copy_(g[2].reshape(T, HID))
            dz = torch.ops.aten.silu_backward(duvqk, pre)
            d_b = torch.sum(dz, dim=0)
            d_normed = torch.mm(dz, w_uvqk.t())
            d_w_uvqk = torch.mm(normed.t(), dz)
            dx, d_in_w, d_in_b = triton_weighted_layer_norm_bwd(dy=d_normed, x=x, weight=in_w, bias=in_b, mean=mean, rstd=rstd, learnable=True, eps=1e-5,
                                                                BLOCK_D=bd, num_warps=nw, dx_accumulate=grad)
            return dx, None, None, d_w_uvqk, d_b, d_w_proj, d_in_w, d_in_b, d_out_w, d_out_b

    return (lambda x, cu, S, *params: RefFusedLayer.apply(x, cu, S, *params)), state


# ---------------------------------------------------------------------------------------------------------------- attention arms
def ours_attention():
    from hstu import hstu_attn_varlen_func
    alpha = 1.0 / math.sqrt(DH)

    def fn(q, k, v, cu, S):
        return hstu_attn_varlen_func(q, k, v, cu, cu, None, None, S, S, S, None, None, 1, (-1, 0), alpha)
    return fn


def reference_attention():
    sys.path.insert(0, os.path.join(ROOT, "baseline", "_ref"))
    from hstu_blackwell import hstu_ops_gpu as refk
    alpha = 1.0 / math.sqrt(DH)

    class RefAttn(torch.autograd.Function):
        @staticmethod
        def forward(ctx, q, k, v, cu, S):
            r = refk.hstu_varlen_fwd_100(q, k, v, cu, cu, S, S, None, None, 1, -1, 0, alpha, None, None)
            out = r[0] if isinstance(r, (tuple, list)) else r
            ctx.save_for_backward(q, k, v, cu)
            ctx.S = S
            return out

        @staticmethod
        def backward(ctx, dout):
            q, k, v, cu = ctx.saved_tensors
            # the reference backward rejects the strided views of the fused uvqk buffer ("stride_order"): its caller has to copy
            g = refk.hstu_varlen_bwd_100(dout.contiguous(), q.contiguous(), k.contiguous(), v.contiguous(), cu, cu, ctx.S, ctx.S, None, None, None, None, None,
                                         1, -1, 0, alpha, None, False, None, False)
            return g[0], g[1], g[2], None, None

    return lambda q, k, v, cu, S: RefAttn.apply(q, k, v, cu, S)


# ---------------------------------------------------------------------------------------------------------------- embedding arms
def ours_embedding(dev, world, capacity, n_ids):
    from dynamicemb import (BatchedDynamicEmbeddingTablesV2, DynamicEmbInitializerArgs, DynamicEmbInitializerMode, DynamicEmbPoolingMode,
                            DynamicEmbScoreStrategy, DynamicEmbTableOptions, EmbOptimType)
    opt = DynamicEmbTableOptions(dim=DEMB, max_capacity=capacity, local_hbm_for_values=1 << 50, score_strategy=DynamicEmbScoreStrategy.STEP,
                                 initializer_args=DynamicEmbInitializerArgs(mode=DynamicEmbInitializerMode.UNIFORM, lower=-0.01, upper=0.01))
    m = BatchedDynamicEmbeddingTablesV2([opt], table_names=["item"], pooling_mode=DynamicEmbPoolingMode.NONE, optimizer=EmbOptimType.EXACT_ADAGRAD,
                                        learning_rate=0.1, eps=1e-8, device=dev)
    m.train()
    if world > 1:
        from dynamicemb.shard import RowWiseShardedDynamicEmbedding
        model = RowWiseShardedDynamicEmbedding(m, None, dist_type="hash_roundrobin", max_ids_per_step=n_ids, recv_capacity=2 * n_ids)
        lengths = torch.ones(n_ids, dtype=torch.int64, device=dev)
        return lambda ids: model(ids, lengths)
    offsets = torch.arange(0, n_ids + 1, dtype=torch.int64, device=dev)
    return lambda ids: m(ids, offsets)


def reference_embedding(dev, world, capacity, n_ids):
    so = os.path.join(ROOT, "baseline", "_ref", "dynamicemb_ext", "dynamicemb_extensions.so")
    spec = importlib.util.spec_from_file_location("dynamicemb_extensions", so)
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    D, C = DEMB, 128
    nb = capacity // C
    # table image + flat value table exactly as scored_hashtable.py:378-425 / key_value_table.py:346-356 lay them out
    storage = torch.empty(17 * C * nb, dtype=torch.uint8, device=dev)
    keys_, dig_, sc_ = ref.table_partition(storage, [torch.int64, torch.uint8, torch.uint64], C, nb)
    keys_.fill_(-1); sc_.fill_(0)

    def fm(k):
        k ^= k >> 33; k = (k * 0xFF51AFD7ED558CCD) & (2 ** 64 - 1); k ^= k >> 33; k = (k * 0xC4CEB9FE1A85EC53) & (2 ** 64 - 1); k ^= k >> 33
        return k
    dig_.fill_((fm(0xFFFFFFFFFFFFFFFF) >> 32) & 0xFF)
    off = torch.tensor([0, nb], dtype=torch.int64, device=dev)
    bsz = torch.zeros(nb, dtype=torch.int32, device=dev)
    ctr = torch.zeros(capacity, dtype=torch.int32, device=dev)
    values = torch.zeros(capacity, 2 * D, device=dev)
    ptrs = torch.tensor([values.data_ptr()], dtype=torch.int64, device=dev)
    vd = torch.tensor([2 * D], dtype=torch.int64, device=dev)
    ed = torch.tensor([D], dtype=torch.int64, device=dev)
    P = ref.ScorePolicy
    state = {"score": 1}

    class RefLookup(torch.autograd.Function):
        """One HBM-direct training lookup in the reference's op order; the optimizer runs in backward (fused), as in the reference."""

        @staticmethod
        def forward(ctx, ids, dummy):
            n = ids.numel()
            segr = torch.tensor([0, n], dtype=torch.int64, device=dev)
            nu_t, uk, rev, _toffs, _ = ref.segmented_unique_cuda(ids, segr, 1, None)
            nu = int(nu_t.item())                                          # reference host sync (batched_dynamicemb_function.py:141)
            uk = uk[:nu]
            tid = torch.zeros(nu, dtype=torch.int64, device=dev)
            score = state["score"]; state["score"] += 1
            sc = torch.full((nu,), score, dtype=torch.int64, device=dev).view(torch.uint64)
            _, found, idx = ref.table_lookup(storage, off, C, uk, tid, sc, P.ASSIGN)
            miss = (~found).nonzero(as_tuple=True)[0]                      # flagged_compact + host sync in the reference
            if miss.numel():
                mk = uk[miss].contiguous()
                mt = torch.zeros(mk.numel(), dtype=torch.int64, device=dev)
                ms = torch.full((mk.numel(),), score, dtype=torch.int64, device=dev).view(torch.uint64)
                new = ref.table_insert(storage, off, C, bsz, mk, mt, ms, P.ASSIGN, ctr)
                init = torch.empty(mk.numel(), 2 * D, device=dev).uniform_(-0.01, 0.01); init[:, D:] = 0
                ref.store_to_flat_table_value(ptrs, new, mt, init, vd, ed, D, True)
                idx[miss] = new
            emb = torch.empty(nu, D, device=dev)
            ref.load_from_flat_table_emb(ptrs, idx, tid, emb, vd, ed, D, True)
            out = torch.empty(n, D, device=dev)
            ref.gather_embedding(emb, out, rev)
            ctx.save_for_backward(rev, idx, tid)
            ctx.nu = nu
            return out

        @staticmethod
        def backward(ctx, grad):
            rev, idx, tid = ctx.saved_tensors
            ug = ref.reduce_grads(rev, grad.contiguous(), ctx.nu, 1, D)
            ref.adagrad_update_for_flat_table(ug, idx, ptrs, tid, vd, ed, 0.1, 1e-8, D, True, 0)
            return None, None

    dummy = torch.zeros(1, device=dev, requires_grad=True)
    if world == 1:
        return lambda ids: RefLookup.apply(ids, dummy)
    # N > 1: the TorchRec data flow (bucketize -> all_to_all(lengths, ids) -> lookup -> all_to_all(rows)) with the reference's kernels
    from dynamicemb.input_dist import rw_sharded_lookup
    blk = torch.tensor([1 << 62], dtype=torch.int64, device=dev)
    dts = torch.tensor([2], dtype=torch.int32, device=dev)                # hash_roundrobin
    lengths = torch.ones(n_ids, dtype=torch.int64, device=dev)

    def bucketize(l, i):
        r = ref.block_bucketize_sparse_features(l, i, False, True, dts, blk, world, None, None, l.numel(), None)
        return r[0], r[1], r[4]

    def unique(i, trange, nf):
        nu_t, uk, rev, toffs, _ = ref.segmented_unique_cuda(i, trange, nf, None)
        return int(nu_t.item()), uk, rev, toffs

    reduce_fn = lambda idx, g, nrows: ref.reduce_grads(idx, g.contiguous(), nrows, 1, D)
    local = lambda ids_fm, offsets_fm: RefLookup.apply(ids_fm.contiguous(), dummy)
    return lambda ids: rw_sharded_lookup(ids, lengths, 1, None, local_fn=local, bucketize_fn=bucketize, unique_fn=unique, reduce_fn=reduce_fn)


# ---------------------------------------------------------------------------------------------------------------- driver
def run(arm: str, dev, world: int, rank: int, steps: int = 6, warmup: int = 3, B: int = 32, S: int = 4096, capacity: int = 16 * 1024 * 1024):
    T = B * S
    ours = arm in ("ours", "ours_fused")
    emb = (ours_embedding if ours else reference_embedding)(dev, world, capacity, T)
    ref_state = None
    if arm == "ours_fused":
        dense = FusedStack(dev)
    elif arm == "reference_fused":
        layer_fn, ref_state = reference_fused_layer()
        dense = FusedStack(dev, layer_fn)
    else:
        dense = DenseStack(dev, ours_attention() if ours else reference_attention())
    cu = torch.arange(0, T + 1, S, dtype=torch.int32, device=dev)
    gen = torch.Generator(device=dev).manual_seed(777 + rank)
    batches = [power_law_ids(T, gen, dev) for _ in range(steps + warmup)]

    def step(ids):
        rows = emb(ids)
        loss = dense(rows, cu, S)
        loss.backward()
        for p in dense.parameters():
            p.grad = None                       # dense optimizer is not part of either hot path; its cost would be identical in both arms
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    t0 = time.time()
    for i in range(warmup):
        step(batches[i])
    barrier()
    jit_s = time.time() - t0
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for i in range(steps):
        step(batches[warmup + i])
    e1.record()
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1) / steps], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item())
    # the embedding part alone (forward + backward with a constant gradient)
    g = torch.randn(T, DEMB, device=dev)
    for i in range(2):
        emb(batches[i]).backward(g)
    barrier()
    e0.record()
    for i in range(steps):
        emb(batches[warmup + i]).backward(g)
    e1.record()
    barrier()
    ems = torch.tensor([e0.elapsed_time(e1) / steps], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ems, op=dist.ReduceOp.MAX)
    extra = {"attention_bwd_wrote_into_uvqk_gradient_slices": ref_state["strided_grads_ok"]} if ref_state is not None else {}
    return {**extra, "arm": arm, "ms_per_step": ms, "samples_per_s": B * world / ms * 1e3, "embedding_fwd_bwd_ms": float(ems.item()), "warmup_incl_jit_s": round(jit_s, 1),
            "tokens_per_gpu": T, "layers": LAYERS, "hidden": HID, "heads": HEADS, "head_dim": DH, "emb_dim": DEMB, "table_rows_per_gpu": capacity}


def run_both(dev, world, rank, **kw):
    out = {"config": "HSTU-large (8 layers, hidden 1024, 8 heads x 128, bf16) + DynamicEmb D=128 fp32 row-wise sharded, B=32 x S=4096 per GPU, "
                     "dense parts = torch bf16 ops of the same shapes in both arms (SURVEY 8(d) cfg 4)"}
    for arm in ("ours", "ours_fused", "reference", "reference_fused"):
        try:
            out[arm] = run(arm, dev, world, rank, **kw)
        except Exception as e:  # noqa: BLE001
            import traceback
            out[arm] = {"error": repr(e)[:300], "trace": traceback.format_exc()[-800:]}
        torch.cuda.empty_cache()
    if "ms_per_step" in out.get("ours", {}) and "ms_per_step" in out.get("reference", {}):
        out["ratio_samples_per_s_ours_over_reference"] = out["reference"]["ms_per_step"] / out["ours"]["ms_per_step"]
        out["ratio_embedding_ours_over_reference"] = out["reference"]["embedding_fwd_bwd_ms"] / out["ours"]["embedding_fwd_bwd_ms"]
    if "ms_per_step" in out.get("ours_fused", {}) and "ms_per_step" in out.get("reference_fused", {}):
        # fused layer against fused layer: our glue kernels + attention vs the reference's Triton glue kernels + its attention (the 1.2x target's ratio)
        out["ratio_samples_per_s_ours_fused_over_reference_fused"] = out["reference_fused"]["ms_per_step"] / out["ours_fused"]["ms_per_step"]
    refs = [out[a]["ms_per_step"] for a in ("reference", "reference_fused") if "ms_per_step" in out.get(a, {})]
    if refs and "ms_per_step" in out.get("ours_fused", {}):
        # the robust headline: our fused layer against the FASTER of the reference's two configurations.  `reference_fused` depends on which
        # configuration Triton's autotuner picks for the reference's layer-norm backward kernels during warm-up: 92-93 ms in four runs,
        # 163-174 ms in two (profiles/r02_e2e_cfg4_n1_fused_layer*.json, r02_bench_line.json)
        out["ratio_samples_per_s_ours_fused_over_best_reference"] = min(refs) / out["ours_fused"]["ms_per_step"]
    if "ms_per_step" in out.get("ours_fused", {}) and "ms_per_step" in out.get("ours", {}):
        out["fused_layer_gain_over_eager_glue_ours"] = out["ours"]["ms_per_step"] / out["ours_fused"]["ms_per_step"]
    if "ms_per_step" in out.get("reference_fused", {}) and "ms_per_step" in out.get("reference", {}):
        out["fused_layer_gain_over_eager_glue_reference"] = out["reference"]["ms_per_step"] / out["reference_fused"]["ms_per_step"]
    return out


if __name__ == "__main__":
    import json
    world = int(os.environ.get("WORLD_SIZE", 1))
    rank = int(os.environ.get("RANK", 0))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    res = run_both(dev, world, rank)
    if rank == 0:
        print(json.dumps(res))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
