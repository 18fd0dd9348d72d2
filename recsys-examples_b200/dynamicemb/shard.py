"""Row-wise model-parallel wrapper around BatchedDynamicEmbeddingTablesV2.

`RowWiseShardedDynamicEmbedding` is the executable equivalent of what TorchRec builds from the reference's sharders
(/root/reference/corelib/dynamicemb/dynamicemb/shard/embedding.py:78 ShardedDynamicEmbeddingCollection.input_dist :277 -> compute ->
output_dist; shard/embeddingbag.py:79; planner/rw_sharding.py:83 RwSequenceDynamicEmbeddingSharding, :189 RwPooledDynamicEmbeddingSharding):
per-rank index dedup -> route every unique id to the rank that owns it -> lookup in the owner's hash table -> rows back -> undo routing +
dedup (sequence) or pool the bags (EmbeddingBagCollection); backward mirrors it with the gradient rows.  Each rank owns an independent
hash table of get_sharded_table_capacity(num_embeddings, world_size) rows.

Where the reference (through TorchRec) runs block_bucketize -> all_to_all(lengths) -> HOST SYNC -> all_to_all(ids) -> lookup ->
all_to_all(rows) with another host sync for the dedup count, this wrapper keeps every count on the device and moves ids, rows and gradient
rows with NVLink peer stores from inside the kernels that produce them (csrc/demb_shard.cu): no NCCL call, no host synchronisation, so the
whole sharded training step is a fixed launch sequence and can be captured in a CUDA graph (`make_graphed_step`).  torch.distributed is used
once, at construction, to exchange the 64-byte IPC handles of the symmetric buffers.

Pooling (SUM / MEAN) happens at the REQUESTER from the unique rows it received — the same accumulation order as the unsharded module,
hence bit-identical pooled outputs (TorchRec pools partial sums per owner and reduce-scatters them; MEAN needs a divisor callback there).

The TorchRec-facing class names of the reference (DynamicEmbeddingCollectionSharder, ...) are kept for drop-in imports; they need torchrec
at construction time (not installed in the build image) and delegate to this wrapper.
"""
import ctypes
from typing import Dict, List, Optional

import torch
import torch.distributed as dist
from torch import nn

from . import _native as N
from . import dynamicemb_extensions as ext
from .batched_dynamicemb_tables import BatchedDynamicEmbeddingTablesV2, PrefetchState
from .types import DynamicEmbPoolingMode


class _RawCuda:
    """__cuda_array_interface__ view of raw device memory (the symmetric buffer is cudaMalloc'ed by the C library, not by torch)."""

    def __init__(self, ptr: int, shape, typestr: str):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 2}


class SymmetricBuffer:
    """One cudaMalloc'ed, zero-filled buffer per rank, IPC-mapped on every rank of the group (all ranks on one NVSwitch box)."""

    def __init__(self, nbytes: int, group, device):
        self.group, self.device, self.nbytes = group, device, int(nbytes)
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        ptr = ctypes.c_void_p()
        handle = (ctypes.c_ubyte * 64)()
        N.check(N.lib.demb_ipc_alloc(self.nbytes, ctypes.byref(ptr), ctypes.cast(handle, ctypes.c_void_p)), "ipc_alloc")
        self.local_ptr = int(ptr.value)
        handles = [None] * self.world
        dist.all_gather_object(handles, bytes(handle), group=group)
        self.ptrs, self._opened = [], []
        for r, h in enumerate(handles):
            if r == self.rank:
                self.ptrs.append(self.local_ptr)
                continue
            p = ctypes.c_void_p()
            hb = (ctypes.c_ubyte * 64).from_buffer_copy(h)
            N.check(N.lib.demb_ipc_open(ctypes.cast(hb, ctypes.c_void_p), ctypes.byref(p)), f"ipc_open(rank {r})")
            self.ptrs.append(int(p.value))
            self._opened.append(int(p.value))
        self.peers = torch.tensor(self.ptrs, dtype=torch.int64, device=device)
        dist.barrier(group=group)                                   # every rank has mapped every buffer before anyone writes

    def view(self, offset: int, shape, dtype=torch.float32) -> torch.Tensor:
        typestr = {torch.float32: "<f4", torch.int64: "<i8", torch.uint8: "|u1"}[dtype]
        return torch.as_tensor(_RawCuda(self.local_ptr + offset, shape, typestr), device=self.device)

    def close(self):
        for p in self._opened:
            N.lib.demb_ipc_close(ctypes.c_void_p(p))
        self._opened = []
        if self.local_ptr:
            N.lib.demb_ipc_free(ctypes.c_void_p(self.local_ptr))
            self.local_ptr = 0


class _ShardedLookup(torch.autograd.Function):
    """forward = the exchange + owner lookup + requester gather; backward = gradient reduce at the requester, peer stores, owner update."""

    @staticmethod
    def forward(ctx, model, ids, offsets, batch_size, dummy):
        out, saved = model._forward_impl(ids, offsets, batch_size)
        ctx.model, ctx.saved = model, saved
        return out

    @staticmethod
    def backward(ctx, grad):
        ctx.model._backward_impl(grad, ctx.saved)
        return None, None, None, None, None


class _ShardCheckpointMixin:
    """dump / load of a row-wise sharded table: every rank writes / reads its own shard in the reference's per-rank file layout
    (`<table>_emb_{keys,values,scores,opt_values}.rank_<r>.world_size_<W>`, dynamicemb/checkpoint.py); a checkpoint written by another
    world size is re-sharded on load by the owner rule of this wrapper's dist_type."""

    def _sync_dist_type(self, dist_type: str) -> None:
        for o in self.local._dynamicemb_options:
            o.dist_type = dist_type

    def dump(self, save_dir: str, optim: bool = False, counter: bool = False, table_names=None) -> None:
        self.local.dump(save_dir, optim=optim, counter=counter, table_names=table_names, pg=self.group)

    def load(self, save_dir: str, optim: bool = False, counter: bool = False, table_names=None) -> None:
        self.local.load(save_dir, optim=optim, counter=counter, table_names=table_names, pg=self.group)


class RowWiseShardedDynamicEmbedding(_ShardCheckpointMixin, nn.Module):
    """forward(ids[n], lengths[F*B]) (KJT of THIS rank's batch, feature-major) ->
         pooling NONE: [n, D] rows in id order (EmbeddingCollection);  SUM / MEAN: [B, F*D] pooled bags (EmbeddingBagCollection).
    `local` is this rank's shard (its pooling_mode selects the output; its tables / value rows / optimizer are used directly).
    Capacities (ids per step): max_ids_per_step = most ids a rank feeds; pair_capacity = most unique ids one rank sends to one owner
    (default: all of them); recv_capacity = most ids an owner accepts in total (default: min(W * pair_capacity, 2 * max_ids_per_step)).
    Exceeding a capacity drops ids (their rows read zeros) and raises at the next `check()` / forward."""

    def __init__(self, local: BatchedDynamicEmbeddingTablesV2, process_group=None, dist_type: str = "hash_roundrobin",
                 num_embeddings_per_feature: Optional[List[int]] = None, use_index_dedup: bool = True, max_ids_per_step: Optional[int] = None,
                 pair_capacity: Optional[int] = None, recv_capacity: Optional[int] = None):
        super().__init__()
        if getattr(local, "_admit_strategy", None) is not None:
            raise NotImplementedError("admission runs the op-by-op prefetch with host-side decisions; the peer-memory step keeps every count "
                                      "on the device — use RowWiseShardedDynamicEmbeddingA2A for a shard with an admission strategy")
        if getattr(local, "_mixed_D", False):
            raise NotImplementedError("mixed embedding dims inside one sharded module are not built: shard the tables of each dim separately")
        if getattr(local, "_caching", False) or getattr(local, "_hybrid", False):
            raise NotImplementedError("the cache / hybrid tiers compact their misses on the host; the peer-memory step keeps every count on the device — "
                                      "use RowWiseShardedDynamicEmbeddingA2A for such a shard")
        self.local = local
        self.group = process_group if process_group is not None else dist.group.WORLD
        self.world_size = dist.get_world_size(self.group)
        self.rank = dist.get_rank(self.group)
        self.use_index_dedup = use_index_dedup        # kept for API parity: ids are always deduplicated per table before they travel
        self.pooling_mode = local.pooling_mode
        self._sync_dist_type(dist_type)
        F, T = local.feature_num, len(local._dynamicemb_options)
        dev = local._device
        self._dev = dev
        if num_embeddings_per_feature is None:
            num_embeddings_per_feature = [1 << 62] * F
        per_table = [num_embeddings_per_feature[local.table_offsets_in_feature[t]] for t in range(T)]
        self._dist_type = torch.full((T,), ext.DIST_TYPE[dist_type], dtype=torch.int32, device=dev)
        # continuous: block size = ceil(hash_size / W) (input_dist.py:231-234)
        self._block_sizes = torch.tensor([(n + self.world_size - 1) // self.world_size for n in per_table], dtype=torch.int64, device=dev)
        self._caps = (max_ids_per_step, pair_capacity, recv_capacity)
        self._buf: Optional[SymmetricBuffer] = None
        self._n_cap = 0
        self._err = torch.zeros(1, dtype=torch.int32, device=dev)
        self._err_host = torch.zeros(1, dtype=torch.int32).pin_memory()
        self._epochs = torch.zeros(4, dtype=torch.int64, device=dev)
        self._empty = nn.Parameter(torch.empty(1, device=dev))        # gives autograd a reason to call backward
        self._prep_req = self._prep_own = None                         # side streams of the two backward sorts
        self.prepare_sorts = 3                                         # bit 0: requester's sort, bit 1: owner's sort launched early on a side stream
        # owner side: rows of the ids the prefetch's lookup stage FOUND are final (and pinned) before insert / evict / row init run, so their
        # NVLink copy back to the requesters starts right after that stage, on a side stream, and overlaps the rest of the prefetch
        self.overlap_hits = True
        self._g2p_stream = self._ev_lookup = self._ev_g2p = None

    # ------------------------------------------------------------------ setup
    def _ensure_buffers(self, n: int) -> None:
        if self._buf is not None and n <= self._n_cap:
            return
        want = torch.tensor([max(n, self._caps[0] or 0)], dtype=torch.int64, device=self._dev)
        dist.all_reduce(want, op=dist.ReduceOp.MAX, group=self.group)   # (re)allocation only: every rank sizes its buffer alike
        n_cap = int(want.item())
        W, D, T = self.world_size, self.local.max_D, len(self.local._dynamicemb_options)
        pair_cap = int(self._caps[1] or n_cap)
        recv_cap = int(self._caps[2] or min(W * pair_cap, 2 * n_cap))
        lay = (ctypes.c_int64 * 6)()
        N.check(N.lib.demb_shard_layout(W, pair_cap, n_cap, D, ctypes.cast(lay, ctypes.c_void_p)), "shard_layout")
        if self._buf is not None:
            torch.cuda.synchronize(self._dev)
            dist.barrier(group=self.group)
            self._buf.close()
        self._buf = SymmetricBuffer(int(lay[5]), self.group, self._dev)
        self._n_cap, self._pair_cap, self._recv_cap = n_cap, pair_cap, recv_cap
        self._lay = [int(x) for x in lay]
        self._rows_back = self._buf.view(self._lay[3], (n_cap, D))
        self._grads_in = self._buf.view(self._lay[4], (W * pair_cap, D))
        dev = self._dev
        self._route_state = torch.zeros(N.lib.demb_shard_route_workspace_bytes(n_cap, W, T), dtype=torch.uint8, device=dev)
        self._recv_ws = torch.zeros(N.lib.demb_shard_recv_workspace_bytes(W, T), dtype=torch.uint8, device=dev)
        self._uscratch_req = ext.unique_scratch(n_cap, T, dev)
        self._uscratch_own = ext.unique_scratch(recv_cap, T, dev)
        self._epochs.zero_()
        if self._g2p_stream is None:
            self._g2p_stream = torch.cuda.Stream(device=dev)
            self._ev_lookup, self._ev_g2p = torch.cuda.Event(), torch.cuda.Event()
            self._ev_lookup.record(); self._ev_g2p.record()             # materialise the cudaEvent handles (outside any graph capture)

    def _barrier(self) -> None:
        N.check(N.launch("peer_barrier", 1, N.lib.demb_peer_barrier, self.world_size, self.rank, self._pair_cap, self._n_cap, self.local.max_D,
                         N.ptr(self._buf.peers), N.ptr(self._err), 0, N.ptr(self._epochs), N.stream()), "peer_barrier")

    def check(self) -> None:
        """Host check of the device error flag (synchronises)."""
        e = int(self._err.item())
        if e:
            self._err.zero_()
            raise RuntimeError("row-wise sharded lookup: " + ("a peer did not reach the barrier (timeout)" if e >= 2 else
                               "an exchange capacity was exceeded and ids were dropped; raise max_ids_per_step / pair_capacity / recv_capacity"))

    # ------------------------------------------------------------------ step
    def _forward_impl(self, ids: torch.Tensor, offsets: torch.Tensor, B: int):
        m = self.local
        W, D, T, F = self.world_size, m.max_D, len(m._dynamicemb_options), m.feature_num
        n = ids.numel()
        dev = self._dev
        # ---- requester: per-table dedup, route the unique ids to their owners (ids + counts are stored into the owners' buffers)
        trange = ext.get_table_range(offsets, m.feature_offsets, F) if T > 1 else None
        num_u, uk, rev, _toffs, _f, utids = ext.segmented_unique_cuda(ids, trange, T, None, want_table_ids=True, scratch=self._uscratch_req)
        train = m.training
        if train and self._prep_req is None:
            self._prep_req, self._prep_own = ext.BackwardPrep(dev), ext.BackwardPrep(dev)
        # the gradient-independent halves of both backward passes (pair list + radix sort) start now, on side streams, under the exchange
        pooled = self.pooling_mode != DynamicEmbPoolingMode.NONE
        _pm = self.prepare_sorts
        prep_req = ext.backward_prepare(self._prep_req, D, rev, max(n, 1)) if (train and not pooled and (_pm & 1)) else None
        send_pos = torch.empty(n, dtype=torch.int64, device=dev)
        ug_addr = torch.empty(n, dtype=torch.int64, device=dev)
        N.check(N.launch("shard_route", 3, N.lib.demb_shard_route, W, self.rank, T, D, self._pair_cap, self._n_cap, N.ptr(self._buf.peers), N.ptr(self._err), n,
                         N.ptr(num_u), N.ptr(uk), N.ptr(utids), N.ptr(self._dist_type), N.ptr(self._block_sizes), N.ptr(send_pos), N.ptr(ug_addr),
                         N.ptr(self._route_state), self._route_state.numel(), N.stream()), "shard_route")
        self._barrier()
        # ---- owner: received ids -> one table-major list; fused prefetch (dedup across sources, probe, insert/evict, init, pin)
        R = self._recv_cap
        ids_recv = torch.empty(R, dtype=ids.dtype, device=dev)
        trange_r = torch.empty(T + 1, dtype=torch.int64, device=dev)
        n_recv = torch.empty(1, dtype=torch.int64, device=dev)
        src_pos = torch.empty(R, dtype=torch.int64, device=dev)
        dst_addr = torch.empty(R, dtype=torch.int64, device=dev)
        N.check(N.launch("shard_recv", 2, N.lib.demb_shard_recv, W, self.rank, T, D, self._pair_cap, self._n_cap, R, N.ptr(self._buf.peers), N.ptr(self._err),
                         N.ptr(ids_recv), N.ptr(trange_r), N.ptr(n_recv), N.ptr(src_pos), N.ptr(dst_addr), N.ptr(self._recv_ws), self._recv_ws.numel(),
                         N.stream()), "shard_recv")
        overlap = self.overlap_hits
        if overlap:
            hit = torch.empty(R, dtype=torch.int8, device=dev)
            N.check(N.lib.demb_train_prefetch_hook(ctypes.c_void_p(self._ev_lookup.cuda_event), N.ptr(hit)), "train_prefetch_hook")
        st = m._prefetch_device_count(ids_recv, trange_r if T > 1 else None, T, n_recv, self._uscratch_own)
        prep_own = ext.backward_prepare(self._prep_own, D, st.reverse_indices, R, n_dev=n_recv, grad_row_of=src_pos) if (train and (_pm & 2)) else None
        if overlap:
            main, side = torch.cuda.current_stream(), self._g2p_stream
            side.wait_event(self._ev_lookup)                            # recorded by the prefetch right after its lookup stage
            with torch.cuda.stream(side):
                N.check(N.launch("gather_to_peers.hits", 1, N.lib.demb_shard_gather_to_peers_part, N.ptr(m._values), m._values.stride(0), D, R, N.ptr(n_recv),
                                 N.ptr(st.rows), N.ptr(st.reverse_indices), N.ptr(dst_addr), N.ptr(hit), 1, N.stream()), "gather_to_peers.hits")
                self._ev_g2p.record(side)
            N.check(N.launch("gather_to_peers", 1, N.lib.demb_shard_gather_to_peers_part, N.ptr(m._values), m._values.stride(0), D, R, N.ptr(n_recv),
                             N.ptr(st.rows), N.ptr(st.reverse_indices), N.ptr(dst_addr), N.ptr(hit), 2, N.stream()), "gather_to_peers")
            main.wait_event(self._ev_g2p)
        else:
            N.check(N.launch("gather_to_peers", 1, N.lib.demb_shard_gather_to_peers, N.ptr(m._values), m._values.stride(0), D, R, N.ptr(n_recv), N.ptr(st.rows),
                             N.ptr(st.reverse_indices), N.ptr(dst_addr), N.stream()), "gather_to_peers")
        self._barrier()
        # ---- requester: one gather from rows_back undoes routing + dedup; pooled modes pool here
        out = ext.gather_forward(self._rows_back, D, send_pos, rev, n, offsets=offsets if pooled else None, batch_size=B if pooled else 0,
                                 num_features=F if pooled else 0, combiner=int(self.pooling_mode) if pooled else -1, out_dtype=m.output_dtype)
        return out, (rev, ug_addr, offsets, B, st, n_recv, src_pos, n, prep_req, prep_own)

    def _backward_impl(self, grad: torch.Tensor, saved) -> None:
        rev, ug_addr, offsets, B, st, n_recv, src_pos, n, prep_req, prep_own = saved
        m = self.local
        D, F = m.max_D, m.feature_num
        pooled = self.pooling_mode != DynamicEmbPoolingMode.NONE
        grad = grad.contiguous().to(torch.float32)
        opt = m._optimizer
        if opt.args.gradient_clipping:
            grad = grad.clamp(-opt.args.max_gradient, opt.args.max_gradient)
        # requester: reduce the gradient rows per unique id; every reduced row is stored straight into its owner's grads_in segment
        ext.backward(None, D, rev, max(n, 1), None, grad, offsets=offsets if pooled else None, batch_size=B if pooled else 0,
                     num_features=F if pooled else 0, combiner=int(self.pooling_mode) if pooled else -1, unique_grad_addr=ug_addr, prepared=prep_req)
        self._barrier()
        # owner: fused reduce across sources + optimizer row update on the received gradient rows
        opt.step()
        ext.backward(m._values, D, st.reverse_indices, self._recv_cap, st.rows, self._grads_in, n_dev=n_recv, grad_row_of=src_pos, prepared=prep_own,
                     **opt.kernel_kwargs())
        m._unpin(st)

    def forward(self, ids: torch.Tensor, lengths: torch.Tensor) -> torch.Tensor:
        m = self.local
        F = m.feature_num
        if ids.dtype != m.index_type:
            ids = ids.to(m.index_type)
        ids = ids.contiguous()
        self._ensure_buffers(ids.numel())
        assert lengths.numel() % F == 0, "lengths must be [F * B] (feature-major)"
        B = lengths.numel() // F
        offsets = torch.zeros(lengths.numel() + 1, dtype=torch.int64, device=ids.device)
        torch.cumsum(lengths, 0, out=offsets[1:])
        if not torch.cuda.is_current_stream_capturing():
            if int(self._err_host[0]) != 0:                         # value copied at the previous step (no synchronisation here)
                self.check()
            self._err_host.copy_(self._err, non_blocking=True)
        if not (m.training and torch.is_grad_enabled()):
            out, saved = self._forward_impl(ids, offsets, B)
            m._unpin(saved[4])
            return out
        return _ShardedLookup.apply(self, ids, offsets, B, self._empty)

    # ------------------------------------------------------------------ CUDA-graph training step
    def make_graphed_step(self, ids_static: torch.Tensor, lengths: torch.Tensor, grad_static: torch.Tensor, with_loss: bool = True):
        """The whole sharded training step (exchange, owner prefetch, rows back, gather, gradient exchange, owner update) as ONE CUDA graph
        per rank — possible because nothing in it reads a count back to the host.  Every rank must call this, and replay, in lockstep
        (the in-graph barriers wait for all ranks).  Returns (graph-like object with replay(), out, loss).  Same restrictions and side
        effects as BatchedDynamicEmbeddingTablesV2.make_graphed_step (three real warm-up steps before capture)."""
        m = self.local
        assert m.training and m._fused_prefetch
        self._ensure_buffers(ids_static.numel())
        F = m.feature_num
        B = lengths.numel() // F
        offsets = torch.zeros(lengths.numel() + 1, dtype=torch.int64, device=self._dev)
        torch.cumsum(lengths, 0, out=offsets[1:])

        def step():
            out_, saved = self._forward_impl(ids_static, offsets, B)
            loss_ = out_.sum() if with_loss else None
            self._backward_impl(grad_static, saved)
            return out_, loss_

        cur = torch.cuda.current_stream(self._dev)
        side = torch.cuda.Stream(self._dev)
        side.wait_stream(cur)
        with torch.no_grad(), torch.cuda.stream(side):
            for _ in range(3):
                step()
        cur.wait_stream(side)
        torch.cuda.synchronize(self._dev)
        dist.barrier(group=self.group)
        graph = torch.cuda.CUDAGraph()
        host_scores, host_iter = dict(m._scores), m._optimizer.iter
        with torch.no_grad(), torch.cuda.graph(graph):
            out, loss = step()
        m._scores, m._optimizer.iter = host_scores, host_iter
        from .batched_dynamicemb_tables import _GraphedStep
        return _GraphedStep(m, graph, keepalive=(ids_static, lengths, offsets, grad_static, self)), out, loss


class _PoolRows(torch.autograd.Function):
    """[n, D] rows of the ids of this rank's batch (feature-major) -> [B, F*D] pooled bags, by the pooled gather kernel over the rows as a
    value table; backward hands every id its bag's gradient (divided by the bag length for MEAN) through the reduce kernel."""

    @staticmethod
    def forward(ctx, rows, offsets, batch_size, num_features, combiner):
        n, D = rows.shape
        ctx.save_for_backward(offsets)
        ctx.shape, ctx.batch_size, ctx.num_features, ctx.combiner = (n, D), batch_size, num_features, combiner
        ident = torch.arange(n, dtype=torch.int64, device=rows.device)
        vals = rows.to(torch.float32).contiguous()                 # the gather kernel reads fp32 value rows (the shard may emit bf16 / fp16)
        ctx.row_dtype = rows.dtype
        return ext.gather_forward(vals, D, ident, None, n, offsets=offsets, batch_size=batch_size, num_features=num_features, combiner=combiner,
                                  out_dtype=rows.dtype, out=torch.empty(batch_size, num_features * D, dtype=rows.dtype, device=rows.device))

    @staticmethod
    def backward(ctx, grad):
        (offsets,) = ctx.saved_tensors
        n, D = ctx.shape
        ident = torch.arange(n, dtype=torch.int64, device=grad.device)
        g = ext.reduce_grads(ident, grad.contiguous().to(torch.float32), n, ctx.batch_size, D, offsets=offsets, combiner=ctx.combiner,
                             total_D=ctx.num_features * D)
        return g.to(ctx.row_dtype), None, None, None, None


class RowWiseShardedDynamicEmbeddingA2A(_ShardCheckpointMixin, nn.Module):
    """The TorchRec-shaped variant of the same data flow over torch.distributed collectives (block bucketize -> all_to_all(lengths, ids) ->
    lookup -> all_to_all(rows), dynamicemb/input_dist.py): for process groups whose ranks do not share an NVLink/NVSwitch domain (multi
    node), for shards that decide on the host (admission, cache / hybrid tiers) and for the CPU (gloo) tests of the host logic.  Two host
    synchronisations per step, as in TorchRec.  `local` is a sequence-mode module; `pooling_mode` SUM / MEAN makes the wrapper an
    EmbeddingBagCollection: the rows come back per id and are pooled at the requester (like the peer-memory wrapper)."""

    def __init__(self, local: BatchedDynamicEmbeddingTablesV2, process_group=None, dist_type: str = "hash_roundrobin",
                 num_embeddings_per_feature: Optional[List[int]] = None, use_index_dedup: bool = True,
                 pooling_mode: DynamicEmbPoolingMode = DynamicEmbPoolingMode.NONE):
        super().__init__()
        assert local.pooling_mode == DynamicEmbPoolingMode.NONE, "the shard looks rows up per id; the wrapper pools (pooling_mode=)"
        self.pooling_mode = DynamicEmbPoolingMode(pooling_mode)
        self.local, self.group, self.use_index_dedup = local, process_group, use_index_dedup
        self.world_size = dist.get_world_size(process_group)
        self._sync_dist_type(dist_type)
        F, dev = local.feature_num, local._device
        self._dist_type = torch.full((F,), ext.DIST_TYPE[dist_type], dtype=torch.int32, device=dev)
        if num_embeddings_per_feature is None:
            num_embeddings_per_feature = [1 << 62] * F
        self._block_sizes = torch.tensor([(n + self.world_size - 1) // self.world_size for n in num_embeddings_per_feature], dtype=torch.int64, device=dev)

    def _bucketize(self, lengths, ids):
        B = lengths.numel() // self.local.feature_num
        nl, ni, _, perm = ext.block_bucketize_sparse_features(lengths, ids, B, self.world_size, self._block_sizes, self._dist_type, sequence=True)
        return nl, ni, perm

    def _unique(self, ids, trange, F):
        num_u, uk, reverse, toffs, _ = ext.segmented_unique_cuda(ids, trange, F, None)
        return int(num_u.item()), uk, reverse, toffs                 # host sync: the exchange sizes depend on it

    def forward(self, ids: torch.Tensor, lengths: torch.Tensor) -> torch.Tensor:
        from .input_dist import rw_sharded_lookup
        rows = rw_sharded_lookup(ids, lengths, self.local.feature_num, self.group, local_fn=self.local, bucketize_fn=self._bucketize,
                                 unique_fn=self._unique if self.use_index_dedup else None,
                                 reduce_fn=lambda idx, grad, num_rows: ext.reduce_grads(idx, grad, num_rows, 0, grad.shape[1]))
        if self.pooling_mode == DynamicEmbPoolingMode.NONE:
            return rows
        F = self.local.feature_num
        offsets = torch.zeros(lengths.numel() + 1, dtype=torch.int64, device=ids.device)
        torch.cumsum(lengths.to(torch.int64), 0, out=offsets[1:])
        return _PoolRows.apply(rows, offsets, lengths.numel() // F, F, int(self.pooling_mode))


def _need_torchrec():
    try:
        import torchrec  # noqa: F401
    except ImportError as e:
        raise ImportError("the TorchRec-facing sharder classes need `torchrec` (release/V1.5.0 in the reference image, docker/Dockerfile:33); "
                          "use RowWiseShardedDynamicEmbedding directly when torchrec is not installed") from e


class DynamicEmbParameterConstraints:
    """planner/planner.py: TorchRec ParameterConstraints + use_dynamicemb + dynamicemb_options."""

    def __init__(self, use_dynamicemb: bool = False, dynamicemb_options=None, **kwargs):
        self.use_dynamicemb, self.dynamicemb_options, self.kwargs = use_dynamicemb, dynamicemb_options, kwargs


def plan_row_wise(constraints: Dict[str, DynamicEmbParameterConstraints], num_embeddings: Dict[str, int], world_size: int) -> Dict[str, dict]:
    """The part of DynamicEmbeddingShardingPlanner that matters for DynamicEmb tables (planner/planner.py:213 + dynamicemb_config.py:696-765):
    every dynamicemb table is sharded ROW_WISE over all ranks with the CUSTOMIZED compute kernel, and each rank's table gets
    max_capacity = get_sharded_table_capacity(num_embeddings, W, bucket_capacity).  Returns {table: {sharding_type, compute_kernel, ranks,
    local_capacity, dist_type}} and writes the per-rank capacity into the table's options (as the reference's planner does)."""
    from .types import get_sharded_table_capacity
    plan = {}
    for name, c in constraints.items():
        if not c.use_dynamicemb:
            continue
        o = c.dynamicemb_options
        cap = get_sharded_table_capacity(num_embeddings[name], world_size, o.bucket_capacity)
        o.max_capacity = cap
        if o.init_capacity is None or o.init_capacity > cap:
            o.init_capacity = cap
        plan[name] = {"sharding_type": "row_wise", "compute_kernel": "customized_kernel", "ranks": list(range(world_size)), "local_capacity": cap,
                      "dist_type": o.dist_type}
    return plan


class _SharderBase:
    """Common part of the two sharders: given the unsharded module's tables and a row-wise plan, build this rank's
    BatchedDynamicEmbeddingTablesV2 shard and wrap it.  `shard()` keeps TorchRec's ModuleSharder signature (module, params, env, device);
    `module` only needs `.embedding_configs()` / `.embedding_bag_configs()` (TorchRec EmbeddingCollection / EmbeddingBagCollection do)."""
    pooled = False

    def __init__(self, qcomm_codecs_registry=None, use_index_dedup: bool = False, fused_params=None):
        self.qcomm_codecs_registry, self.use_index_dedup, self.fused_params = qcomm_codecs_registry, use_index_dedup, dict(fused_params or {})

    def sharding_types(self, compute_device_type: str) -> List[str]:
        return ["row_wise"]

    def compute_kernels(self, sharding_type: str, compute_device_type: str) -> List[str]:
        return ["customized_kernel"]

    def shard(self, module, params: Dict[str, dict], env=None, device: Optional[torch.device] = None, module_fqn: Optional[str] = None):
        from .types import DynamicEmbTableOptions, EmbOptimType
        configs = module.embedding_bag_configs() if self.pooled else module.embedding_configs()
        pg = getattr(env, "process_group", None)
        opts, names, fmap, hash_sizes = [], [], [], []
        for t, cfg in enumerate(configs):
            p = params[cfg.name]
            assert p["sharding_type"] == "row_wise" and p["compute_kernel"] == "customized_kernel", "DynamicEmb tables shard row-wise (planner/rw_sharding.py)"
            o: DynamicEmbTableOptions = self.fused_params.get("dynamicemb_options", {}).get(cfg.name) or p.get("dynamicemb_options")
            assert o is not None, f"no DynamicEmbTableOptions for table {cfg.name}"
            o.dim, o.max_capacity = cfg.embedding_dim, p["local_capacity"]
            opts.append(o); names.append(cfg.name)
            for _ in cfg.feature_names:
                fmap.append(t); hash_sizes.append(cfg.num_embeddings)
        pooling = DynamicEmbPoolingMode.NONE
        if self.pooled:
            kinds = {str(getattr(c, "pooling", "SUM")).split(".")[-1].upper() for c in configs}
            assert len(kinds) == 1, "one pooling type per EmbeddingBagCollection shard"
            pooling = DynamicEmbPoolingMode.MEAN if kinds.pop() == "MEAN" else DynamicEmbPoolingMode.SUM
        fp = {k: v for k, v in self.fused_params.items() if k not in ("dynamicemb_options",)}
        optimizer = fp.pop("optimizer", EmbOptimType.SGD)
        # the module's own tier rule (batched_dynamicemb_tables.py: value bytes against local_hbm_for_values), evaluated up front
        from .optimizer import get_optimizer_state_dim
        max_d = max(o.dim for o in opts)
        vdim = (max_d + get_optimizer_state_dim(optimizer if opts[0].training else EmbOptimType.NONE, max_d) + 3) // 4 * 4
        total, local_hbm = sum(o.max_capacity * 4 * vdim for o in opts), sum(o.local_hbm_for_values for o in opts)
        tiered = (total > local_hbm) if opts[0].caching else (0 < local_hbm < total)
        if opts[0].admit_strategy is not None or tiered:
            # admission / cache / hybrid tiers decide on the host side of the op sequence: the all_to_all wrapper carries them; the shard
            # looks rows up per id and the wrapper pools for an EmbeddingBagCollection
            local = BatchedDynamicEmbeddingTablesV2(opts, table_names=names, feature_table_map=fmap, pooling_mode=DynamicEmbPoolingMode.NONE,
                                                    device=device, optimizer=optimizer, **fp)
            return RowWiseShardedDynamicEmbeddingA2A(local, pg, dist_type=opts[0].dist_type, num_embeddings_per_feature=hash_sizes,
                                                     use_index_dedup=self.use_index_dedup, pooling_mode=pooling)
        local = BatchedDynamicEmbeddingTablesV2(opts, table_names=names, feature_table_map=fmap, pooling_mode=pooling, device=device,
                                                optimizer=optimizer, **fp)
        return RowWiseShardedDynamicEmbedding(local, pg, dist_type=opts[0].dist_type, num_embeddings_per_feature=hash_sizes,
                                              use_index_dedup=self.use_index_dedup)


class DynamicEmbeddingCollectionSharder(_SharderBase):
    """shard/embedding.py:343 — (qcomm_codecs_registry, use_index_dedup, fused_params): sequence embeddings."""
    pooled = False


class DynamicEmbeddingBagCollectionSharder(_SharderBase):
    """shard/embeddingbag.py:79 — pooled embeddings."""
    pooled = True

    def __init__(self, qcomm_codecs_registry=None, fused_params=None):
        super().__init__(qcomm_codecs_registry, False, fused_params)


class DynamicEmbeddingShardingPlanner:
    """planner/planner.py:213 — same constructor arguments as the reference (`eb_configs, topology, batch_size, enumerator,
    storage_reservation, proposer, partitioner, performance_model, stats, constraints, debug`); plans the DynamicEmb tables row-wise itself
    (`plan_row_wise`: what the reference's constructor does in `_prepare_dynemb_table_options` + its `_dyn_emb_plan` loop, :271-350); the
    TorchRec-side arguments are kept for tables without `use_dynamicemb`, which are left to TorchRec's EmbeddingShardingPlanner (needs torchrec).
    World size: `world_size=` if given, else `topology.world_size`, else the default process group's size, else 1.
    Also accepted (this package's earlier call form): `DynamicEmbeddingShardingPlanner(constraints_dict, world_size=W).plan({table: rows})`."""

    def __init__(self, eb_configs=None, topology=None, batch_size: Optional[int] = None, enumerator=None, storage_reservation=None, proposer=None,
                 partitioner=None, performance_model=None, stats=None, constraints: Optional[Dict[str, DynamicEmbParameterConstraints]] = None,
                 debug: bool = True, world_size: Optional[int] = None):
        if isinstance(eb_configs, dict) and constraints is None:          # constraints passed first
            constraints, eb_configs = eb_configs, None
        self.constraints = constraints or {}
        self.eb_configs = list(eb_configs or [])
        self.topology, self.batch_size, self.debug = topology, batch_size, debug
        self.torchrec_args = dict(enumerator=enumerator, storage_reservation=storage_reservation, proposer=proposer, partitioner=partitioner,
                                  performance_model=performance_model, stats=stats)
        self.world_size = world_size if world_size is not None else getattr(topology, "world_size", None)

    def _world(self, pg=None) -> int:
        if self.world_size is not None:
            return int(self.world_size)
        if dist.is_initialized():
            return dist.get_world_size(pg) if pg is not None else dist.get_world_size()
        return 1

    def plan(self, num_embeddings: Optional[Dict[str, int]] = None, sharders=None, pg=None) -> Dict[str, dict]:
        """{table: {sharding_type, compute_kernel, ranks, local_capacity, dist_type}} for every `use_dynamicemb` table.  `num_embeddings`
        defaults to the `num_embeddings` of the constructor's `eb_configs` (the reference reads them from there)."""
        if num_embeddings is not None and not isinstance(num_embeddings, dict):       # reference call form: plan(module, sharders)
            num_embeddings = None
        if num_embeddings is None:
            num_embeddings = {c.name: c.num_embeddings for c in self.eb_configs}
        missing = [n for n, c in self.constraints.items() if c.use_dynamicemb and n not in num_embeddings]
        if missing:
            raise ValueError(f"no num_embeddings for DynamicEmb tables {missing}: pass eb_configs to the constructor or a dict to plan()")
        out = plan_row_wise(self.constraints, num_embeddings, self._world(pg))
        for name in out:
            out[name]["dynamicemb_options"] = self.constraints[name].dynamicemb_options
        return out

    def collective_plan(self, module=None, sharders=None, pg=None) -> Dict[str, dict]:
        """planner/planner.py:351 — every rank computes the same plan (it depends only on the constraints and the world size), so there is
        nothing to broadcast."""
        return self.plan(None, sharders, pg)


ShardedDynamicEmbedding = RowWiseShardedDynamicEmbedding
