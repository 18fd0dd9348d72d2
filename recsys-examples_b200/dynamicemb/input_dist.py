"""Row-wise sharded input/output dist: the only collectives on the embedding path.

Restates, over plain torch.distributed (NCCL on GPUs over NVLink5/NVSwitch, gloo in the CPU tests), what the reference gets
from TorchRec: `RwSparseFeaturesDist` = block_bucketize_sparse_features + KJTAllToAll (lengths, then ids)
(/root/reference/corelib/dynamicemb/dynamicemb/input_dist.py:225-285) and `SequenceEmbeddingsAllToAll` (variable-split
all_to_all of [n, D] fp32 rows, autograd-mirrored) (planner/rw_sharding.py:83).  The device work (bucketize, lookup) is injected
as callables so this host logic is testable on CPU with world_size > 1.

Layout conventions (TorchRec KJT): ids are feature-major, slot = f*B + b; after the exchange a rank holds, for every source rank r,
that rank's ids for every (feature, sample) slot: received slot order = (feature, source rank, sample) once regrouped by feature.
"""
from dataclasses import dataclass
from typing import Callable, List, Optional, Tuple

import torch
import torch.distributed as dist


@dataclass
class DistContext:
    """Everything output_dist needs to route rows back."""
    send_splits: List[int]        # ids sent to each rank
    recv_splits: List[int]        # ids received from each rank
    unbucketize_permute: torch.Tensor   # position of original id i inside the bucketized (rank-major) send buffer
    recv_order: Optional[torch.Tensor]   # permutation applied to received ids to make them feature-major for the local lookup (None = identity)
    num_ids: int


def _all_to_all_single(out: torch.Tensor, inp: torch.Tensor, out_splits, in_splits, group) -> None:
    dist.all_to_all_single(out, inp, output_split_sizes=out_splits, input_split_sizes=in_splits, group=group)


class _RowsAllToAll(torch.autograd.Function):
    """[n_in, D] -> [n_out, D] variable-split all_to_all; backward is the mirrored exchange (TorchRec SequenceEmbeddingsAllToAll)."""

    @staticmethod
    def forward(ctx, rows, out_splits, in_splits, group):
        ctx.out_splits, ctx.in_splits, ctx.group = out_splits, in_splits, group
        out = rows.new_empty(sum(out_splits), rows.shape[1])
        _all_to_all_single(out, rows.contiguous(), out_splits, in_splits, group)
        return out

    @staticmethod
    def backward(ctx, grad):
        g = grad.new_empty(sum(ctx.in_splits), grad.shape[1])
        _all_to_all_single(g, grad.contiguous(), ctx.in_splits, ctx.out_splits, ctx.group)
        return g, None, None, None


class _PermuteRows(torch.autograd.Function):
    """out = x[idx] for a PERMUTATION idx.  backward is the gather by the inverse permutation — never torch's
    index_put_(accumulate=True), whose sort-based kernel took 40 ms per step on a [2^20, 128] gradient."""

    @staticmethod
    def forward(ctx, x, idx, inv_idx):
        ctx.save_for_backward(inv_idx)
        return x.index_select(0, idx)

    @staticmethod
    def backward(ctx, grad):
        (inv_idx,) = ctx.saved_tensors
        return grad.index_select(0, inv_idx), None, None


class _ExpandRows(torch.autograd.Function):
    """out = x[idx] where idx repeats rows (un-dedup + un-bucketize in one gather).  backward reduces the gradient rows per source
    row with the injected `reduce_fn(idx, grad, num_rows)` (the CUDA reduce_grads kernels on GPUs: fixed summation order)."""

    @staticmethod
    def forward(ctx, x, idx, reduce_fn):
        ctx.save_for_backward(idx)
        ctx.num_rows, ctx.reduce_fn = x.shape[0], reduce_fn
        return x.index_select(0, idx)

    @staticmethod
    def backward(ctx, grad):
        (idx,) = ctx.saved_tensors
        return ctx.reduce_fn(idx, grad.contiguous(), ctx.num_rows), None, None


def _inverse_permutation(perm: torch.Tensor) -> torch.Tensor:
    inv = torch.empty_like(perm)
    inv[perm] = torch.arange(perm.numel(), device=perm.device, dtype=perm.dtype)
    return inv


def rw_input_dist(ids: torch.Tensor, lengths: torch.Tensor, batch_size: int, num_features: int, group,
                  bucketize_fn: Callable) -> Tuple[torch.Tensor, torch.Tensor, DistContext]:
    """ids[n] + lengths[F*B] (feature-major) of THIS rank's batch -> (ids_recv feature-major, lengths_recv[F*W*B], ctx).

    bucketize_fn(lengths, ids) -> (new_lengths[W*F*B] rank-major, new_ids[n] rank-major, unbucketize_permute[n])."""
    W = dist.get_world_size(group)
    S = num_features * batch_size
    new_lengths, new_ids, perm = bucketize_fn(lengths, ids)
    # --- exchange lengths: every rank sends its [F*B] length vector for each destination
    recv_lengths = torch.empty_like(new_lengths)
    _all_to_all_single(recv_lengths, new_lengths, [S] * W, [S] * W, group)
    splits = torch.cat([new_lengths.view(W, S).sum(dim=1), recv_lengths.view(W, S).sum(dim=1)]).tolist()   # ONE host sync (TorchRec KJTAllToAll syncs here too)
    send_splits, recv_splits = splits[:W], splits[W:]
    ids_recv = new_ids.new_empty(sum(recv_splits))
    _all_to_all_single(ids_recv, new_ids, recv_splits, send_splits, group)
    # --- received layout is (source rank, feature, sample); the lookup wants feature-major: (feature, source rank, sample)
    if num_features == 1:                 # one feature: the received order already is feature-major, no regrouping pass
        ctx = DistContext(send_splits, recv_splits, perm, None, int(ids.numel()))
        return ids_recv, recv_lengths, ctx
    rl = recv_lengths.view(W, num_features, batch_size)
    lengths_fm = rl.permute(1, 0, 2).reshape(-1)                        # [F * W * B]
    # permutation of ids: compute start offset of each (r, f, b) segment in received order, then gather in (f, r, b) order
    seg_len = rl.reshape(-1)
    seg_start = torch.cumsum(seg_len, 0) - seg_len
    order_segments = torch.arange(W * S, device=ids.device).view(W, num_features, batch_size).permute(1, 0, 2).reshape(-1)
    lens_o = seg_len[order_segments]
    starts_o = seg_start[order_segments]
    total = int(ids_recv.numel())
    if total > 0:
        out_start = torch.cumsum(lens_o, 0) - lens_o
        seg_of = torch.repeat_interleave(torch.arange(W * S, device=ids.device), lens_o, output_size=total)
        recv_order = starts_o[seg_of] + (torch.arange(total, device=ids.device) - out_start[seg_of])
    else:
        recv_order = torch.empty(0, dtype=torch.int64, device=ids.device)
    ids_fm = ids_recv[recv_order]
    ctx = DistContext(send_splits, recv_splits, perm, recv_order, int(ids.numel()))
    return ids_fm, lengths_fm, ctx


def rw_output_dist(rows_fm: torch.Tensor, ctx: DistContext, group, expand: Optional[torch.Tensor] = None,
                   reduce_fn: Optional[Callable] = None) -> torch.Tensor:
    """rows of the locally looked-up ids (feature-major order) -> rows for this rank's original ids, original order.

    expand: reverse indices of the per-rank index dedup (original id i -> position of its unique id), folded into the final gather;
    reduce_fn(idx, grad[n, D], num_rows) -> [num_rows, D] sums the gradient rows that share a source row (needed with `expand`)."""
    # undo the feature-major regrouping, send every row back to the rank that asked, undo the bucketize permutation
    rows_recv_order = rows_fm if ctx.recv_order is None else _PermuteRows.apply(rows_fm, _inverse_permutation(ctx.recv_order), ctx.recv_order)
    back = _RowsAllToAll.apply(rows_recv_order, ctx.send_splits, ctx.recv_splits, group)
    if expand is None:
        return _PermuteRows.apply(back, ctx.unbucketize_permute, _inverse_permutation(ctx.unbucketize_permute))
    assert reduce_fn is not None
    return _ExpandRows.apply(back, ctx.unbucketize_permute[expand], reduce_fn)


def rw_sharded_lookup(ids: torch.Tensor, lengths: torch.Tensor, num_features: int, group, *, local_fn: Callable, bucketize_fn: Callable,
                      unique_fn: Optional[Callable] = None, reduce_fn: Optional[Callable] = None) -> torch.Tensor:
    """The whole row-wise sharded sequence lookup of one rank's batch, device work injected as callables (so the host logic — dedup
    re-spread, exchanges, regrouping, un-bucketize / un-dedup and their backward — runs under gloo on CPU in the tests):

      unique_fn(ids, table_range[F+1], F) -> (num_unique: int, unique_ids[>=num_unique], reverse[n], unique_offsets[F+1])   (None = no dedup)
      bucketize_fn(lengths[F*B], ids)     -> (new_lengths[W*F*B] rank-major, new_ids rank-major, unbucketize_permute)
      local_fn(ids_fm, offsets_fm)        -> rows [n_recv, D] of this rank's table for the received ids (feature-major), differentiable
      reduce_fn(idx, grad[n, D], num_rows)-> [num_rows, D] per-source-row sums (backward of the final expanding gather)

    Mirrors ShardedDynamicEmbeddingCollection.input_dist / compute / output_dist (shard/embedding.py:183-340)."""
    F = num_features
    B = lengths.numel() // F
    reverse = None
    if unique_fn is not None:
        # _dedup_indices (shard/embedding.py:183-275): send each distinct id once per rank.  Unique ids of a feature are
        # re-spread over that feature's B slots (compute_dedup_lengths, unique_op.cu:753) so the KJT stays well-formed.
        offsets = torch.zeros(lengths.numel() + 1, dtype=torch.int64, device=ids.device)
        torch.cumsum(lengths, 0, out=offsets[1:])
        trange = offsets[:: B].contiguous()                      # one "table" per feature for dedup purposes
        nu, uk, reverse, toffs = unique_fn(ids, trange, F)
        ids = uk[:nu]
        per_f = (toffs[1:] - toffs[:-1])                          # unique count per feature
        base = per_f // B
        rem = per_f - base * B
        lengths = (base[:, None] + (torch.arange(B, device=ids.device)[None, :] < rem[:, None]).to(torch.int64)).reshape(-1)
    ids_fm, lengths_fm, ctx = rw_input_dist(ids, lengths.to(torch.int64), B, F, group, bucketize_fn)
    offsets_fm = torch.zeros(lengths_fm.numel() + 1, dtype=torch.int64, device=ids.device)
    torch.cumsum(lengths_fm, 0, out=offsets_fm[1:])
    rows = local_fn(ids_fm, offsets_fm)
    return rw_output_dist(rows, ctx, group, expand=reverse, reduce_fn=reduce_fn)

