"""`hstu_attn_varlen_func` — same 23-argument signature, validation errors and autograd behaviour as the reference
(third_party/FBGEMM/fbgemm_gpu/experimental/hstu/hstu/cuda_hstu_attention.py:677-775, HstuAttnVarlenFunc :248-670).
One deliberate fix: `scaling_seqlen` IS honoured (the reference's sm100 path silently scales by 1/max_seqlen_q,
hstu_blackwell/hstu_fwd.py:1777, unlike its own sm80/sm90 kernels)."""
from typing import Optional, Tuple

import torch

from . import hstu_ops_gpu


class HstuAttnVarlenFunc(torch.autograd.Function):
    @staticmethod
    def forward(ctx, q, k, v, cu_seqlens_q, cu_seqlens_k, seqused_q, seqused_k, max_seqlen_q, max_seqlen_k, scaling_seqlen, num_contexts,
                num_targets, target_group_size, window_size=(-1, -1), alpha=1.0, rab=None, has_drab=False, func=None, kv_cache=None,
                page_offsets=None, page_ids=None, last_page_lens=None, quant_mode=-1):
        assert q.dim() == 3, "q shape should be (L, num_heads, head_dim)"
        assert k.dim() == 3, "k shape should be (L, num_heads, head_dim)"
        assert v.dim() == 3, "v shape should be (L, num_heads, hidden_dim)"
        assert seqused_q is None and seqused_k is None, "HSTU-Blackwell does not support seqused_q and seqused_k"
        assert rab is None and not has_drab, "rab is not supported on sm100"
        assert quant_mode in (-1, None), "fp8 is not supported on sm100"
        out, _ = hstu_ops_gpu.hstu_varlen_fwd_100(q, k, v, cu_seqlens_q, cu_seqlens_k, max_seqlen_q, max_seqlen_k, num_contexts, num_targets,
                                                  target_group_size, window_size[0], window_size[1], alpha, rab, func, paged_kv=kv_cache,
                                                  page_ids=page_ids, page_indptrs=page_offsets, scaling_seqlen=scaling_seqlen)
        ctx.save_for_backward(q, k, v, cu_seqlens_q, cu_seqlens_k, num_contexts, num_targets)
        ctx.max_seqlen_q, ctx.max_seqlen_k, ctx.scaling_seqlen = max_seqlen_q, max_seqlen_k, scaling_seqlen
        ctx.target_group_size, ctx.window_size, ctx.alpha = target_group_size, window_size, alpha
        return out

    @staticmethod
    def backward(ctx, dout):
        q, k, v, cu_q, cu_k, num_contexts, num_targets = ctx.saved_tensors
        dq, dk, dv, _ = hstu_ops_gpu.hstu_varlen_bwd_100(dout, q, k, v, cu_q, cu_k, ctx.max_seqlen_q, ctx.max_seqlen_k, None, None, None,
                                                         num_contexts, num_targets, ctx.target_group_size, ctx.window_size[0],
                                                         ctx.window_size[1], ctx.alpha, None, False, None, False,
                                                         scaling_seqlen=ctx.scaling_seqlen)
        return (dq, dk, dv) + (None,) * 20


def hstu_attn_varlen_func(
    q: torch.Tensor,
    k: torch.Tensor,
    v: torch.Tensor,
    cu_seqlens_q: torch.Tensor,
    cu_seqlens_k: torch.Tensor,
    seqused_q: Optional[torch.Tensor],
    seqused_k: Optional[torch.Tensor],
    max_seqlen_q: int,
    max_seqlen_k: int,
    scaling_seqlen: int,
    num_contexts: torch.Tensor,
    num_targets: torch.Tensor,
    target_group_size: int = 1,
    window_size: Tuple[int, int] = (-1, -1),
    alpha: float = 1.0,
    rab: Optional[torch.Tensor] = None,
    has_drab: bool = False,
    kv_cache: Optional[torch.Tensor] = None,
    page_offsets: Optional[torch.Tensor] = None,
    page_ids: Optional[torch.Tensor] = None,
    last_page_lens: Optional[torch.Tensor] = None,
    func: Optional[torch.Tensor] = None,
    quant_mode: Optional[int] = -1,
):
    """q,k,v: (total, nheads, headdim) bf16; cu_seqlens_*: (batch+1,) int32; returns out (total, nheads, headdim).
    See the reference docstring (cuda_hstu_attention.py:702-728) for every argument."""
    if has_drab and (rab is None):
        raise ValueError("AssertError: rab is None, but has_drab is True, is not allowed in backward")
    if num_contexts is not None and window_size != (-1, 0):
        raise ValueError("AssertError: context is True and causal is not True, this is undefined behavior")
    if num_targets is not None and window_size != (-1, 0):
        raise ValueError("AssertError: target is True and causal is not True, this is undefined behavior")
    if num_targets is None and target_group_size < 1:
        raise ValueError("AssertError: target_group_size should be greater than 0 when target is True")
    if max_seqlen_q > max_seqlen_k:
        raise ValueError("AssertError: seq_len_q >= seq_len_k, this is undefined behavior")
    return HstuAttnVarlenFunc.apply(q, k, v, cu_seqlens_q, cu_seqlens_k, seqused_q, seqused_k, max_seqlen_q, max_seqlen_k, scaling_seqlen,
                                    num_contexts, num_targets, target_group_size, window_size, alpha, rab, has_drab, func, kv_cache,
                                    page_offsets, page_ids, last_page_lens, quant_mode)
