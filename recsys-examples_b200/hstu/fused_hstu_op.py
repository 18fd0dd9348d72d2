"""FusedHSTULayerFunction / fused_hstu_op — one HSTU layer as a single autograd function.

Mirror of /root/reference/examples/hstu/ops/fused_hstu_op.py:75-1180 (same argument list, same saved-tensor / recompute switches, same
returned gradients) on this library's kernels:

    y = layer_norm(input, input_norm_weight, input_norm_bias)          hstu_layer_norm_fwd            (csrc/hstu_glue.cu)
    y = addmm(linear_uvqk_bias, y, linear_uvqk_weight)                 cuBLAS (torch.addmm) — on sm_100 the reference also uses torch's
    y = silu(y)                                                        hstu_silu_fwd                   addmm + silu here (fused_hstu_op.py:43-48)
    u, v, q, k = split(y)                                              views, consumed in place
    attn = hstu_attn(q, k, v)                                          hstu_fwd_sm100                  (csrc/hstu_fwd.cu, tcgen05)
    y = dropout(layer_norm(attn, out_w, out_b) * u)                    hstu_ln_mul_dropout_fwd
    out = addmm(input, y, linear_proj_weight)                          cuBLAS

Backward, where it differs from the reference's op order:
  * the output-norm backward kernel also RECOMPUTES y (the proj GEMM's weight-gradient operand): y is not saved by the forward;
  * the attention backward writes dq / dk / dv contiguous and the SiLU backward kernel gathers du / dv / dq / dk from where they are
    (the reference lets its attention kernel write into slices of one uvqk gradient buffer; ours needs contiguous outputs);
  * the input-norm backward adds the residual gradient into dx in the same pass (the reference: dx_accumulate, same thing);
  * the uvqk bias gradient (column sums of the SiLU backward's output) is accumulated by the SiLU backward kernel itself (the reference sums
    dz in a separate pass, triton_addmm.py:293-294).
`wgrad_stream` / `wgrad_event` (weight-gradient GEMMs on a side stream) are honoured the way the reference does it: the two weight-gradient
GEMMs run on `wgrad_stream` when given, and `wgrad_event` is recorded after them.
Limits: Blackwell attention kernel only (attn_backend is accepted and ignored), linear_dim_per_head == attention_dim_per_head in (64, 128),
hidden size and heads * linear_dim <= 8192 (<= 1024: warp-per-row glue kernels, above: CTA-per-row), no contextual tokens (as the reference on sm_100, :60-71).
"""
from typing import Optional, Union

import torch

from . import hstu_ops_gpu as _attn
from .layer_ops import (layer_norm_mul_dropout_bwd, layer_norm_mul_dropout_fwd, silu_bwd_segments, silu_fwd, weighted_layer_norm_bwd,
                        weighted_layer_norm_fwd)


def _no_contexts(num_contexts):
    if num_contexts is None:
        return None
    if isinstance(num_contexts, int):
        if num_contexts == 0:
            return None
        raise ValueError("Blackwell fused_hstu_op does not support contextual tokens")
    if torch.count_nonzero(num_contexts).item() == 0:
        return None
    raise ValueError("Blackwell fused_hstu_op does not support contextual tokens")


class FusedHSTULayerFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, input, seqlen_offsets, max_seqlen, scaling_seqlen, linear_uvqk_weight, linear_uvqk_bias, linear_proj_weight,  # noqa: A002
                num_heads, linear_dim_per_head, attention_dim_per_head, ln_eps, dropout_ratio, training, input_norm_weight=None,
                input_norm_bias=None, output_norm_weight=None, output_norm_bias=None, attn_backend=None, num_targets=None,
                num_contextuals=None, target_group_size=1, alpha=1.0, causal=True, seed=None, residual=True, wgrad_stream=None,
                wgrad_event=None, recompute_input_layernorm=False, recompute_input_silu=False):
        assert input.dim() == 2, "input tensor must be 2D"
        assert linear_uvqk_bias.dim() == 1, "linear_uvqk_bias must be 1D"
        learn_in, learn_out = input_norm_weight is not None, output_norm_weight is not None
        assert not learn_in or input_norm_bias is not None
        assert learn_out and output_norm_bias is not None, "output_norm_weight / output_norm_bias must be provided"
        assert causal, "the fused layer is causal (window (-1, 0))"
        if linear_dim_per_head != attention_dim_per_head:
            raise ValueError("linear_dim_per_head must equal attention_dim_per_head")
        if attention_dim_per_head not in (64, 128):
            raise ValueError(f"Blackwell fwd only supports head_dim in (64, 128), got {attention_dim_per_head}")
        num_contexts = _no_contexts(num_contextuals)
        H, Dh = num_heads, attention_dim_per_head
        split = [linear_dim_per_head * H, linear_dim_per_head * H, Dh * H, Dh * H]
        T = input.shape[0]

        normed, in_mean, in_rstd, _, _ = weighted_layer_norm_fwd(input, input_norm_weight, input_norm_bias, ln_eps)
        pre = torch.addmm(linear_uvqk_bias, normed, linear_uvqk_weight)                 # silu_input
        act = silu_fwd(pre)
        tu, tv, tq, tk = torch.split(act, split, dim=-1)
        q, k, v = tq.view(T, H, Dh), tk.view(T, H, Dh), tv.view(T, H, Dh)
        attn, _ = _attn.hstu_varlen_fwd_100(q, k, v, seqlen_offsets, seqlen_offsets, max_seqlen, max_seqlen, num_contexts, num_targets,
                                            target_group_size, -1, 0, alpha, scaling_seqlen=scaling_seqlen)
        attn2d = attn.view(T, H * Dh)
        y, out_mean, out_rstd, _, _, used_seed = layer_norm_mul_dropout_fwd(attn2d, tu, output_norm_weight, output_norm_bias, ln_eps, dropout_ratio,
                                                                           training, False, seed)
        out = torch.addmm(input, y, linear_proj_weight) if residual else torch.mm(y, linear_proj_weight)

        ctx.save_for_backward(input, input_norm_weight, input_norm_bias, in_mean, in_rstd, None if recompute_input_layernorm else normed,
                              linear_uvqk_weight, pre, None if recompute_input_silu else act, attn2d, output_norm_weight, output_norm_bias,
                              out_mean, out_rstd, linear_proj_weight, seqlen_offsets, num_targets)
        ctx.cfg = dict(H=H, Dh=Dh, split=split, eps=ln_eps, dropout_ratio=dropout_ratio, training=training, seed=used_seed, residual=residual,
                       max_seqlen=max_seqlen, scaling_seqlen=scaling_seqlen, target_group_size=target_group_size, alpha=alpha,
                       learn_in=learn_in, wgrad_stream=wgrad_stream, wgrad_event=wgrad_event, has_bias_grad=linear_uvqk_bias.requires_grad)
        return out

    @staticmethod
    def backward(ctx, grad_output):
        (input, in_w, in_b, in_mean, in_rstd, normed, w_uvqk, pre, act, attn2d, out_w, out_b, out_mean, out_rstd, w_proj, seqlen_offsets,  # noqa: A001
         num_targets) = ctx.saved_tensors
        c = ctx.cfg
        H, Dh, T = c["H"], c["Dh"], input.shape[0]
        grad_output = grad_output.contiguous()
        cur = torch.cuda.current_stream()
        ws_stream = c["wgrad_stream"]

        def wgrad(a_t, b_):
            """a_t^T @ b_ on the weight-gradient stream when one was given (its inputs are ready on the current stream now)."""
            if ws_stream is None:
                return torch.mm(a_t.t(), b_)
            ws_stream.wait_stream(cur)
            with torch.cuda.stream(ws_stream):
                r = torch.mm(a_t.t(), b_)
            a_t.record_stream(ws_stream); b_.record_stream(ws_stream)
            return r

        if act is None:                                                   # recompute_input_silu
            act = silu_fwd(pre)
        tu, tv, tq, tk = torch.split(act, c["split"], dim=-1)
        # 1. proj GEMM dgrad; its wgrad needs y, which the next kernel recomputes
        dy = torch.mm(grad_output, w_proj.t())
        dattn, du, d_out_w, d_out_b, y = layer_norm_mul_dropout_bwd(dy, attn2d, tu, out_w, out_b, out_mean, out_rstd, 0, 0, c["eps"], c["training"],
                                                                    c["dropout_ratio"], c["seed"], False, True)
        d_w_proj = wgrad(y, grad_output)
        # 2. attention backward
        dq, dk, dv, _ = _attn.hstu_varlen_bwd_100(dattn.view(T, H, Dh), tq.view(T, H, Dh), tk.view(T, H, Dh), tv.view(T, H, Dh), seqlen_offsets,
                                                  seqlen_offsets, c["max_seqlen"], c["max_seqlen"], None, None, None, None, num_targets,
                                                  c["target_group_size"], -1, 0, c["alpha"], scaling_seqlen=c["scaling_seqlen"])
        # 3. silu backward over (du | dv | dq | dk), read in place -> gradient of the uvqk GEMM output
        if c["has_bias_grad"]:                                            # the bias gradient is the column sum of dpre: same kernel
            dpre, d_bias = silu_bwd_segments([du, dv.view(T, H * Dh), dq.view(T, H * Dh), dk.view(T, H * Dh)], pre, with_bias_grad=True)
            d_bias = d_bias.to(pre.dtype)
        else:
            dpre, d_bias = silu_bwd_segments([du, dv.view(T, H * Dh), dq.view(T, H * Dh), dk.view(T, H * Dh)], pre), None
        d_normed = torch.mm(dpre, w_uvqk.t())
        if normed is None:                                                # recompute_input_layernorm
            normed = weighted_layer_norm_fwd(input, in_w, in_b, c["eps"])[0]
        d_w_uvqk = wgrad(normed, dpre)
        # 4. input layer norm backward (+ the residual branch's gradient, same pass)
        d_input, d_in_w, d_in_b = weighted_layer_norm_bwd(d_normed, input, in_w, in_b, in_mean, in_rstd, c["learn_in"], c["eps"], 0, 0,
                                                          grad_output if c["residual"] else None)
        if ws_stream is not None:
            if c["wgrad_event"] is not None:
                c["wgrad_event"].record(ws_stream)
            cur.wait_stream(ws_stream)                                    # autograd consumes the weight gradients on the current stream
        return (d_input, None, None, None, d_w_uvqk, d_bias, d_w_proj, None, None, None, None, None, None, d_in_w, d_in_b, d_out_w, d_out_b,
                None, None, None, None, None, None, None, None, None, None, None, None)


def fused_hstu_op(input: torch.Tensor, seqlen_offsets: torch.Tensor, max_seqlen: int, scaling_seqlen: int, linear_uvqk_weight: torch.Tensor,  # noqa: A002
                  linear_uvqk_bias: torch.Tensor, linear_proj_weight: torch.Tensor, num_heads: int, linear_dim_per_head: int,
                  attention_dim_per_head: int, ln_eps: float, dropout_ratio: float, training: bool,
                  input_norm_weight: Optional[torch.Tensor] = None, input_norm_bias: Optional[torch.Tensor] = None,
                  output_norm_weight: Optional[torch.Tensor] = None, output_norm_bias: Optional[torch.Tensor] = None, attn_backend=None,
                  num_targets: Optional[torch.Tensor] = None, num_contextuals: Union[int, Optional[torch.Tensor]] = None,
                  target_group_size: int = 1, alpha: float = 1.0, causal: bool = True, seed: Optional[int] = None, residual: bool = True,
                  wgrad_stream: Optional[torch.cuda.Stream] = None, wgrad_event: Optional[torch.cuda.Event] = None,
                  recompute_input_layernorm: bool = False, recompute_input_silu: bool = False) -> torch.Tensor:
    """examples/hstu/ops/fused_hstu_op.py:1104-1180 — same positional order."""
    return FusedHSTULayerFunction.apply(input, seqlen_offsets, max_seqlen, scaling_seqlen, linear_uvqk_weight, linear_uvqk_bias, linear_proj_weight,
                                        num_heads, linear_dim_per_head, attention_dim_per_head, ln_eps, dropout_ratio, training, input_norm_weight,
                                        input_norm_bias, output_norm_weight, output_norm_bias, attn_backend, num_targets, num_contextuals,
                                        target_group_size, alpha, causal, seed, residual, wgrad_stream, wgrad_event, recompute_input_layernorm,
                                        recompute_input_silu)
