"""ORACLE — TEST INFRASTRUCTURE ONLY.  CPU/eager restatement of the HSTU jagged attention the product kernels implement.

Follows the reference's own checkers:
  * mask: third_party/FBGEMM/fbgemm_gpu/experimental/hstu/test/hstu_test.py:86-171 (`construct_mask`) — identical, for the
    causal + target(+group) + context cases, to examples/hstu/ops/pt_ops/pt_hstu_attention.py:46-105 (`_get_valid_attn_mask`);
  * math: hstu_test.py:553-634 (`_hstu_attention_maybe_from_cache`): pad -> einsum QK^T -> *alpha -> silu -> /scaling_seqlen
    -> *mask -> einsum PV -> unpad, in fp32 (upcast=True) or in the input dtype (the "torch bf16" run that sets the tolerance);
    `scaling_seqlen` as in pt_hstu_attention.py:150-196.
Parity: pinned by tests/golden/hstu_*.npz, produced by tests/golden/gen_golden_cpu.py from the reference's own Python
(`pytorch_hstu_mha` and `construct_mask`, executed unmodified).  Gradients come from autograd of this function.
Tolerance rule (hstu_test.py:885,956-964; examples/commons/utils/hstu_assert_close.py:42-57):
  fwd: max|kernel - ref32| <= 2 * max|ref_bf16 - ref32|;  dq/dk/dv: <= 5 *.
"""
import math
from typing import Optional, Tuple

import torch
import torch.nn.functional as F


def build_mask(seqlens, N: int, num_contexts=None, num_targets=None, target_group_size: int = 1, window: Tuple[int, int] = (-1, 0)) -> torch.Tensor:
    """bool [B, N, N]; entry [b, i, j] = query i may attend key j (both < seqlen_b)."""
    B = len(seqlens)
    mask = torch.zeros(B, N, N, dtype=torch.bool)
    wl, wr = window
    rows = torch.arange(N).view(N, 1)
    cols = torch.arange(N).view(1, N)
    for b in range(B):
        L = int(seqlens[b])
        nc = int(num_contexts[b]) if num_contexts is not None else 0
        nt = int(num_targets[b]) if num_targets is not None else 0
        m = torch.ones(N, N, dtype=torch.bool)
        if wr >= 0:
            m &= cols <= rows + wr
        if wl >= 0:
            m &= cols >= rows - wl
        if wl < 0 and wr == 0:   # causal: target / context rules (hstu_test.py:118-152)
            h = L - nt            # history end = target start (contexts count as history)
            if nt > 0:
                tgt = (rows >= h) & (cols >= h) & (cols < h + ((rows - h) // target_group_size) * target_group_size)
                m &= ~tgt
            if nc > 0:
                m |= (rows < nc) & (cols < h)
        m &= (rows < L) & (cols < L)
        mask[b] = m
    return mask


def _pad(x, cu, B, N):
    T, H, D = x.shape
    out = x.new_zeros(B, N, H, D)
    for b in range(B):
        s, e = int(cu[b]), int(cu[b + 1])
        out[b, : e - s] = x[s:e]
    return out


def _unpad(x, cu, B):
    return torch.cat([x[b, : int(cu[b + 1]) - int(cu[b])] for b in range(B)], dim=0)


def hstu_attention(q, k, v, cu_seqlens, max_seqlen: int, alpha: float, scaling_seqlen: int = -1, num_contexts=None, num_targets=None,
                   target_group_size: int = 1, window: Tuple[int, int] = (-1, 0), upcast: bool = True) -> torch.Tensor:
    """q,k,v: (T,H,D).  Returns (T,H,D) in q.dtype.  Differentiable."""
    cu = cu_seqlens.tolist() if torch.is_tensor(cu_seqlens) else list(cu_seqlens)
    B = len(cu) - 1
    N = max_seqlen
    if scaling_seqlen is None or scaling_seqlen <= 0:
        scaling_seqlen = max_seqlen
    dt = q.dtype
    pq, pk, pv = _pad(q, cu, B, N), _pad(k, cu, B, N), _pad(v, cu, B, N)
    if upcast:
        pq, pk, pv = pq.float(), pk.float(), pv.float()
    s = torch.einsum("bnhd,bmhd->bhnm", pq, pk) * alpha
    p = F.silu(s) / scaling_seqlen
    mask = build_mask([cu[i + 1] - cu[i] for i in range(B)], N, num_contexts, num_targets, target_group_size, window).to(q.device)
    p = p * mask.unsqueeze(1).to(p.dtype)
    o = torch.einsum("bhnm,bmhd->bnhd", p, pv)
    return _unpad(o, cu, B).to(dt)


def fwd_bwd(q, k, v, dout, cu_seqlens, max_seqlen, alpha, scaling_seqlen=-1, num_contexts=None, num_targets=None, target_group_size=1,
            window=(-1, 0), upcast=True):
    q, k, v = (t.detach().clone().requires_grad_(True) for t in (q, k, v))
    o = hstu_attention(q, k, v, cu_seqlens, max_seqlen, alpha, scaling_seqlen, num_contexts, num_targets, target_group_size, window, upcast)
    o.backward(dout.to(o.dtype))
    return o.detach(), q.grad, k.grad, v.grad


def fwd_flops(seqlens, H: int, D: int, num_targets=None, num_contexts=None) -> float:
    """examples/commons/utils/perf.py:697-740 (_compute_attn_fwd_flops): 2 GEMMs over the unmasked area; causal: 2*H*D*L^2 per sequence."""
    tot = 0.0
    for i, L in enumerate(seqlens):
        nt = int(num_targets[i]) if num_targets is not None else 0
        nc = int(num_contexts[i]) if num_contexts is not None else 0
        hist = L - nt - nc
        tot += 4 * H * L * (nc + hist) * D - 2 * H * hist * hist * D + 4 * H * nt * D
    return tot
