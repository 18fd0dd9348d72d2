"""GPU parity of the HSTU attention kernels against the eager oracle with the reference's own tolerance rule
(third_party/FBGEMM/.../hstu/test/hstu_test.py:885,956-964; examples/commons/utils/hstu_assert_close.py:42-57):
    fwd : max|kernel - ref_fp32| <= 2 x max|torch_bf16 - ref_fp32|        dq/dk/dv : <= 5 x
Cases follow hstu_test.py:645-725 (jagged batches, targets / groups / contexts, local windows, alpha, hdim 64/128) and
test_fwd_qkv_noclone.py (q/k/v as strided views of one fused buffer)."""
import math

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_sm100_descriptor_probe(cuda):
    """Pins the hand-encoded UMMA descriptors (K-major / MN-major SWIZZLE_128B, LBO/SBO, k-step advance) and the TMEM layout."""
    from hstu import hstu_ops_gpu as ops
    torch.manual_seed(0)
    A = torch.randn(128, 128, device=cuda).to(torch.bfloat16)
    Bm = torch.randn(128, 128, device=cuda).to(torch.bfloat16)
    Af, Bf = A.float(), Bm.float()
    want = {0: Af @ Bf.t(), 1: Af @ Bf, 2: Af.t() @ Bf.t(), 3: Af.t() @ Bf, 4: Af @ Bf.t(), 5: Af @ Bf}   # 4/5: A operand in tensor memory (.ts)
    msgs = []
    for v in range(6):
        C = ops.probe_gemm(A, Bm, v)
        torch.cuda.synchronize()
        err = (C - want[v]).abs().max().item()
        msgs.append(f"variant {v}: max err {err:.4g}")
    print("\n".join(msgs))
    for v in range(6):
        C = ops.probe_gemm(A, Bm, v)
        assert torch.allclose(C, want[v], rtol=1e-3, atol=1e-2), msgs[v]


def _inputs(cuda, lens, H, D, fused=True, seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    T = int(sum(lens))
    if fused:   # u, v, q, k views of one (T, 4*H*D) buffer, as fused_hstu_op.py:494-501
        buf = torch.randn(T, 4 * H * D, generator=g).to(torch.bfloat16).to(cuda)
        _, v, q, k = (t.view(T, H, D) for t in buf.split(H * D, dim=-1))
    else:
        q, k, v = (torch.randn(T, H, D, generator=g).to(torch.bfloat16).to(cuda) for _ in range(3))
    dout = torch.randn(T, H, D, generator=g).to(torch.bfloat16).to(cuda)
    cu = torch.tensor(np.concatenate([[0], np.cumsum(lens)]), dtype=torch.int32, device=cuda)
    return q, k, v, dout, cu


def _check(cuda, lens, H, D, nt=None, nc=None, G=1, window=(-1, 0), alpha=None, scaling=-1, fused=True, seed=0):
    from hstu import hstu_attn_varlen_func
    from oracle import hstu_attn as orc
    alpha = 1.0 / math.sqrt(D) if alpha is None else alpha
    q, k, v, dout, cu = _inputs(cuda, lens, H, D, fused, seed)
    N = int(max(lens))
    ntt = None if nt is None else torch.tensor(nt, dtype=torch.int32, device=cuda)
    nct = None if nc is None else torch.tensor(nc, dtype=torch.int32, device=cuda)
    qq, kk, vv = (t.detach().clone().requires_grad_(True) if not fused else t.detach().requires_grad_(True) for t in (q, k, v))
    out = hstu_attn_varlen_func(qq, kk, vv, cu, cu, None, None, N, N, scaling, nct, ntt, G, window, alpha)
    out.backward(dout)
    torch.cuda.synchronize()
    ref = orc.fwd_bwd(q.float().cpu(), k.float().cpu(), v.float().cpu(), dout.float().cpu(), cu.cpu(), N, alpha, scaling, nc, nt, G, window, upcast=True)
    low = orc.fwd_bwd(q.cpu(), k.cpu(), v.cpu(), dout.cpu(), cu.cpu(), N, alpha, scaling, nc, nt, G, window, upcast=False)
    got = (out, qq.grad, kk.grad, vv.grad)
    names = ("out", "dq", "dk", "dv")
    report = []
    ok = True
    for nm, g_, r32, r16, mult in zip(names, got, ref, low, (2, 5, 5, 5)):
        assert g_ is not None, nm
        err = (g_.float().cpu() - r32.float()).abs().max().item()
        base = (r16.float() - r32.float()).abs().max().item()
        report.append(f"{nm}: err {err:.3e} vs {mult}x bf16-eager {base:.3e}")
        ok &= err <= mult * base + 1e-6 and math.isfinite(err)
    assert ok, " | ".join(report)
    return report


@pytest.mark.parametrize("D", [128, 64])
def test_causal_jagged(cuda, D):
    _check(cuda, [257, 1, 128, 640, 63, 129, 32, 500], 2, D)


def test_single_tile_and_exact_multiples(cuda):
    _check(cuda, [128], 1, 128, fused=False)
    _check(cuda, [256, 384], 2, 128)
    _check(cuda, [5], 1, 64)


@pytest.mark.parametrize("G", [1, 3])
def test_targets(cuda, G):
    _check(cuda, [300, 77, 512, 140], 2, 128, nt=[40, 0, 100, 139], G=G)


def test_contexts_and_targets(cuda):
    _check(cuda, [300, 200, 513], 2, 128, nt=[20, 0, 64], nc=[5, 3, 0], G=2)
    _check(cuda, [400, 150], 1, 64, nc=[130, 7])          # contexts spanning more than one 64/128-row tile


def test_local_window_and_full(cuda):
    _check(cuda, [300, 260], 2, 128, window=(100, 0))
    _check(cuda, [300, 260], 2, 64, window=(64, 33))
    _check(cuda, [200, 129], 2, 128, window=(-1, -1))


def test_alpha_and_scaling_seqlen(cuda):
    _check(cuda, [333, 200], 2, 128, alpha=0.1, scaling=1000)       # scaling_seqlen honoured (reference sm100 ignores it)
    _check(cuda, [333, 200], 2, 128, alpha=1.0)


def test_long_sequence(cuda):
    _check(cuda, [2048, 1500], 4, 128, nt=[256, 0])


def test_argument_validation(cuda):
    from hstu import hstu_attn_varlen_func
    q, k, v, dout, cu = _inputs(cuda, [64], 1, 64)
    nt = torch.tensor([3], dtype=torch.int32, device=cuda)
    with pytest.raises(ValueError):
        hstu_attn_varlen_func(q, k, v, cu, cu, None, None, 64, 64, -1, None, nt, 1, (-1, -1), 1.0)     # targets need causal
    with pytest.raises(ValueError):
        hstu_attn_varlen_func(q, k, v, cu, cu, None, None, 64, 32, -1, None, None, 1, (-1, 0), 1.0)    # max_seqlen_q > max_seqlen_k
    with pytest.raises(ValueError):
        hstu_attn_varlen_func(q, k, v, cu, cu, None, None, 64, 64, -1, None, None, 1, (-1, 0), 1.0, None, True)   # has_drab without rab
