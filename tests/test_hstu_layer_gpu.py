"""GPU: the HSTU layer glue kernels (csrc/hstu_glue.cu) and FusedHSTULayerFunction against plain PyTorch fp32 references of the same ops.

The reference checks its Triton kernels the same way (examples/hstu/test/test_hstu_layer.py / test_hstu_op.py: fused op vs the eager
`pytorch_norm_mul_dropout` / F.layer_norm / F.silu path, bf16 tolerances).  Tolerances here: fp32 I/O 2e-5 (one pass, fp32 math on both
sides); bf16 / fp16 I/O: the kernel's fp32 result rounded once must lie within one output ulp of the fp32 reference computed from the same
(rounded) inputs -> rtol 2^-7 (bf16) / 2^-10 (fp16) plus a small atol; weight / bias gradients (sums over rows): relative to the column's
sum of magnitudes."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

DT = {"fp32": torch.float32, "bf16": torch.bfloat16, "fp16": torch.float16}
TOL = {"fp32": dict(rtol=2e-5, atol=2e-5), "bf16": dict(rtol=2 ** -7, atol=2e-2), "fp16": dict(rtol=2 ** -10, atol=3e-3)}


def _mk(cuda, shape, dtype, seed, scale=1.0, shift=0.0):
    g = torch.Generator(device=cuda).manual_seed(seed)
    return (torch.randn(*shape, device=cuda, generator=g) * scale + shift).to(dtype)


@pytest.mark.parametrize("dt", ["fp32", "bf16", "fp16"])
@pytest.mark.parametrize("rows,D", [(1, 8), (37, 256), (513, 264), (300, 512), (4099, 1024), (129, 1000),
                                    (300, 1032), (257, 2048), (700, 4096), (33, 8192), (1, 5000)])      # > 1024: the CTA-per-row kernels
@pytest.mark.parametrize("learnable", [True, False])
def test_layer_norm_fwd_bwd_matches_torch(cuda, dt, rows, D, learnable):
    from hstu import layer_ops as L
    dtype, eps = DT[dt], 1e-5
    x = _mk(cuda, (rows, D), dtype, 1, 2.0, 0.5)
    w = _mk(cuda, (D,), dtype, 2, 0.5, 1.0) if learnable else None
    b = _mk(cuda, (D,), dtype, 3, 0.5) if learnable else None
    dy = _mk(cuda, (rows, D), dtype, 4)
    res = _mk(cuda, (rows, D), dtype, 5)
    y, mean, rstd, _, _ = L.triton_weighted_layer_norm_fwd(x, w, b, eps)
    xr = x.float().requires_grad_(True)
    wr = w.float().requires_grad_(True) if learnable else None
    br = b.float().requires_grad_(True) if learnable else None
    yr = F.layer_norm(xr, (D,), wr, br, eps)
    torch.testing.assert_close(y.float(), yr.detach(), **TOL[dt])
    torch.testing.assert_close(mean, x.float().mean(1), rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(rstd, 1.0 / torch.sqrt(x.float().var(1, unbiased=False) + eps), rtol=1e-4, atol=1e-5)
    yr.backward(dy.float())
    for acc in (None, res):
        dx, dw, db = L.triton_weighted_layer_norm_bwd(dy, x, w, b, mean, rstd, learnable, eps, 0, 0, acc)
        want = xr.grad + (acc.float() if acc is not None else 0)
        torch.testing.assert_close(dx.float(), want, **TOL[dt])
        if learnable:
            scale_w = (dy.float().abs() * ((x.float() - mean[:, None]) * rstd[:, None]).abs()).sum(0)
            assert bool(((dw.float() - wr.grad).abs() <= TOL[dt]["rtol"] * scale_w + 1e-4).all())
            assert bool(((db.float() - br.grad).abs() <= TOL[dt]["rtol"] * dy.float().abs().sum(0) + 1e-4).all())
        else:
            assert dw is None and db is None


def test_layer_norm_strided_rows_and_errors(cuda):
    from hstu import layer_ops as L
    big = _mk(cuda, (200, 4096), torch.bfloat16, 7)
    x = big[:, 1024:2048]                                   # a slice of a fused buffer: row stride 4096, read in place
    w, b = _mk(cuda, (1024,), torch.bfloat16, 8, 0.1, 1.0), _mk(cuda, (1024,), torch.bfloat16, 9, 0.1)
    y, mean, rstd, _, _ = L.weighted_layer_norm_fwd(x, w, b, 1e-6)
    y2 = L.weighted_layer_norm_fwd(x.contiguous(), w, b, 1e-6)[0]
    assert torch.equal(y, y2)
    dy = _mk(cuda, (200, 4096), torch.bfloat16, 10)[:, 2048:3072]
    dx1 = L.weighted_layer_norm_bwd(dy, x, w, b, mean, rstd, True, 1e-6)
    dx2 = L.weighted_layer_norm_bwd(dy.contiguous(), x.contiguous(), w, b, mean, rstd, True, 1e-6)
    assert all(torch.equal(a, c) for a, c in zip(dx1, dx2))
    with pytest.raises(ValueError, match="<= 8192"):
        L.weighted_layer_norm_fwd(_mk(cuda, (4, 16384), torch.bfloat16, 1), None, None, 1e-5)
    with pytest.raises(ValueError):
        L.weighted_layer_norm_fwd(_mk(cuda, (4, 100), torch.bfloat16, 1), None, None, 1e-5)        # width not a multiple of 8
    e = L.weighted_layer_norm_fwd(torch.empty(0, 64, device=cuda, dtype=torch.bfloat16), None, None, 1e-5)
    assert e[0].shape == (0, 64)


@pytest.mark.parametrize("dt", ["fp32", "bf16"])
@pytest.mark.parametrize("rows,D", [(65, 128), (1000, 1024), (777, 520), (200, 2048), (64, 4104), (40, 8192)])
@pytest.mark.parametrize("ratio,training", [(0.0, True), (0.25, True), (0.25, False)])
def test_ln_mul_dropout_fwd_bwd_matches_torch(cuda, dt, rows, D, ratio, training):
    """y = dropout(LN(x) * u) with the kernel's own mask (exported by the C ABI) applied to the torch reference
    (pytorch_norm_mul_dropout, examples/hstu/ops/pt_ops/pt_norm_mul_dropout.py:19-64, with a fixed mask instead of F.dropout)."""
    from hstu import layer_ops as L
    dtype, eps, seed = DT[dt], 1e-5, 1234567
    x = _mk(cuda, (rows, D), dtype, 11, 1.5, -0.3)
    ubuf = _mk(cuda, (rows, 4 * D), dtype, 12)
    u = ubuf[:, :D]                                           # strided row view, like the u slice of the uvqk activation
    w, b = _mk(cuda, (D,), dtype, 13, 0.3, 1.0), _mk(cuda, (D,), dtype, 14, 0.3)
    dy = _mk(cuda, (rows, D), dtype, 15)
    y, mean, rstd, _, _, used = L.triton_layer_norm_mul_dropout_fwd(x, u, w, b, eps, ratio, training, False, seed)
    drop = training and ratio > 0
    if drop:
        keep = L.dropout_mask(rows, D, ratio, used, cuda)
        thr = round(ratio * 65536)
        scale = 1.0 / (1.0 - thr / 65536.0)
        frac = 1.0 - keep.float().mean().item()
        assert abs(frac - ratio) < 4 * math.sqrt(ratio * (1 - ratio) / (rows * D)) + 1e-3
        assert not torch.equal(keep, L.dropout_mask(rows, D, ratio, used + 1, cuda))
        m = keep.float() * scale
    else:
        m = torch.ones(rows, D, device=cuda)
    xr, ur = x.float().requires_grad_(True), u.float().requires_grad_(True)
    wr, br = w.float().requires_grad_(True), b.float().requires_grad_(True)
    yr = F.layer_norm(xr, (D,), wr, br, eps) * ur * m
    torch.testing.assert_close(y.float(), yr.detach(), **TOL[dt])
    if drop:
        assert bool((y[~keep] == 0).all())
    yr.backward(dy.float())
    du_buf = torch.zeros(rows, 2 * D, dtype=dtype, device=cuda)
    for du_out in (None, du_buf[:, D:]):                     # fresh buffer / strided slice of a larger gradient buffer
        dx, du, dw, db, y2 = L.triton_layer_norm_mul_dropout_bwd(dy, x, u, w, b, mean, rstd, 0, 0, eps, training, ratio, used, False, True, None, du_out)
        torch.testing.assert_close(dx.float(), xr.grad, **TOL[dt])
        torch.testing.assert_close(du.float(), ur.grad, **TOL[dt])
        torch.testing.assert_close(y2.float(), y.float(), rtol=2 ** -7 if dt == "bf16" else 1e-6, atol=1e-6)   # the recomputed forward output
        xh = (x.float() - mean[:, None]) * rstd[:, None]
        dln = (dy.float() * m * u.float()).abs()
        assert bool(((dw.float() - wr.grad).abs() <= TOL[dt]["rtol"] * (dln * xh.abs()).sum(0) + 1e-4).all())
        assert bool(((db.float() - br.grad).abs() <= TOL[dt]["rtol"] * dln.sum(0) + 1e-4).all())
    assert bool((du_buf[:, :D] == 0).all())


@pytest.mark.parametrize("dt", ["fp32", "bf16", "fp16"])
def test_silu_fwd_bwd_matches_torch(cuda, dt):
    from hstu import layer_ops as L
    dtype = DT[dt]
    rows, H, Dh = 333, 8, 128
    W = 4 * H * Dh
    x = _mk(cuda, (rows, W), dtype, 21, 3.0)
    y = L.triton_silu_fwd(x)
    torch.testing.assert_close(y.float(), F.silu(x.float()), **TOL[dt])
    # gradient pieces as the layer produces them: du a strided slice, dv / dq / dk contiguous (T, H, D)
    dub = _mk(cuda, (rows, 2 * H * Dh), dtype, 22)
    du = dub[:, H * Dh:]
    dv, dq, dk = (_mk(cuda, (rows, H, Dh), dtype, s) for s in (23, 24, 25))
    dx = L.silu_bwd_segments([du, dv.view(rows, -1), dq.view(rows, -1), dk.view(rows, -1)], x)
    xr = x.float().requires_grad_(True)
    g = torch.cat([du.float(), dv.view(rows, -1).float(), dq.view(rows, -1).float(), dk.view(rows, -1).float()], dim=1)
    F.silu(xr).backward(g)
    torch.testing.assert_close(dx.float(), xr.grad, **TOL[dt])
    torch.testing.assert_close(L.triton_silu_bwd(g.to(dtype), x).float(), xr.grad, **TOL[dt])


@pytest.mark.parametrize("dt", ["fp32", "bf16"])
@pytest.mark.parametrize("rows,W", [(777, 4096), (1000, 192), (33, 1000), (5, 8 * 2377)])
def test_silu_bwd_with_bias_grad(cuda, dt, rows, W):
    """The column sums of the SiLU backward's output (the bias gradient of the GEMM in front) accumulated by the same kernel; widths whose
    W / 8 cannot be tiled onto the launch (the last case) fall back to a separate sum."""
    from hstu import layer_ops as L
    dtype = DT[dt]
    x = _mk(cuda, (rows, W), dtype, 31, 2.0)
    cuts = [0, W // 8 * 2, W // 8 * 5, W] if W % 64 == 0 else [0, W]
    g = _mk(cuda, (rows, W), dtype, 32)
    segs = [g[:, a:b] for a, b in zip(cuts[:-1], cuts[1:])]
    dx, dbias = L.silu_bwd_segments(segs, x, with_bias_grad=True)
    xr = x.float().requires_grad_(True)
    F.silu(xr).backward(g.float())
    torch.testing.assert_close(dx.float(), xr.grad, **TOL[dt])
    assert dbias.dtype == torch.float32 and dbias.shape == (W,)
    assert torch.equal(dx, L.silu_bwd_segments(segs, x))                       # same dx with and without the column sums
    assert bool(((dbias - xr.grad.sum(0)).abs() <= TOL[dt]["rtol"] * xr.grad.abs().sum(0) + 1e-4).all())


def _layer_ref(x, cu, S, p, H, Dh, eps, alpha, num_targets=None):
    """The layer in fp32 torch ops with the oracle attention (oracle/hstu_attn.py)."""
    from oracle.hstu_attn import hstu_attention
    T = x.shape[0]
    n = F.layer_norm(x, (x.shape[1],), p["in_w"], p["in_b"], eps)
    uvqk = F.silu(torch.addmm(p["b_uvqk"], n, p["w_uvqk"]))
    u, v, q, k = uvqk.split(H * Dh, dim=-1)
    a = hstu_attention(q.reshape(T, H, Dh), k.reshape(T, H, Dh), v.reshape(T, H, Dh), cu, S, alpha, S, None, num_targets, 1, (-1, 0),
                       upcast=x.dtype == torch.float32)          # bf16 run: the "torch bf16" path that sets the tolerance
    y = F.layer_norm(a.reshape(T, H * Dh), (H * Dh,), p["out_w"], p["out_b"], eps) * u
    return torch.addmm(x, y, p["w_proj"])


@pytest.mark.parametrize("recompute", [(False, False), (True, True)])
@pytest.mark.parametrize("with_targets", [False, True])
def test_fused_hstu_layer_matches_fp32_reference(cuda, recompute, with_targets):
    """FusedHSTULayerFunction (bf16) against the same layer in fp32 torch + the oracle attention: the reference's acceptance rule
    (hstu_test.py:885,956-964 / commons/utils/hstu_assert_close.py:42-57): error of the fused bf16 op <= k x error of the bf16 eager path."""
    from hstu.fused_hstu_op import fused_hstu_op
    H, Dh, HID, eps = 4, 64, 256, 1e-5
    lens = [50, 128, 1, 77, 200]
    cu = torch.tensor([0] + list(torch.tensor(lens).cumsum(0)), dtype=torch.int32, device=cuda)
    T, S = int(cu[-1]), 256
    nt = torch.tensor([5, 0, 1, 10, 20], dtype=torch.int32, device=cuda) if with_targets else None
    alpha = 1.0 / math.sqrt(Dh)
    g = torch.Generator(device=cuda).manual_seed(0)
    rnd = lambda *s, sc=1.0: torch.randn(*s, device=cuda, generator=g) * sc        # noqa: E731
    p32 = dict(in_w=1 + rnd(HID, sc=0.1), in_b=rnd(HID, sc=0.1), w_uvqk=rnd(HID, 4 * H * Dh, sc=HID ** -0.5), b_uvqk=rnd(4 * H * Dh, sc=0.1),
               out_w=1 + rnd(H * Dh, sc=0.1), out_b=rnd(H * Dh, sc=0.1), w_proj=rnd(H * Dh, HID, sc=(H * Dh) ** -0.5))
    x32 = rnd(T, HID)
    dout32 = rnd(T, HID)
    bf = lambda t: t.to(torch.bfloat16)                                           # noqa: E731

    def run(dtype, fused):
        x = x32.detach().clone().to(dtype).requires_grad_(True)            # fresh leaves every run
        p = {k: v.detach().clone().to(dtype).requires_grad_(True) for k, v in p32.items()}
        if fused:
            out = fused_hstu_op(x, cu, S, S, p["w_uvqk"], p["b_uvqk"], p["w_proj"], H, Dh, Dh, eps, 0.0, True, p["in_w"], p["in_b"], p["out_w"], p["out_b"],
                                None, nt, None, 1, alpha, True, None, True, None, None, recompute[0], recompute[1])
        else:
            out = _layer_ref(x, cu, S, p, H, Dh, eps, alpha, nt)
        out.backward(dout32.to(dtype))
        return [out.detach().float(), x.grad.float()] + [p[k].grad.float() for k in sorted(p)]

    # references computed from the bf16-rounded parameters, in fp32
    x32 = bf(x32).float(); dout32 = bf(dout32).float(); p32 = {k: bf(v).float() for k, v in p32.items()}
    ref = run(torch.float32, False)
    eager = run(torch.bfloat16, False)
    ours = run(torch.bfloat16, True)
    names = ["out", "d_input"] + ["d_" + k for k in sorted(p32)]
    for name, r, e, o in zip(names, ref, eager, ours):
        err_e, err_o = (e - r).abs().max().item(), (o - r).abs().max().item()
        assert err_o <= 3 * err_e + 1e-3 * r.abs().max().item() + 1e-6, f"{name}: fused {err_o:.3e} vs eager bf16 {err_e:.3e}"


def test_fused_hstu_layer_dropout_and_wgrad_stream(cuda):
    """Dropout: same seed -> same result, backward consistent with the forward mask (finite-difference-free check: the gradient of sum(out)
    w.r.t. the proj weight equals y^T 1 with the y the backward recomputed); weight gradients on a side stream equal the in-stream ones."""
    from hstu.fused_hstu_op import fused_hstu_op
    H, Dh, HID, eps = 2, 64, 128, 1e-5
    cu = torch.tensor([0, 100, 164, 300], dtype=torch.int32, device=cuda)
    T, S = 300, 256
    g = torch.Generator(device=cuda).manual_seed(3)
    rnd = lambda *s, sc=1.0: (torch.randn(*s, device=cuda, generator=g) * sc).to(torch.bfloat16)     # noqa: E731
    base = dict(w_uvqk=rnd(HID, 4 * H * Dh, sc=0.1), b_uvqk=rnd(4 * H * Dh, sc=0.1), w_proj=rnd(H * Dh, HID, sc=0.1), in_w=1 + rnd(HID, sc=0.1),
                in_b=rnd(HID, sc=0.1), out_w=1 + rnd(H * Dh, sc=0.1), out_b=rnd(H * Dh, sc=0.1))
    x0 = rnd(T, HID)

    def run(seed, ratio, stream=None):
        x = x0.clone().requires_grad_(True)
        p = {k: v.clone().requires_grad_(True) for k, v in base.items()}
        ev = torch.cuda.Event() if stream is not None else None
        out = fused_hstu_op(x, cu, S, S, p["w_uvqk"], p["b_uvqk"], p["w_proj"], H, Dh, Dh, eps, ratio, True, p["in_w"], p["in_b"], p["out_w"], p["out_b"],
                            None, None, None, 1, 0.125, True, seed, True, stream, ev)
        out.float().sum().backward()
        torch.cuda.synchronize()
        return out.detach(), x.grad, {k: v.grad for k, v in p.items()}

    a, b, c = run(11, 0.3), run(11, 0.3), run(12, 0.3)
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1]) and all(torch.equal(a[2][k], b[2][k]) for k in base)
    assert not torch.equal(a[0], c[0])
    nodrop = run(11, 0.0)
    assert not torch.equal(a[0], nodrop[0])
    s = torch.cuda.Stream()
    side = run(11, 0.3, s)
    assert torch.equal(a[0], side[0]) and torch.equal(a[1], side[1]) and all(torch.equal(a[2][k], side[2][k]) for k in base)
