"""Row-wise glue ops of the HSTU layer over the C ABI of librecsys_b200.so (include/hstu_b200.h, csrc/hstu_glue.cu).

Same names, argument order and return tuples as the reference's Triton op layer, so `fused_hstu_op.py` reads like the reference's:
    triton_weighted_layer_norm_fwd / _bwd        examples/hstu/ops/triton_ops/triton_layer_norm.py:313, :386
    triton_layer_norm_mul_dropout_fwd / _bwd     examples/hstu/ops/triton_ops/triton_norm_mul_dropout.py:361, :426   (concat_ux=False)
    triton_silu_fwd / _bwd                       examples/hstu/ops/triton_ops/triton_silu.py:91, :108
(the `triton_` prefix is kept as an alias of each function: these are AOT sm_100a kernels, nothing is JIT-compiled; BLOCK_D / num_warps in
the return tuples are launch details of the Triton kernels and are returned as 0).  No CPU fallback: without the library the import raises.
"""
import ctypes
from typing import Optional, Sequence, Tuple

import torch

from dynamicemb import _native as N

P, I32, I64, U64, F32 = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_uint64, ctypes.c_float
_SIGS = {
    "hstu_glue_workspace_bytes": (I64, [I32]),
    "hstu_layer_norm_fwd": (I32, [P, I64, P, P, P, I64, P, P, I64, I32, F32, I32, P]),
    "hstu_layer_norm_bwd": (I32, [P, I64, P, I64, P, P, P, P, I64, P, I64, P, P, P, I64, I64, I32, I32, P]),
    "hstu_ln_mul_dropout_fwd": (I32, [P, I64, P, I64, P, P, P, I64, P, P, I64, I32, F32, F32, U64, I32, I32, P]),
    "hstu_ln_mul_dropout_bwd": (I32, [P, I64, P, I64, P, I64, P, P, P, P, P, I64, P, I64, P, I64, P, P, P, I64, I64, I32, F32, U64, I32, I32, P]),
    "hstu_silu_fwd": (I32, [P, P, I64, I32, P]),
    "hstu_silu_bwd": (I32, [I32, P, P, P, P, P, I64, I32, P]),
    "hstu_silu_bwd_bias_workspace_bytes": (I64, [I32]),
    "hstu_silu_bwd_bias": (I32, [I32, P, P, P, P, P, P, P, I64, I64, I32, P]),
    "hstu_dropout_mask": (I32, [I64, I32, F32, U64, P, P]),
}
for _name, (_res, _args) in _SIGS.items():
    _f = getattr(N.lib, _name)
    _f.restype, _f.argtypes = _res, _args

_DTYPE = {torch.float32: 0, torch.float16: 1, torch.bfloat16: 2}
_ERR = {-1100: "invalid argument (pointers 16-byte aligned, strides / widths multiples of 8 elements)",
        -1101: "unsupported: the normalised width must be <= 8192", -1102: "workspace too small"}
MAX_NORM_DIM = 8192          # <= 1024: warp-per-row kernels (the HSTU-large case); up to 8192: CTA-per-row kernels


def _check(rc: int, what: str) -> None:
    if rc == 0:
        return
    if rc in _ERR:
        raise ValueError(f"{what}: {_ERR[rc]}")
    raise RuntimeError(f"{what}: CUDA error {-rc}")


def _rows(t: torch.Tensor) -> torch.Tensor:
    """2-D, unit inner stride, row stride a multiple of 8 elements, 16-byte aligned — else a contiguous copy (switch_to_contiguous_if_needed)."""
    assert t.dim() == 2
    if t.stride(1) != 1 or t.stride(0) % 8 != 0 or t.data_ptr() % 16 != 0:
        t = t.contiguous()
    return t


def _param(p: Optional[torch.Tensor], like: torch.Tensor) -> Optional[torch.Tensor]:
    if p is None:
        return None
    return p.detach().to(like.dtype).contiguous()


def _workspace(D: int, device) -> torch.Tensor:
    return N.workspace(int(N.lib.hstu_glue_workspace_bytes(D)), device)


def weighted_layer_norm_fwd(x: torch.Tensor, weight: Optional[torch.Tensor], bias: Optional[torch.Tensor], eps: float,
                            mean: Optional[torch.Tensor] = None, rstd: Optional[torch.Tensor] = None
                            ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, int, int]:
    """(y, mean, rstd, BLOCK_D, num_warps).  mean / rstd given = buffers to fill (the statistics are always recomputed)."""
    assert x.dim() == 2, f"x.dim() == {x.dim()}, expected 2"
    x = _rows(x)
    n, D = x.shape
    if weight is not None:
        assert bias is not None and weight.dim() == 1 and bias.dim() == 1 and weight.numel() == D and bias.numel() == D
    y = torch.empty(n, D, dtype=x.dtype, device=x.device)
    mean = torch.empty(n, dtype=torch.float32, device=x.device) if mean is None else mean
    rstd = torch.empty(n, dtype=torch.float32, device=x.device) if rstd is None else rstd
    w, b = _param(weight, x), _param(bias, x)
    if n == 0:
        return y, mean, rstd, 0, 0
    _check(N.launch("hstu_layer_norm_fwd", 1, N.lib.hstu_layer_norm_fwd, N.ptr(x), x.stride(0), N.ptr(w), N.ptr(b), N.ptr(y), y.stride(0), N.ptr(mean),
                    N.ptr(rstd), n, D, float(eps), _DTYPE[x.dtype], N.stream()), "layer_norm_fwd")
    return y, mean, rstd, 0, 0


def weighted_layer_norm_bwd(dy: torch.Tensor, x: torch.Tensor, weight: Optional[torch.Tensor], bias: Optional[torch.Tensor], mean: torch.Tensor,
                            rstd: torch.Tensor, learnable: bool, eps: float, BLOCK_D: int = 0, num_warps: int = 0,
                            dx_accumulate: Optional[torch.Tensor] = None, wait_event: Optional[torch.cuda.Event] = None
                            ) -> Tuple[torch.Tensor, Optional[torch.Tensor], Optional[torch.Tensor]]:
    """(dx, dweight, dbias): dx = LN backward of dy + dx_accumulate (the residual branch's gradient), weight / bias gradients in the same
    pass.  dweight / dbias have the weight's dtype (fp32 accumulation), None when not learnable."""
    if wait_event is not None:
        wait_event.wait(torch.cuda.current_stream())
    x, dy = _rows(x), _rows(dy.to(x.dtype))
    n, D = x.shape
    acc = _rows(dx_accumulate.to(x.dtype)) if dx_accumulate is not None else None
    dx = torch.empty(n, D, dtype=x.dtype, device=x.device)
    dwb = torch.empty(2, D, dtype=torch.float32, device=x.device) if learnable else None
    w = _param(weight, x) if learnable else None
    ws = _workspace(D, x.device)
    _check(N.launch("hstu_layer_norm_bwd", 2 if learnable else 1, N.lib.hstu_layer_norm_bwd, N.ptr(dy), dy.stride(0), N.ptr(x), x.stride(0), N.ptr(w),
                    N.ptr(mean), N.ptr(rstd), N.ptr(acc), acc.stride(0) if acc is not None else 0, N.ptr(dx), dx.stride(0),
                    N.ptr(dwb[0]) if learnable else None, N.ptr(dwb[1]) if learnable else None, N.ptr(ws), ws.numel(), n, D, _DTYPE[x.dtype],
                    N.stream()), "layer_norm_bwd")
    if not learnable:
        return dx, None, None
    return dx, dwb[0].to(weight.dtype), dwb[1].to(weight.dtype)


def _seed(seed: Optional[int]) -> int:
    if seed is None:
        seed = int(torch.randint(low=0, high=2 ** 62, size=(1,), dtype=torch.int64).item())
    return int(seed) & 0xFFFFFFFFFFFFFFFF


def _u2d(u: torch.Tensor) -> torch.Tensor:
    if u.dim() == 3:                       # (T, heads, dim) view of the uvqk buffer: heads are adjacent, so it is a 2-D row view
        u = u.reshape(u.size(0), -1) if u.stride(1) != u.size(2) or u.stride(2) != 1 else u.as_strided((u.size(0), u.size(1) * u.size(2)), (u.stride(0), 1))
    return _rows(u)


def layer_norm_mul_dropout_fwd(x: torch.Tensor, u: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor, eps: float, dropout_ratio: float,
                               training: bool, concat_ux: bool = False, seed: Optional[int] = None
                               ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, int, int, int]:
    """(y, mean, rstd, BLOCK_D, num_warps, seed):  y = dropout(LN(x) * u).  `u` may be a strided row view (a slice of the uvqk buffer)."""
    if concat_ux:
        raise NotImplementedError("concat_ux=True is not used by the fused HSTU layer (fused_hstu_op.py:447) and is not built")
    assert x.dim() == 2
    x = _rows(x)
    u = _u2d(u if u.dtype == x.dtype else u.to(x.dtype))
    n, D = x.shape
    assert u.shape == (n, D) and weight.numel() == D and bias.numel() == D
    y = torch.empty(n, D, dtype=x.dtype, device=x.device)
    mean = torch.empty(n, dtype=torch.float32, device=x.device)
    rstd = torch.empty(n, dtype=torch.float32, device=x.device)
    if n == 0:
        return y, mean, rstd, 0, 0, 0
    seed = _seed(seed) if (training and dropout_ratio > 0.0) else (0 if seed is None else int(seed) & 0xFFFFFFFFFFFFFFFF)
    w, b = _param(weight, x), _param(bias, x)
    _check(N.launch("hstu_ln_mul_dropout_fwd", 1, N.lib.hstu_ln_mul_dropout_fwd, N.ptr(x), x.stride(0), N.ptr(u),
                    u.stride(0), N.ptr(w), N.ptr(b), N.ptr(y), y.stride(0), N.ptr(mean), N.ptr(rstd), n, D, float(eps), float(dropout_ratio), seed,
                    1 if training else 0, _DTYPE[x.dtype], N.stream()), "layer_norm_mul_dropout_fwd")
    return y, mean, rstd, 0, 0, seed


def layer_norm_mul_dropout_bwd(dy: torch.Tensor, x: torch.Tensor, u: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor, mean: torch.Tensor,
                               rstd: torch.Tensor, BLOCK_D: int = 0, num_warps: int = 0, eps: float = 1e-5, training: bool = True,
                               dropout_ratio: float = 0.0, seed: Optional[int] = None, concat_ux: bool = False, compute_y: bool = False,
                               wait_event: Optional[torch.cuda.Event] = None, du: Optional[torch.Tensor] = None
                               ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor, Optional[torch.Tensor]]:
    """(dx, du, dweight, dbias, y).  One pass over dy / x / u; y (the forward output) is recomputed on the way when compute_y.
    `du` given = the buffer to write (may be a strided slice of the uvqk gradient buffer)."""
    if concat_ux:
        raise NotImplementedError("concat_ux=True is not built")
    if wait_event is not None:
        wait_event.wait(torch.cuda.current_stream())
    x = _rows(x)
    dy, u = _rows(dy.to(x.dtype)), _u2d(u if u.dtype == x.dtype else u.to(x.dtype))
    n, D = x.shape
    dx = torch.empty(n, D, dtype=x.dtype, device=x.device)
    if du is None:
        du = torch.empty(n, D, dtype=x.dtype, device=x.device)
    assert du.shape == (n, D) and du.dtype == x.dtype and du.stride(1) == 1 and du.stride(0) % 8 == 0 and du.data_ptr() % 16 == 0
    y = torch.empty(n, D, dtype=x.dtype, device=x.device) if compute_y else None
    dwb = torch.empty(2, D, dtype=torch.float32, device=x.device)
    w, b = _param(weight, x), _param(bias, x)
    ws = _workspace(D, x.device)
    drop = training and dropout_ratio > 0.0
    assert not drop or seed is not None, "the seed the forward returned is needed to regenerate the dropout mask"
    _check(N.launch("hstu_ln_mul_dropout_bwd", 2, N.lib.hstu_ln_mul_dropout_bwd, N.ptr(dy), dy.stride(0), N.ptr(x), x.stride(0), N.ptr(u), u.stride(0),
                    N.ptr(w), N.ptr(b), N.ptr(mean), N.ptr(rstd), N.ptr(dx), dx.stride(0), N.ptr(du), du.stride(0), N.ptr(y),
                    y.stride(0) if y is not None else 0, N.ptr(dwb[0]), N.ptr(dwb[1]), N.ptr(ws), ws.numel(), n, D, float(dropout_ratio),
                    (int(seed) & 0xFFFFFFFFFFFFFFFF) if seed is not None else 0, 1 if training else 0, _DTYPE[x.dtype], N.stream()),
           "layer_norm_mul_dropout_bwd")
    return dx, du, dwb[0].to(weight.dtype), dwb[1].to(weight.dtype), y


def silu_fwd(input: torch.Tensor) -> torch.Tensor:  # noqa: A002
    x = input.contiguous()
    assert x.numel() % 8 == 0, "element count must be a multiple of 8"
    y = torch.empty_like(x)
    _check(N.launch("hstu_silu_fwd", 1, N.lib.hstu_silu_fwd, N.ptr(x), N.ptr(y), x.numel(), _DTYPE[x.dtype], N.stream()), "silu_fwd")
    return y.view(input.shape)


def silu_bwd_segments(grad_segments: Sequence[torch.Tensor], input: torch.Tensor, with_bias_grad: bool = False):  # noqa: A002
    """d(silu) for `input` [rows, W] whose output gradient arrives as 1..4 column segments [rows, w_i] (sum w_i = W), each a row view with
    its own base and stride — du / dv / dq / dk are read where the layer-norm and attention backward kernels left them (no torch.cat).
    with_bias_grad: also return dx.sum(0) (fp32 [W]) — the bias gradient of the GEMM in front of the SiLU — accumulated by the same kernel
    instead of a second pass over dx; returns (dx, dbias)."""
    x = input.contiguous()
    rows, W = x.shape
    segs = [_rows(g.reshape(rows, -1).to(x.dtype)) for g in grad_segments]
    assert 1 <= len(segs) <= 4 and sum(s.shape[1] for s in segs) == W
    dx = torch.empty_like(x)
    ptrs = (ctypes.c_void_p * len(segs))(*[s.data_ptr() for s in segs])
    strides = (ctypes.c_int64 * len(segs))(*[s.stride(0) for s in segs])
    widths = (ctypes.c_int32 * len(segs))(*[s.shape[1] for s in segs])
    if with_bias_grad:
        ws_bytes = int(N.lib.hstu_silu_bwd_bias_workspace_bytes(W))
        if ws_bytes > 0:
            dbias = torch.empty(W, dtype=torch.float32, device=x.device)
            ws = N.workspace(ws_bytes, x.device)
            _check(N.launch("hstu_silu_bwd_bias", 2, N.lib.hstu_silu_bwd_bias, len(segs), ctypes.cast(ptrs, P), ctypes.cast(strides, P),
                            ctypes.cast(widths, P), N.ptr(x), N.ptr(dx), N.ptr(dbias), N.ptr(ws), ws.numel(), rows, _DTYPE[x.dtype], N.stream()), "silu_bwd_bias")
            return dx, dbias
    _check(N.launch("hstu_silu_bwd", 1, N.lib.hstu_silu_bwd, len(segs), ctypes.cast(ptrs, P), ctypes.cast(strides, P), ctypes.cast(widths, P), N.ptr(x),
                    N.ptr(dx), rows, _DTYPE[x.dtype], N.stream()), "silu_bwd")
    if with_bias_grad:                                      # width that cannot be tiled onto the launch: separate sum
        return dx, dx.sum(dim=0, dtype=torch.float32)
    return dx


def silu_bwd(grad_output: torch.Tensor, input: torch.Tensor) -> torch.Tensor:  # noqa: A002
    x2 = input.reshape(-1, input.shape[-1])
    return silu_bwd_segments([grad_output.reshape(x2.shape)], x2).view(input.shape)


def dropout_mask(rows: int, D: int, dropout_ratio: float, seed: int, device) -> torch.Tensor:
    """test aid: the boolean keep mask the ln_mul_dropout kernels apply for (ratio, seed)."""
    keep = torch.empty(rows, D, dtype=torch.uint8, device=device)
    _check(N.lib.hstu_dropout_mask(rows, D, float(dropout_ratio), int(seed) & 0xFFFFFFFFFFFFFFFF, N.ptr(keep), N.stream()), "dropout_mask")
    return keep.bool()


# the reference's names
triton_weighted_layer_norm_fwd = weighted_layer_norm_fwd
triton_weighted_layer_norm_bwd = weighted_layer_norm_bwd
triton_layer_norm_mul_dropout_fwd = layer_norm_mul_dropout_fwd
triton_layer_norm_mul_dropout_bwd = layer_norm_mul_dropout_bwd
triton_silu_fwd = silu_fwd
triton_silu_bwd = silu_bwd
