"""sm_100 op layer of the HSTU attention: `hstu_varlen_fwd_100` / `hstu_varlen_bwd_100`, the two entry points the reference's
autograd function calls on Blackwell (third_party/FBGEMM/fbgemm_gpu/experimental/hstu/src/hstu_blackwell/hstu_ops_gpu.py:85-252,
:257-512), over the C ABI of librecsys_b200.so (include/hstu_b200.h).  No JIT, no compile cache: kernels are AOT sm_100a."""
import ctypes
from typing import Optional, Tuple

import torch

from dynamicemb import _native as N   # shared ctypes loader of librecsys_b200.so (raises if the library is missing)

P, I32, F32 = ctypes.c_void_p, ctypes.c_int, ctypes.c_float
N.lib.hstu_fwd_sm100.restype = I32
N.lib.hstu_fwd_sm100.argtypes = [P, P, P, P, P, P, P, I32, I32, I32, I32, I32, I32, I32, I32, I32, F32, P, P, ctypes.c_int64, P]
N.lib.hstu_bwd_sm100.restype = I32
N.lib.hstu_bwd_sm100.argtypes = [P, P, P, P, P, P, P, P, P, P, I32, I32, I32, I32, I32, I32, I32, I32, I32, F32, P, P, ctypes.c_int64, P]
N.lib.hstu_workspace_bytes.restype = ctypes.c_int64
N.lib.hstu_workspace_bytes.argtypes = []
_WS_BYTES = int(N.lib.hstu_workspace_bytes())
N.lib.sm100_probe_gemm.restype = I32
N.lib.sm100_probe_gemm.argtypes = [P, P, P, I32, P, P]

HSTU_ERR = {-1100: "invalid argument", -1101: "unsupported configuration (head_dim must be 64 or 128)", -1102: "workspace"}


def _check(rc, what):
    if rc == 0:
        return
    if rc in HSTU_ERR:
        raise ValueError(f"{what}: {HSTU_ERR[rc]}")
    if rc <= -2000:
        raise RuntimeError(f"{what}: TMA tensor map creation failed ({rc})")
    raise RuntimeError(f"{what}: CUDA error {-rc}")


def _supports_layout(t: torch.Tensor) -> bool:
    """unit last stride, token/head strides multiples of 8 elements, 16-byte aligned base (hstu_ops_gpu.py:77-82)"""
    return t.stride(2) == 1 and t.stride(0) % 8 == 0 and t.stride(1) % 8 == 0 and t.data_ptr() % 16 == 0


def _prep(t: torch.Tensor) -> torch.Tensor:
    return t if _supports_layout(t) else t.contiguous()


def _i32(t: Optional[torch.Tensor]) -> Optional[torch.Tensor]:
    if t is None:
        return None
    return t if t.dtype == torch.int32 and t.is_contiguous() else t.to(torch.int32).contiguous()


def hstu_varlen_fwd_100(q, k, v, cu_seqlens_q, cu_seqlens_k, max_seqlen_q, max_seqlen_k, num_contexts, num_targets, target_group_size,
                        window_size_left, window_size_right, alpha, rab=None, func=None, paged_kv=None, page_ids=None, page_indptrs=None,
                        scaling_seqlen: int = -1) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
    assert rab is None and func is None and paged_kv is None, "rab / arbitrary mask / paged KV are not on the training hot path (DESIGN.md)"
    assert q.dtype == torch.bfloat16 and k.dtype == q.dtype and v.dtype == q.dtype, "bf16 only"
    # self attention only (seqlen_q == seqlen_k, the training case).  Distinct cu_seqlens tensors are accepted on shape / max_seqlen
    # agreement; comparing their contents would be a host sync on every call.
    assert cu_seqlens_q.data_ptr() == cu_seqlens_k.data_ptr() or (cu_seqlens_q.shape == cu_seqlens_k.shape and max_seqlen_q == max_seqlen_k
                                                                  and k.shape[0] == q.shape[0]), "self attention (seqlen_q == seqlen_k) only"
    q, k, v = _prep(q), _prep(k), _prep(v)
    T, H, D = q.shape
    out = torch.empty(T, H, D, dtype=q.dtype, device=q.device)
    cu = _i32(cu_seqlens_q)
    B = cu.numel() - 1
    if scaling_seqlen is None or scaling_seqlen <= 0:
        scaling_seqlen = max_seqlen_q
    strides = (ctypes.c_int64 * 6)(q.stride(0), q.stride(1), k.stride(0), k.stride(1), v.stride(0), v.stride(1))
    nc, nt = _i32(num_contexts), _i32(num_targets)
    ws = torch.empty(_WS_BYTES, dtype=torch.uint8, device=q.device)      # tile counter of this launch (stream-ordered allocation)
    _check(N.lib.hstu_fwd_sm100(N.ptr(q), N.ptr(k), N.ptr(v), N.ptr(out), N.ptr(cu), N.ptr(nc), N.ptr(nt), B, H, D, T, int(max_seqlen_q),
                                int(scaling_seqlen), int(target_group_size), int(window_size_left), int(window_size_right), float(alpha),
                                ctypes.cast(strides, P), N.ptr(ws), ws.numel(), N.stream()), "hstu_fwd_sm100")
    return out, None


def hstu_varlen_bwd_100(do, q, k, v, cu_seqlens_q, cu_seqlens_k, max_seqlen_q, max_seqlen_k, dq, dk, dv, num_contexts, num_targets,
                        target_group_size, window_size_left, window_size_right, alpha, rab=None, has_drab=False, func=None,
                        deterministic=False, scaling_seqlen: int = -1):
    assert rab is None and not has_drab and func is None
    q, k, v, dout = _prep(q), _prep(k), _prep(v), _prep(do)          # parameter names are the reference's (hstu_blackwell/hstu_ops_gpu.py:257)
    T, H, D = q.shape
    dq = torch.empty(T, H, D, dtype=q.dtype, device=q.device) if dq is None else dq
    dk = torch.empty(T, H, D, dtype=q.dtype, device=q.device) if dk is None else dk
    dv = torch.empty(T, H, D, dtype=q.dtype, device=q.device) if dv is None else dv
    cu = _i32(cu_seqlens_q)
    B = cu.numel() - 1
    if scaling_seqlen is None or scaling_seqlen <= 0:
        scaling_seqlen = max_seqlen_q
    strides = (ctypes.c_int64 * 8)(q.stride(0), q.stride(1), k.stride(0), k.stride(1), v.stride(0), v.stride(1), dout.stride(0), dout.stride(1))
    nc, nt = _i32(num_contexts), _i32(num_targets)
    ws = torch.empty(_WS_BYTES, dtype=torch.uint8, device=q.device)      # tile counters of the two launches
    _check(N.lib.hstu_bwd_sm100(N.ptr(dout), N.ptr(q), N.ptr(k), N.ptr(v), N.ptr(dq), N.ptr(dk), N.ptr(dv), N.ptr(cu), N.ptr(nc), N.ptr(nt),
                                B, H, D, T, int(max_seqlen_q), int(scaling_seqlen), int(target_group_size), int(window_size_left),
                                int(window_size_right), float(alpha), ctypes.cast(strides, P), N.ptr(ws), ws.numel(), N.stream()), "hstu_bwd_sm100")
    return dq, dk, dv, None


def probe_gemm(A: torch.Tensor, Bm: torch.Tensor, variant: int, overrides=None) -> torch.Tensor:
    """development aid: one-CTA tcgen05 GEMM (csrc/sm100_probe.cu)"""
    C = torch.zeros(128, 128, dtype=torch.float32, device=A.device)
    ov = None
    if overrides is not None:
        ov = (ctypes.c_uint32 * 8)(*overrides)
    _check(N.lib.sm100_probe_gemm(N.ptr(A), N.ptr(Bm), N.ptr(C), variant, ctypes.cast(ov, P) if ov is not None else None, N.stream()), "probe")
    return C
