"""Micro-benchmark of the HSTU layer glue kernels (csrc/hstu_glue.cu) at the HSTU-large shape: T = 32 x 4096 tokens, hidden 1024, bf16.

Every kernel is HBM-bound; `frac` = algorithmic bytes / time / the measured copy bandwidth (MEASURED_PEAKS.json `hbm_gbs`, else 6586 GB/s).
The same op through eager torch (F.layer_norm / F.silu autograd, what tools/e2e_harness.py's DenseStack runs) is timed beside it.
CUDA events, 20 iterations after 5 warm-ups; inputs are 268 MB..1 GB each, far larger than the 126 MB L2.
    python tools/bench_hstu_glue.py            -> one JSON line
"""
import json
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "recsys-examples_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)


def timeit(fn, iters=20, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def run(dev=None):
    from hstu import layer_ops as L
    dev = torch.device("cuda", 0) if dev is None else dev
    T, D, W = 32 * 4096, 1024, 4096
    peak = 6586.1
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:  # noqa: BLE001
        pass
    bf = torch.bfloat16
    x = torch.randn(T, D, device=dev, dtype=bf)
    dy = torch.randn(T, D, device=dev, dtype=bf)
    res = torch.randn(T, D, device=dev, dtype=bf)
    w = torch.ones(D, device=dev, dtype=bf)
    b = torch.zeros(D, device=dev, dtype=bf)
    pre = torch.randn(T, W, device=dev, dtype=bf)
    act = torch.randn(T, W, device=dev, dtype=bf)
    u = act[:, :D]
    du = torch.randn(T, D, device=dev, dtype=bf)
    dv, dq, dk = (torch.randn(T, D, device=dev, dtype=bf) for _ in range(3))
    rowb = T * D * 2                     # bytes of one [T, 1024] bf16 tensor
    out = {"shape": {"tokens": T, "hidden": D, "uvqk_width": W, "dtype": "bf16"}, "hbm_peak_gbs": peak, "kernels": {}}

    def rec(name, ms, nbytes, torch_ms=None):
        out["kernels"][name] = {"ms": round(ms, 4), "algorithmic_MB": round(nbytes / 1e6, 1), "GBps": round(nbytes / ms / 1e6, 1),
                                "frac": round(nbytes / ms / 1e6 / peak, 3), "torch_eager_ms": None if torch_ms is None else round(torch_ms, 4)}

    y, mean, rstd, _, _ = L.weighted_layer_norm_fwd(x, w, b, 1e-5)
    rec("layer_norm_fwd", timeit(lambda: L.weighted_layer_norm_fwd(x, w, b, 1e-5)), 2 * rowb, timeit(lambda: F.layer_norm(x, (D,), w, b, 1e-5)))
    xg = x.clone().requires_grad_(True)
    wg, bg = w.clone().requires_grad_(True), b.clone().requires_grad_(True)

    def torch_ln_bwd():
        yy = F.layer_norm(xg, (D,), wg, bg, 1e-5)
        torch.autograd.grad(yy, (xg, wg, bg), dy)
    t_fb = timeit(torch_ln_bwd)
    t_f = timeit(lambda: F.layer_norm(xg, (D,), wg, bg, 1e-5))
    rec("layer_norm_bwd(+residual add, dw, db)", timeit(lambda: L.weighted_layer_norm_bwd(dy, x, w, b, mean, rstd, True, 1e-5, 0, 0, res)), 4 * rowb,
        t_fb - t_f)
    rec("silu_fwd [T,4096]", timeit(lambda: L.silu_fwd(pre)), 2 * T * W * 2, timeit(lambda: F.silu(pre)))
    pg = pre.clone().requires_grad_(True)

    def torch_silu_bwd():
        g = torch.cat([du, dv, dq, dk], dim=1)
        torch.autograd.grad(F.silu(pg), pg, g)
    rec("silu_bwd (4 gradient segments read in place)", timeit(lambda: L.silu_bwd_segments([du, dv, dq, dk], pre)), 3 * T * W * 2,
        timeit(torch_silu_bwd) - timeit(lambda: F.silu(pg)))
    rec("silu_bwd + uvqk bias gradient (column sums in the same kernel)", timeit(lambda: L.silu_bwd_segments([du, dv, dq, dk], pre, with_bias_grad=True)),
        3 * T * W * 2, timeit(lambda: L.silu_bwd_segments([du, dv, dq, dk], pre).sum(dim=0)))
    out["kernels"]["silu_bwd + uvqk bias gradient (column sums in the same kernel)"]["torch_eager_ms_is"] = "our silu_bwd followed by a separate dx.sum(0)"
    y2, m2, r2, _, _, seed = L.layer_norm_mul_dropout_fwd(x, u, w, b, 1e-5, 0.0, True)
    rec("ln_mul_dropout_fwd", timeit(lambda: L.layer_norm_mul_dropout_fwd(x, u, w, b, 1e-5, 0.0, True)), 3 * rowb,
        timeit(lambda: F.layer_norm(x, (D,), w, b, 1e-5) * u))
    rec("ln_mul_dropout_fwd (ratio 0.2)", timeit(lambda: L.layer_norm_mul_dropout_fwd(x, u, w, b, 1e-5, 0.2, True, False, 7)), 3 * rowb)
    ug = u.clone().requires_grad_(True)

    def torch_nmd_bwd():
        yy = F.layer_norm(xg, (D,), wg, bg, 1e-5) * ug
        torch.autograd.grad(yy, (xg, ug, wg, bg), dy)
    t_nf = timeit(lambda: F.layer_norm(xg, (D,), wg, bg, 1e-5) * ug)
    rec("ln_mul_dropout_bwd (dx, du, dw, db + recomputed y)", timeit(lambda: L.layer_norm_mul_dropout_bwd(dy, x, u, w, b, m2, r2, 0, 0, 1e-5, True, 0.0, seed,
                                                                                                      False, True)), 6 * rowb, timeit(torch_nmd_bwd) - t_nf)
    return out


if __name__ == "__main__":
    print(json.dumps(run()))
