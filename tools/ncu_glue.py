"""ncu target for the HSTU layer glue kernels: every kernel of csrc/hstu_glue.cu once at the HSTU-large shape (T = 32 x 4096, hidden 1024,
bf16) between cudaProfilerStart / Stop, after one un-profiled warm-up launch each.
    ncu --set full --clock-control none --import-source on --profile-from-start off -o gpurun_out/glue python tools/ncu_glue.py
    python tools/ncu_summary.py gpurun_out/glue.ncu-rep profiles/r02_hstu_glue_ncu_full_summary.csv
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "recsys-examples_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)


def main():
    from hstu import layer_ops as L
    dev = torch.device("cuda", 0)
    T, D, W, bf = 32 * 4096, 1024, 4096, torch.bfloat16
    x, dy, res = (torch.randn(T, D, device=dev, dtype=bf) for _ in range(3))
    w, b = torch.ones(D, device=dev, dtype=bf), torch.zeros(D, device=dev, dtype=bf)
    pre, act = (torch.randn(T, W, device=dev, dtype=bf) for _ in range(2))
    u = act[:, :D]
    du, dv, dq, dk = (torch.randn(T, D, device=dev, dtype=bf) for _ in range(4))

    def everything():
        y, mean, rstd, _, _ = L.weighted_layer_norm_fwd(x, w, b, 1e-5)
        L.weighted_layer_norm_bwd(dy, x, w, b, mean, rstd, True, 1e-5, 0, 0, res)
        L.silu_fwd(pre)
        L.silu_bwd_segments([du, dv, dq, dk], pre)
        y2, m2, r2, _, _, seed = L.layer_norm_mul_dropout_fwd(x, u, w, b, 1e-5, 0.0, True)
        L.layer_norm_mul_dropout_bwd(dy, x, u, w, b, m2, r2, 0, 0, 1e-5, True, 0.0, seed, False, True)
        L.layer_norm_mul_dropout_fwd(x, u, w, b, 1e-5, 0.2, True, False, 7)
    everything()
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()
    everything()
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()


if __name__ == "__main__":
    main()
