import math
import numpy as np
import torch


def run(dev):
    from hstu import hstu_attn_varlen_func
    from oracle import hstu_attn as orc
    lens, H, D = [200, 77, 256], 2, 128
    T = sum(lens)
    g = torch.Generator().manual_seed(0)
    q, k, v, dout = (torch.randn(T, H, D, generator=g).to(torch.bfloat16).to(dev) for _ in range(4))
    cu = torch.tensor(np.concatenate([[0], np.cumsum(lens)]), dtype=torch.int32, device=dev)
    qq, kk, vv = (t.clone().requires_grad_(True) for t in (q, k, v))
    out = hstu_attn_varlen_func(qq, kk, vv, cu, cu, None, None, 256, 256, -1, None, None, 1, (-1, 0), 1 / math.sqrt(D))
    out.backward(dout)
    ref = orc.fwd_bwd(q.float().cpu(), k.float().cpu(), v.float().cpu(), dout.float().cpu(), cu.cpu(), 256, 1 / math.sqrt(D))
    low = orc.fwd_bwd(q.cpu(), k.cpu(), v.cpu(), dout.cpu(), cu.cpu(), 256, 1 / math.sqrt(D), upcast=False)
    for nm, got, r32, r16, mult in zip(("out", "dq", "dk", "dv"), (out, qq.grad, kk.grad, vv.grad), ref, low, (2, 5, 5, 5)):
        err = (got.float().cpu() - r32).abs().max().item()
        base = (r16.float() - r32).abs().max().item()
        assert err <= mult * base + 1e-6, f"hstu {nm}: {err} > {mult} x {base}"
    print("smoke: hstu ok")
