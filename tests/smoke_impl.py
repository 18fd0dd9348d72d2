"""__graft_entry__.smoke(): one tiny invocation of each hot path on cuda:0, checked against the CPU oracle."""
import numpy as np
import torch


def run():
    assert torch.cuda.is_available(), "smoke() needs a GPU"
    dev = torch.device("cuda", 0)
    from dynamicemb import dynamicemb_extensions as ext
    from dynamicemb.scored_hashtable import LinearBucketTable, ScoreArg, ScoreSpec
    from dynamicemb.dynamicemb_extensions import ScorePolicy
    from oracle import dynamicemb as orc

    # --- DynamicEmb: insert -> fused probe+gather, vs oracle
    rng = np.random.default_rng(0)
    t = LinearBucketTable([128 * 16], [ScoreSpec("s", ScorePolicy.ASSIGN)], bucket_capacity=128, device=dev)
    o = orc.OracleTable([128 * 16], 128)
    keys = np.unique(rng.integers(0, 1 << 50, size=1500, dtype=np.int64))
    sc = np.ones(keys.size, dtype=np.int64)
    idx = t.insert(torch.from_numpy(keys).to(dev), torch.zeros(keys.size, dtype=torch.int64, device=dev), ScoreArg("s", torch.from_numpy(sc).to(dev)))
    oidx, _, _, _ = o.insert(keys, None, policy=1, score_in=sc)
    assert np.array_equal(idx.cpu().numpy(), oidx), "slot assignment differs from the oracle"
    assert np.array_equal(t.table_storage_.cpu().numpy(), o.storage), "table image differs from the oracle"
    D = 128
    values = torch.randn(t.capacity_, D, device=dev)
    ids = torch.from_numpy(keys[rng.integers(0, keys.size, size=4096)]).to(dev)
    out = ext.lookup_forward(t.table_storage_, t.table_bucket_offsets_, t.bucket_capacity_, values, D, ids)
    slot_of = dict(zip(keys.tolist(), oidx.tolist()))
    exp = orc.gather_rows(values.cpu().numpy(), D, np.array([slot_of[int(k)] for k in ids.cpu().numpy()]))
    assert np.array_equal(out.cpu().numpy(), exp), "fused lookup+gather differs from the oracle"
    print("smoke: dynamicemb ok")

    # --- HSTU attention fwd+bwd, vs oracle
    from tests import smoke_hstu
    smoke_hstu.run(dev)

    # --- HSTU layer glue: one fused layer (layer norm -> uvqk GEMM -> SiLU -> attention -> LN*u -> proj + residual) fwd + bwd vs fp32 torch
    import math
    import torch.nn.functional as F
    from hstu.fused_hstu_op import fused_hstu_op
    from oracle.hstu_attn import hstu_attention
    H, Dh, HID = 2, 64, 128
    cu = torch.tensor([0, 90, 200], dtype=torch.int32, device=dev)
    T, S, alpha = 200, 128, 1.0 / math.sqrt(Dh)
    g = torch.Generator(device=dev).manual_seed(5)
    mk = lambda *s, sc=1.0: (torch.randn(*s, device=dev, generator=g) * sc).to(torch.bfloat16)      # noqa: E731
    x, dout = mk(T, HID), mk(T, HID)
    P = dict(w_uvqk=mk(HID, 4 * H * Dh, sc=0.1), b_uvqk=mk(4 * H * Dh, sc=0.1), w_proj=mk(H * Dh, HID, sc=0.1), in_w=1 + mk(HID, sc=0.1),
             in_b=mk(HID, sc=0.1), out_w=1 + mk(H * Dh, sc=0.1), out_b=mk(H * Dh, sc=0.1))
    xq = x.clone().requires_grad_(True)
    pq = {k: v.clone().requires_grad_(True) for k, v in P.items()}
    out = fused_hstu_op(xq, cu, S, S, pq["w_uvqk"], pq["b_uvqk"], pq["w_proj"], H, Dh, Dh, 1e-5, 0.0, True, pq["in_w"], pq["in_b"], pq["out_w"], pq["out_b"],
                        None, None, None, 1, alpha)
    out.backward(dout)
    xr = x.float().requires_grad_(True)
    pr = {k: v.float().requires_grad_(True) for k, v in P.items()}
    n = F.layer_norm(xr, (HID,), pr["in_w"], pr["in_b"], 1e-5)
    u, v, q, k = F.silu(torch.addmm(pr["b_uvqk"], n, pr["w_uvqk"])).split(H * Dh, dim=-1)
    a = hstu_attention(q.reshape(T, H, Dh), k.reshape(T, H, Dh), v.reshape(T, H, Dh), cu, S, alpha, S)
    ref = torch.addmm(xr, F.layer_norm(a.reshape(T, H * Dh), (H * Dh,), pr["out_w"], pr["out_b"], 1e-5) * u, pr["w_proj"])
    ref.backward(dout.float())
    for name, got, want in [("out", out, ref), ("d_input", xq.grad, xr.grad), ("d_w_uvqk", pq["w_uvqk"].grad, pr["w_uvqk"].grad),
                            ("d_out_w", pq["out_w"].grad, pr["out_w"].grad)]:
        err, scale = (got.float() - want).abs().max().item(), want.abs().max().item()
        assert err <= 0.06 * scale + 1e-3, f"fused HSTU layer {name}: max abs err {err:.3e} against scale {scale:.3e}"
    print("smoke: hstu fused layer ok")
