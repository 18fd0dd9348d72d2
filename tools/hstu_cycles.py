"""Cycle accounting of CTA 0 of the HSTU forward kernel (development aid)."""
import ctypes, math, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "recsys-examples_b200"))
from hstu import hstu_ops_gpu as ops
from dynamicemb import _native as N
dev = torch.device("cuda", 0)
B, S, H, D = 8, 4096, 8, 128
T = B * S
q, k, v = (torch.randn(T, H, D, device=dev, dtype=torch.bfloat16) for _ in range(3))
cu = torch.arange(0, T + 1, S, dtype=torch.int32, device=dev)
a = 1 / math.sqrt(D)
for _ in range(2): ops.hstu_varlen_fwd_100(q, k, v, cu, cu, S, S, None, None, 1, -1, 0, a)
torch.cuda.synchronize()
N.lib.hstu_set_debug_buffer.argtypes = [ctypes.c_void_p]
e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
for Bf in (8, 32):
    Tf = Bf * S
    qf, kf, vf = (torch.randn(Tf, H, D, device=dev, dtype=torch.bfloat16) for _ in range(3))
    cuf = torch.arange(0, Tf + 1, S, dtype=torch.int32, device=dev)
    for _ in range(2): ops.hstu_varlen_fwd_100(qf, kf, vf, cuf, cuf, S, S, None, None, 1, -1, 0, a)
    torch.cuda.synchronize()
    e0.record(); ops.hstu_varlen_fwd_100(qf, kf, vf, cuf, cuf, S, S, None, None, 1, -1, 0, a); e1.record(); torch.cuda.synchronize()
    t_prod = e0.elapsed_time(e1)
    dbgd = torch.zeros(128, dtype=torch.int32, device=dev)          # device buffer: every CTA adds its counters (units of 16 cycles)
    N.lib.hstu_set_debug_buffer(ctypes.c_void_p(dbgd.data_ptr()))
    e0.record(); ops.hstu_varlen_fwd_100(qf, kf, vf, cuf, cuf, S, S, None, None, 1, -1, 0, a); e1.record(); torch.cuda.synchronize()
    N.lib.hstu_set_debug_buffer(None)
    d = [16 * x for x in dbgd.tolist()]
    iters, ctas = d[8] // 16, d[9] // 16
    print(f"forward B={Bf}: production kernel {t_prod:.3f} ms, accounting build {e0.elapsed_time(e1):.3f} ms; {ctas} CTAs, {iters} iterations ({iters / 148:.0f} per SM)")
    print(f"  CTA lifetime {d[10] / ctas:9.0f} cyc avg = prologue {d[11] / ctas:.0f} + loop {(d[10] - d[11] - d[12]) / ctas:.0f} + epilogue {d[12] / ctas:.0f};"
          f"  sum of lifetimes / 148 SMs = {d[10] / 148 / 1e6:.3f} Mcyc -> {d[10] / 148 / t_prod / 1e3:.0f} MHz if back to back")
    names = {40: "mma: wait k_full", 42: "mma: wait v_full", 43: "mma: wait p_full", 48: "silu: wait s_full", 50: "silu: ld + math + st", 52: "silu: fence + arrive"}
    for k_, nm in names.items():
        print(f"  {nm:34s} {d[k_] / iters:8.0f} cyc / iteration")
    for k_, nm in {41: "mma: wait q_full", 53: "silu: between iterations (sum)", 55: "silu: tile fetch + setup", 56: "io: wait o_full", 57: "io: O read-out + stores"}.items():
        print(f"  {nm:34s} {d[k_] / ctas:8.0f} cyc / tile")
    print(f"  per tile: {iters / ctas:.1f} iterations x {(d[48] + d[50] + d[52]) / iters:.0f} = {(d[48] + d[50] + d[52]) / ctas:.0f} cyc in the SiLU loop, tile period {d[10] / ctas:.0f}")
    del qf, kf, vf
# ---- backward: dKV kernel slots 64.., dQ kernel slots 96.. (every CTA adds its counters, units of 16 cycles)
do = torch.randn_like(q)
for Bf in (8, 32):
    Tf = Bf * S
    qf, kf, vf, dof = (torch.randn(Tf, H, D, device=dev, dtype=torch.bfloat16) for _ in range(4))
    cuf = torch.arange(0, Tf + 1, S, dtype=torch.int32, device=dev)
    for _ in range(2): ops.hstu_varlen_bwd_100(dof, qf, kf, vf, cuf, cuf, S, S, None, None, None, None, None, 1, -1, 0, a)
    torch.cuda.synchronize()
    e0.record(); ops.hstu_varlen_bwd_100(dof, qf, kf, vf, cuf, cuf, S, S, None, None, None, None, None, 1, -1, 0, a); e1.record(); torch.cuda.synchronize()
    t_prod = e0.elapsed_time(e1)
    dbgd = torch.zeros(128, dtype=torch.int32, device=dev)
    N.lib.hstu_set_debug_buffer(ctypes.c_void_p(dbgd.data_ptr()))
    ops.hstu_varlen_bwd_100(dof, qf, kf, vf, cuf, cuf, S, S, None, None, None, None, None, 1, -1, 0, a); torch.cuda.synchronize()
    N.lib.hstu_set_debug_buffer(None)
    d = dbgd.tolist()
    print(f"backward B={Bf}: {t_prod:.3f} ms (dKV + dQ kernels)")
    for base, nm in ((64, "dKV"), (96, "dQ")):
        n, tiles = max(d[base + 15], 1), max(d[base + 14], 1)
        print(f" {nm} kernel: {tiles} stationary tiles, {n} iterations of 128x64 ({n / 148:.0f} per SM); MMA-thread lifetime {16 * d[base + 3] / n:.0f} cyc / iteration")
        for off, what in ((0, "mma: wait y_full"), (1, "mma: wait s_empty"), (2, "mma: wait operand tile"), (8, "silu: wait pd_empty"), (9, "silu: wait s_full"),
                          (10, "silu: ld + math + store"), (11, "silu: fence + arrive")):
            print(f"   {what:28s} {16 * d[base + off] / n:8.0f} cyc / iteration")
    del qf, kf, vf, dof
