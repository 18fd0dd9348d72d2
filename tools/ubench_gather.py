"""Micro-benchmark: random 512-B row gather / 1-KiB row read-modify-write as a function of the value-table size (TLB reach study).

  python tools/ubench_gather.py [--sizes 4,32,128] [--out gpurun_out/ubench_gather.json]

For each table size (GiB of [rows, 256] fp32 = 1 KiB value rows) times, with CUDA events:
  gather   : demb_gather_forward (cp.async.bulk staged, fp32 out) of 2^20 ids -> 2^20 x 512 B out
             row sets: unique-random, the same rows sorted by address, Zipf(1.05) duplicates (394 K unique), Zipf + sorted unique
  backward : demb_backward (sort + tiles/windows/spans, Adagrad) on the same id sets
Reports GB/s on the algorithmic bytes and the fraction of MEASURED_PEAKS.json's HBM copy bandwidth.
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "recsys-examples_b200"))


def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(True), torch.cuda.Event(True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sizes", default="4,32,128")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "ubench_gather.json"))
    args = ap.parse_args()
    from dynamicemb import dynamicemb_extensions as ext
    dev = torch.device("cuda", 0)
    peak = 6586.1
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        peak = json.load(open(p))["hbm_gbs"]
    D, V, n = 128, 256, 1 << 20
    res = {"peak_GBps": peak, "n_ids": n, "cases": []}
    g = torch.Generator(device=dev).manual_seed(1)
    for gib in [int(x) for x in args.sizes.split(",")]:
        rows_total = gib * (1 << 30) // (V * 4)
        values = torch.empty(rows_total, V, dtype=torch.float32, device=dev)
        values[: 1 << 20].zero_()
        uniq = torch.randint(0, rows_total, (n,), device=dev, generator=g, dtype=torch.int64)
        # Zipf-like duplicates: power law over a 1e9 key space hashed onto the rows
        x = torch.rand(n, device=dev, dtype=torch.float64, generator=g)
        a = 1.0 - 1.05
        keys = torch.pow(x * (1e9 ** a - 1.0) + 1.0, 1.0 / a).to(torch.int64)
        uk, inv = torch.unique(keys, return_inverse=True)
        nu = uk.numel()
        urows = (uk * 2654435761 % rows_total).to(torch.int64)
        urows_sorted, order = torch.sort(urows)
        inv_sorted = torch.empty_like(order); inv_sorted[order] = torch.arange(nu, device=dev)
        inv_of_sorted = inv_sorted[inv]
        ident = torch.arange(n, device=dev, dtype=torch.int64)
        grad = torch.randn(n, D, device=dev)
        cases = {
            "gather_unique_random": (lambda: ext.gather_forward(values, D, uniq, None, n), n * (8 + 512 + 512)),
            "gather_unique_sorted": (lambda r=torch.sort(uniq).values: ext.gather_forward(values, D, r, None, n), n * (8 + 512 + 512)),
            "gather_zipf": (lambda: ext.gather_forward(values, D, urows, inv, n), n * (8 + 512) + nu * (8 + 512)),
            "gather_zipf_unique_sorted": (lambda: ext.gather_forward(values, D, urows_sorted, inv_of_sorted, n), n * (8 + 512) + nu * (8 + 512)),
            "backward_zipf": (lambda: ext.backward(values, D, inv, nu, urows, grad, opt_type=3, lr=0.01), n * (512 + 8) + nu * (2 * 1024 + 8)),
            "backward_zipf_unique_sorted": (lambda: ext.backward(values, D, inv_of_sorted, nu, urows_sorted, grad, opt_type=3, lr=0.01), n * (512 + 8) + nu * (2 * 1024 + 8)),
            "backward_unique_random": (lambda: ext.backward(values, D, ident, n, uniq, grad, opt_type=3, lr=0.01), n * (512 + 8) + n * (2 * 1024 + 8)),
        }
        for name, (fn, nbytes) in cases.items():
            ms = timeit(fn)
            rec = {"table_GiB": gib, "case": name, "ms": ms, "GBps": nbytes / ms / 1e6, "frac": nbytes / ms / 1e6 / peak, "n_unique": nu}
            res["cases"].append(rec)
            print(json.dumps(rec), flush=True)
        del values
        torch.cuda.empty_cache()
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    json.dump(res, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
