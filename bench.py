#!/usr/bin/env python
"""bench.py — the driver's measurement contract for the two hot paths (DESIGN.md §3 Measurement).

  python bench.py --gpus N --steps K --warmup W            our arm (torchrun launches N>1 ranks)
  python bench.py --impl reference ...                      the reference's CPU path of the same workload (oracle port), rank 0 only

Workload (BASELINE.json configs[1], SURVEY §8(d) cfg 2): DynamicEmb hash-table embedding, key space 1e9, D=128 fp32 rows, fused Adagrad
in backward, ids from the reference's power-law generator (corelib/dynamicemb/benchmark/dataset_generator.py:4-22, alpha 1.05), 2^20 ids
per step per GPU, sequence output.  Table per GPU: 128 Mi rows x 1 KiB ([emb | Adagrad state]) = 128 GiB + 2.1 GiB key map, STEP scores,
prefilled to EVICTION STEADY STATE (every bucket full: ~4e8 power-law draws of the stream, then score-0 filler keys from outside the
key space until the table is full — each new key of a timed step evicts a min-score slot of its bucket).  A "step" =
BatchedDynamicEmbeddingTablesV2.forward (dedup, probe, insert+evict+init of new ids, gather) + backward (gradient reduce + Adagrad row
update).  Extra keys: the same step WITHOUT eviction (key space 64 Mi, load < 0.5), the eager (non-graph) step, the fused eval lookup.
N>1: row-wise sharded (hash_roundrobin), each rank feeds its own 2^20-id batch (weak scaling); + SURVEY cfg 5 batch sweep (key space
1e10).  Second hot path (configs[2]): HSTU attention fwd+bwd, B=32 x S=4096 causal, H=8, D=128 bf16, under "hstu_attn" with the
reference's own sm100 kernels timed beside it when staged (baseline/_ref).  CPU legs: oracle port of this workload, cfg 1
(4 x nn.EmbeddingBag 1M x 64), eager-PyTorch HSTU.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "recsys-examples_b200"))

D = 128
KEY_SPACE = 1_000_000_000
ALPHA = 1.05
METRIC = "128-d embedding lookups/sec (DynamicEmb lookup + fused Adagrad update, Zipf 1.05)"
UNIT = "lookups/s"
FILLER_BASE = 1 << 40          # filler keys live outside every key space used here: never looked up, score 0 => evicted first


def power_law_ids(n, gen, device, key_space=KEY_SPACE):
    """PowerLaw(1, key_space, alpha) — restates dataset_generator.py:4-22 (inverse-CDF on float64)."""
    x = torch.rand(n, device=device, dtype=torch.float64, generator=gen)
    g = 1.0 - ALPHA
    y = torch.pow(x * (key_space ** g - 1.0) + 1.0, 1.0 / g)
    y = torch.clamp(y, max=key_space - 1)
    return y.to(torch.int64)


def _lsr(x, n):
    return (x >> n) & ((1 << (64 - n)) - 1)


def fmix64_torch(x):
    """murmur3 fmix64 on int64 tensors (two's-complement wraparound = uint64 arithmetic); low 63 bits — enough for `% W`, W a power of two."""
    k = x.clone()
    k = k ^ _lsr(k, 33)
    k = k * (-49064778989728563)          # 0xff51afd7ed558ccd as int64
    k = k ^ _lsr(k, 33)
    k = k * (-4265267296055464877)        # 0xc4ceb9fe1a85ec53 as int64
    k = k ^ _lsr(k, 33)
    return k & 0x7FFFFFFFFFFFFFFF


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        j = json.load(open(p))
        return j["hbm_gbs"], j["bf16_tflops"], "measured (MEASURED_PEAKS.json, burst)"
    return 6650.0, 1590.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md clocks line)."""

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        rows = [r for r in self.rows if len(r) >= 6 and r[0].isdigit()]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm = sorted(int(r[0]) for r in rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[2 + i].lower().startswith("active") for r in rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": int(rows[0][1]), "reasons": reasons, "samples": len(rows)}


def workload_config(args, n_ids):
    par = (f"row-wise x{args.gpus} (hash_roundrobin, index dedup, exchange of ids and rows over NVLink; KJT = 1 feature x 4096 samples x "
           f"{n_ids // 4096} ids per rank)") if args.gpus > 1 else "single GPU"
    return {"workload": "dynamicemb_keyspace1e9_d128_fp32_adagrad_zipf1.05_seq", "ids_per_step_per_gpu": n_ids, "table_rows_per_gpu": args.capacity,
            "bucket_capacity": 128, "score_strategy": "STEP", "table_state": "eviction steady state (all buckets full; prefill = power-law stream + score-0 filler)",
            "parallelism": par,
            "l2": "value table (128 GiB) and the per-step id / gradient / output streams (0.5 GB each) exceed the 126 MB L2; a fresh id batch every step"}


# ---------------------------------------------------------------------------------------------------------------- CPU legs (reported baselines)
CPU_TABLE_ROWS = 1 << 22


def cpu_reference_step_factory(seed):
    """The reference's CPU path of this workload, as a port: TorchRec CPU EmbeddingCollection is an nn.Embedding-style dense gather per
    table; dynamic keys need a key->row map first, done by the oracle's restatement of the reference hash table
    (oracle/dynamicemb_oracle.c, sequential C), on a 4 Mi-row table filled to eviction steady state like the GPU arm's.  Gather /
    per-key gradient reduce / Adagrad run in multi-threaded torch CPU ops."""
    from oracle.dynamicemb import OracleTable
    torch.set_num_threads(min(os.cpu_count(), 32))      # more threads only add oversubscription on these small tensors
    cap = CPU_TABLE_ROWS
    tab = OracleTable([cap], 128)
    values = torch.zeros(cap, 2 * D)
    values.fill_(0.0)                                                 # touch every page now, not inside the timed steps
    filler = FILLER_BASE + np.arange(int(cap * 1.3), dtype=np.int64)
    tab.insert(filler, None, policy=1, score_in=np.zeros(filler.size, dtype=np.int64))     # full table: every new key evicts
    gen = torch.Generator().manual_seed(seed)
    state = {"step": 1}

    def step(n_ids):
        ids = power_law_ids(n_ids, gen, "cpu")
        uk, inv = torch.unique(ids, return_inverse=True)
        sc = np.full(uk.numel(), state["step"], dtype=np.int64)
        state["step"] += 1
        _, found, slots = tab.lookup(uk.numpy(), None, policy=1, score_in=sc)
        miss = ~found
        if miss.any():
            new_slots, _, _, _ = tab.insert(uk.numpy()[miss], None, policy=1, score_in=sc[miss])
            slots[miss] = new_slots
            ns = torch.from_numpy(new_slots[new_slots >= 0])
            values[ns, :D] = torch.empty(ns.numel(), D).uniform_(-0.01, 0.01)
            values[ns, D:] = 0
        rows = torch.from_numpy(slots).clamp(min=0)
        out = values[rows[inv], :D]                                   # forward gather
        grad = torch.ones_like(out)
        ug = torch.zeros(uk.numel(), D).index_add_(0, inv, grad)      # reduce per unique key
        st = values[rows, D:] + ug * ug                               # Adagrad
        values[rows, D:] = st
        values[rows, :D] -= 0.01 * ug / (st.sqrt() + 1e-8)
        return n_ids

    return step


def _cpu_pick_step_size(step, n_target, steps, budget_s):
    """Largest power-of-two ids/step <= n_target for which `steps` steps fit in budget_s (calibrated on two 2^17-id steps)."""
    step(1 << 17)
    t0 = time.perf_counter(); step(1 << 17); per_id = (time.perf_counter() - t0) / (1 << 17)
    n = n_target
    while n > (1 << 17) and per_id * n * steps > budget_s:
        n >>= 1
    return n


def cpu_baseline(budget_s=15.0, n_target=1 << 20):
    step = cpu_reference_step_factory(7)
    n_ids = _cpu_pick_step_size(step, n_target, 4, budget_s)
    t0, done = time.perf_counter(), 0
    while time.perf_counter() - t0 < budget_s or done == 0:
        done += step(n_ids)
    dt = time.perf_counter() - t0
    return {"value": done / dt, "unit": UNIT, "cores": min(os.cpu_count(), 32), "kind": "port",
            "sample": f"{done // n_ids} steps x {n_ids} ids of the same power-law stream on a {CPU_TABLE_ROWS >> 20} Mi-row table at eviction steady state, {dt:.1f} s wall"}


def cpu_cfg1_embeddingbag(budget_s=6.0):
    """BASELINE.json configs[0] / SURVEY §8(d) cfg 1: unsharded TorchRec EmbeddingBagCollection on CPU = one nn.EmbeddingBag(mode='sum') per
    table applied to its feature's slice and concatenated: 4 tables x 1M x 64 fp32, B=8192, hotness 10, power-law ids (alpha 1.05)."""
    torch.manual_seed(0)
    torch.set_num_threads(os.cpu_count())
    T, N, Dm, B, H = 4, 1_000_000, 64, 8192, 10
    bags = [torch.nn.EmbeddingBag(N, Dm, mode="sum", _weight=torch.rand(N, Dm) - 0.5) for _ in range(T)]
    gen = torch.Generator().manual_seed(0)
    offsets = torch.arange(0, B * H + 1, H)[:-1]

    def step():
        with torch.no_grad():
            ids = [power_law_ids(B * H, gen, "cpu", key_space=N) for _ in range(T)]
            return torch.cat([bag(i, offsets) for bag, i in zip(bags, ids)], dim=1)

    step()
    best, t_end = 1e9, time.perf_counter() + budget_s
    runs = 0
    while time.perf_counter() < t_end or runs < 5:
        t0 = time.perf_counter(); step(); best = min(best, time.perf_counter() - t0); runs += 1
    return {"value": T * B * H / best, "unit": "ids/s (forward lookup + sum pooling)", "cores": os.cpu_count(), "kind": "port",
            "sample": f"4 x nn.EmbeddingBag(1e6, 64, 'sum'), B=8192 x hotness 10 = {T * B * H} ids/step, best of {runs} ({best * 1e3:.2f} ms); "
                      "stands in for TorchRec's CPU EmbeddingBagCollection (torchrec is not installed: parity with it unpinned)"}


def cpu_eager_hstu(budget_s=10.0):
    """The reference's eager-PyTorch HSTU path (examples/hstu/ops/pt_ops/pt_hstu_attention.py:150-196, restated in oracle/hstu_attn.py) on
    the host cores: fp32, B=2 x S=1024, H=8, D=128 causal, forward + backward; TFLOP/s by the same flop accounting as the GPU numbers."""
    from oracle import hstu_attn as orc
    torch.set_num_threads(os.cpu_count())
    B, S, H, Dh = 2, 1024, 8, 128
    T = B * S
    g = torch.Generator().manual_seed(0)
    q, k, v, do = (torch.randn(T, H, Dh, generator=g) for _ in range(4))
    cu = torch.arange(0, T + 1, S, dtype=torch.int32)
    a = 1 / math.sqrt(Dh)
    fl = orc.fwd_flops([S] * B, H, Dh)
    best_f, best_fb, t_end, runs = 1e9, 1e9, time.perf_counter() + budget_s, 0
    while runs < 2 or time.perf_counter() < t_end:
        t0 = time.perf_counter()
        with torch.no_grad():
            orc.hstu_attention(q, k, v, cu, S, a)
        best_f = min(best_f, time.perf_counter() - t0)
        t0 = time.perf_counter(); orc.fwd_bwd(q, k, v, do, cu, S, a); best_fb = min(best_fb, time.perf_counter() - t0)
        runs += 1
    return {"fwd_ms": best_f * 1e3, "fwd_bwd_ms": best_fb * 1e3, "fwd_tflops": fl / best_f / 1e12, "fwd_bwd_tflops": 3.5 * fl / best_fb / 1e12,
            "cores": os.cpu_count(), "kind": "port",
            "sample": f"B={B} x S={S}, H={H}, D={Dh} fp32 causal, best of {runs}; the full B=32 x S=4096 fp32 score tensor is 17 GB - scale linearly in B and quadratically in S"}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    step = cpu_reference_step_factory(7)
    n_ids = _cpu_pick_step_size(step, args.ids, args.steps + args.warmup, 150.0)      # the full 2^20-id step when the host is fast enough
    for _ in range(args.warmup):
        step(n_ids)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step(n_ids)
    dt = time.perf_counter() - t0
    v = n_ids * args.steps / dt
    cores = min(os.cpu_count(), 32)
    sample = (f"each step = {n_ids} ids of the same power-law stream ({'the full step' if n_ids == args.ids else 'a bounded sample of the ' + str(args.ids) + '-id step'}); the key->row map is a {CPU_TABLE_ROWS >> 20} Mi-row table at eviction "
              f"steady state instead of {args.capacity >> 20} Mi rows (host memory bound of the sample); value rows live in host DRAM")
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": workload_config(args, args.ids),
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------------------------- HSTU side bench
def _p10(f, iters, warm=3):
    for _ in range(warm):
        f()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record(); f(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 10]            # P10 like hstu_attn_kernel_benchmark.py


def bench_hstu(dev, tflops_peak, iters=10, with_reference=True):
    try:
        from hstu import hstu_ops_gpu as ops
        B, S, H, Dh = 32, 4096, 8, 128
        T = B * S
        buf = torch.randn(T, 4 * H * Dh, device=dev, dtype=torch.bfloat16)
        _, v, q, k = (t.view(T, H, Dh) for t in buf.split(H * Dh, dim=-1))
        dout = torch.randn(T, H, Dh, device=dev, dtype=torch.bfloat16)
        cu = torch.arange(0, T + 1, S, dtype=torch.int32, device=dev)
        alpha = 1 / math.sqrt(Dh)
        fwd = lambda: ops.hstu_varlen_fwd_100(q, k, v, cu, cu, S, S, None, None, 1, -1, 0, alpha)
        bwd = lambda: ops.hstu_varlen_bwd_100(dout, q, k, v, cu, cu, S, S, None, None, None, None, None, 1, -1, 0, alpha)
        res = {}
        flops_fwd = 2.0 * H * Dh * S * S * B          # causal: 4*H*D*S^2/2 (examples/commons/utils/perf.py:729-739)
        for name, f, fl in (("fwd", fwd, flops_fwd), ("bwd", bwd, 2.5 * flops_fwd)):
            ms = _p10(f, iters)
            res[name] = {"ms": ms, "tflops": fl / ms / 1e9, "frac_of_bf16_peak": fl / ms / 1e9 / tflops_peak}
        res["samples_per_s_attn_only_8_layers"] = B / ((res["fwd"]["ms"] + res["bwd"]["ms"]) * 8 / 1e3)
        # SURVEY 8(d) config 3, profile (ii): jagged lengths, Zipf alpha 1.2 in [1, 4096] (examples/commons/datasets/hstu_batch.py:156-170, numpy
        # fallback branch), seed 1234; and the uniform profile with 256 targets per sequence (group size 1)
        lens = np.clip(np.random.default_rng(1234).zipf(1.2, size=B), 1, S).astype(np.int64)
        cuj = torch.tensor(np.concatenate([[0], np.cumsum(lens)]), dtype=torch.int32, device=dev)
        Tj = int(lens.sum())
        fl_j = float(2.0 * H * Dh * (lens.astype(np.float64) ** 2).sum())
        nt = torch.full((B,), 256, dtype=torch.int32, device=dev)
        variants = {"jagged_zipf1.2": (lambda: ops.hstu_varlen_fwd_100(q[:Tj], k[:Tj], v[:Tj], cuj, cuj, S, S, None, None, 1, -1, 0, alpha),
                                       lambda: ops.hstu_varlen_bwd_100(dout[:Tj], q[:Tj], k[:Tj], v[:Tj], cuj, cuj, S, S, None, None, None, None, None, 1, -1, 0, alpha), fl_j),
                    "uniform_256_targets": (lambda: ops.hstu_varlen_fwd_100(q, k, v, cu, cu, S, S, None, nt, 1, -1, 0, alpha),
                                            lambda: ops.hstu_varlen_bwd_100(dout, q, k, v, cu, cu, S, S, None, None, None, None, nt, 1, -1, 0, alpha), None)}
        for vn, (f_fwd, f_bwd, fl) in variants.items():
            out_v = {}
            for name, f in (("fwd", f_fwd), ("bwd", f_bwd)):
                for _ in range(2):
                    f()
                torch.cuda.synchronize()
                ts = []
                for _ in range(5):
                    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
                    e0.record(); f(); e1.record(); torch.cuda.synchronize()
                    ts.append(e0.elapsed_time(e1))
                out_v[name + "_ms"] = min(ts)
                if fl is not None:
                    out_v[name + "_tflops"] = fl * (1.0 if name == "fwd" else 2.5) / min(ts) / 1e9
            if vn.startswith("jagged"):
                out_v["tokens"] = Tj
            res[vn] = out_v
        res["config"] = "B=32 S=4096 causal H=8 D=128 bf16, q/k/v strided views of one (T, 4HD) buffer"
        # ---- the reference's own sm100 CuTe-DSL kernels on the same inputs, same box, same run (staged copy under baseline/_ref — never
        # /root/reference; JIT on first call).  ratio > 1 = ours faster.
        ref_dir = os.path.join(ROOT, "baseline", "_ref")
        if with_reference and os.path.isdir(os.path.join(ref_dir, "hstu_blackwell")):
            try:
                sys.path.insert(0, ref_dir)
                from hstu_blackwell import hstu_ops_gpu as refk
                t0 = time.time()
                rf = lambda: refk.hstu_varlen_fwd_100(q, k, v, cu, cu, S, S, None, None, 1, -1, 0, alpha, None, None)
                rf(); torch.cuda.synchronize()
                qc, kc, vc = q.contiguous(), k.contiguous(), v.contiguous()      # the reference bwd rejects the strided uvqk views
                rb = lambda: refk.hstu_varlen_bwd_100(dout, qc, kc, vc, cu, cu, S, S, None, None, None, None, None, 1, -1, 0, alpha, None, False, None, False)
                rb(); torch.cuda.synchronize()
                jit_s = time.time() - t0
                rfm, rbm = _p10(rf, iters), _p10(rb, iters)
                res["reference_sm100_kernels"] = {"fwd_ms": rfm, "bwd_ms": rbm, "fwd_tflops": flops_fwd / rfm / 1e9, "bwd_tflops": 2.5 * flops_fwd / rbm / 1e9,
                                                  "ratio_fwd": rfm / res["fwd"]["ms"], "ratio_bwd": rbm / res["bwd"]["ms"], "jit_s": jit_s,
                                                  "note": "hstu_blackwell CuTe-DSL kernels from baseline/_ref, same inputs; ratio = reference ms / ours ms"}
            except Exception as e:  # noqa: BLE001
                res["reference_sm100_kernels"] = {"error": repr(e)[:300]}
        return res
    except Exception as e:  # report, never kill the embedding line
        return {"error": repr(e)[:300]}


# ---------------------------------------------------------------------------------------------------------------- table setup
def make_module(capacity, dev):
    from dynamicemb import (BatchedDynamicEmbeddingTablesV2, DynamicEmbInitializerArgs, DynamicEmbInitializerMode, DynamicEmbPoolingMode,
                            DynamicEmbScoreStrategy, DynamicEmbTableOptions, EmbOptimType)
    opt = DynamicEmbTableOptions(dim=D, max_capacity=capacity, local_hbm_for_values=1 << 50, score_strategy=DynamicEmbScoreStrategy.STEP,
                                 initializer_args=DynamicEmbInitializerArgs(mode=DynamicEmbInitializerMode.UNIFORM, lower=-0.01, upper=0.01))
    m = BatchedDynamicEmbeddingTablesV2([opt], table_names=["t0"], pooling_mode=DynamicEmbPoolingMode.NONE, optimizer=EmbOptimType.EXACT_ADAGRAD,
                                        learning_rate=0.1, eps=1e-8, device=dev)
    m.train()
    return m


def prefill(m, gen, dev, world, rank, key_space, plaw_draws, to_full, capacity):
    """Power-law stream keys (score 1, rows initialised) for `plaw_draws` draws; then, if `to_full`, sequential filler keys outside the
    key space with score 0 until every bucket is full.  Keys a rank does not own under hash_roundrobin are skipped (sharded)."""
    from dynamicemb import dynamicemb_extensions as ext
    from dynamicemb.dynamicemb_extensions import ScorePolicy
    from dynamicemb.scored_hashtable import ScoreArg
    t0 = time.time()
    chunk = 1 << 24
    drawn = 0
    while drawn < plaw_draws and m.tables.size() < 0.95 * capacity:
        ids = torch.unique(power_law_ids(chunk, gen, dev, key_space))
        drawn += chunk
        if world > 1:
            ids = ids[(fmix64_torch(ids) % world) == rank]
        z = torch.zeros(ids.numel(), dtype=torch.int64, device=dev)
        _, found, _ = m.tables.lookup(ids, z, ScoreArg("score", None, ScorePolicy.CONST))
        ids = ids[~found]
        if ids.numel() == 0:
            continue
        z = torch.zeros(ids.numel(), dtype=torch.int64, device=dev)
        slots = m.tables.insert(ids, z, ScoreArg("score", torch.ones(ids.numel(), dtype=torch.int64, device=dev), ScorePolicy.ASSIGN))
        ext.init_rows(m._values, D, slots, ids, ext.InitializerMode.UNIFORM, -0.01, 0.01, seed=1)
    stream_keys = m.tables.size()
    filler = 0
    if to_full:
        nxt = FILLER_BASE + rank * (1 << 36)
        target = int(1.3 * capacity)
        while filler < target:
            k = min(chunk, target - filler)
            ids = torch.arange(nxt, nxt + k, dtype=torch.int64, device=dev)
            z = torch.zeros(k, dtype=torch.int64, device=dev)
            m.tables.insert(ids, z, ScoreArg("score", z, ScorePolicy.ASSIGN))          # score 0: first to be evicted
            nxt += k
            filler += k
    torch.cuda.synchronize()
    return {"stream_keys": stream_keys, "filler_keys_inserted": filler, "load": m.tables.size() / capacity, "seconds": round(time.time() - t0, 1)}


def time_steps(fn, batches, barrier):
    barrier()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for b in batches:
        fn(b)
    e1.record()
    barrier()
    return e0.elapsed_time(e1)


# ---------------------------------------------------------------------------------------------------------------- main arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--ids", type=int, default=1 << 20)
    ap.add_argument("--capacity", type=int, default=128 * 1024 * 1024)
    ap.add_argument("--plaw-draws", type=int, default=400_000_000, help="power-law draws of the prefill (the rest of the table is filler)")
    ap.add_argument("--no-hstu", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-variants", action="store_true", help="skip the no-eviction variant / cfg-5 sweep")
    ap.add_argument("--no-e2e", action="store_true", help="skip the SURVEY cfg-4 end-to-end harness (HSTU-large + DynamicEmb, ours and reference kernels)")
    ap.add_argument("--opt", default="", help="development toggles, e.g. 5=1,2=0 (demb_set_option)")
    ap.add_argument("--no-overlap-hits", action="store_true", help="N>1: copy the rows back to the requesters only after the whole owner prefetch (A/B)")
    ap.add_argument("--ncu", action="store_true", help="wrap 2 eager steps + 1 eval lookup in cudaProfilerStart/Stop (ncu --profile-from-start off) and exit")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        return run_reference_arm(args)

    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", 1))
    rank = int(os.environ.get("RANK", 0))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from dynamicemb import _native as N
    for kv in filter(None, args.opt.split(",")):
        N.lib.demb_set_option(int(kv.split("=")[0]), int(kv.split("=")[1]))

    n_ids = args.ids
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    m = make_module(args.capacity, dev)
    fill = prefill(m, gen, dev, world, rank, KEY_SPACE, args.plaw_draws, True, args.capacity)
    model = None
    if world > 1:
        from dynamicemb.shard import RowWiseShardedDynamicEmbedding
        # exchange capacities (ids per step): a rank feeds n_ids, sends at most n_ids/2 unique ids to any single owner (hash_roundrobin
        # spreads ~n_unique/W to each), and an owner accepts at most n_ids in total
        model = RowWiseShardedDynamicEmbedding(m, None, dist_type="hash_roundrobin", use_index_dedup=True, max_ids_per_step=n_ids,
                                               pair_capacity=n_ids // 2, recv_capacity=n_ids)
        model.overlap_hits = not args.no_overlap_hits
        samples = 4096                     # KJT shape per rank: one feature, 4096 samples x 256 ids (HSTU-like jagged sequences)
        lengths = torch.full((samples,), n_ids // samples, dtype=torch.int64, device=dev)
        call = lambda ids: model(ids, lengths)
        offsets = None
    else:
        model = None
        offsets = torch.arange(0, n_ids + 1, dtype=torch.int64, device=dev)
        call = lambda ids: m(ids, offsets)
    grad = torch.randn(n_ids, D, device=dev)
    total = args.steps + args.warmup
    batches = [power_law_ids(n_ids, gen, dev) for _ in range(total)]

    def step(ids):
        out = call(ids)
        out.backward(grad)
        return out

    for i in range(args.warmup):
        step(batches[i])
    if args.ncu:
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        for i in range(2):
            step(batches[args.warmup + i])
        if world == 1:
            m.eval()
            for opt_v in (1, 2):                               # the fused probe+gather forward (eval path), both tile-probe variants
                N.lib.demb_set_option(0, opt_v)
                m(batches[-1], offsets)
            N.lib.demb_set_option(0, 1)
            m.train()
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        return
    # ---- timed region: device-resident inputs.  N=1 runs the module's CUDA-graph step (make_graphed_step: prefetch + forward + fused
    # backward captured once — the fused prefetch has no host sync — and replayed; the id batch is copied device-to-device into the
    # graph's static input inside the timed region); the sharded wrapper provides its own step.
    dev_ids = torch.empty(n_ids, dtype=torch.int64, device=dev)
    graphed, graphed_noloss, graph_err = None, None, None
    try:
        dev_ids.copy_(batches[0])
        if world == 1:
            graphed = m.make_graphed_step(dev_ids, offsets, grad)                          # e2e: + loss stand-in (out.sum()) read back every step
            graphed_noloss = m.make_graphed_step(dev_ids, offsets, grad, with_loss=False)   # value: the step alone
        elif hasattr(model, "make_graphed_step"):
            graphed = model.make_graphed_step(dev_ids, lengths, grad)
            graphed_noloss = model.make_graphed_step(dev_ids, lengths, grad, with_loss=False)
    except Exception as e:   # noqa: BLE001  (fall back to the eager step, say so in the JSON)
        graphed, graphed_noloss, graph_err = None, None, repr(e)[:200]

    def timed_step(ids):
        if graphed_noloss is not None:
            dev_ids.copy_(ids, non_blocking=True)
            graphed_noloss[0].replay()
        else:
            step(ids)

    for i in range(args.warmup):
        timed_step(batches[i])
    if world > 1 and graphed_noloss is None:
        for i in range(20):               # eager sharded step: let the caching allocator settle before the timed steps
            timed_step(power_law_ids(n_ids, gen, dev))
    ms = max_over_ranks(time_steps(timed_step, batches[args.warmup:], barrier))
    value = world * n_ids * args.steps / (ms / 1e3)
    # ---- the same K steps through the module's plain forward()/backward() (no CUDA graph): the reference-shaped call sequence
    eager_batches = [power_law_ids(n_ids, gen, dev) for _ in range(args.steps)]
    for i in range(3):
        step(batches[i])
    ms_eager = max_over_ranks(time_steps(step, eager_batches, barrier))
    value_eager = world * n_ids * args.steps / (ms_eager / 1e3)

    # ---- per-kernel times for the roofline: a separate, un-reported pass of eager steps with CUDA events around every native launch
    N.lib.demb_profile_enable(1)
    N.PROFILE = {}
    N.LAUNCHES[0] = 0
    n_prof = min(args.steps, 8)
    nu_prof = []
    for i in range(n_prof):
        ids = power_law_ids(n_ids, gen, dev)                        # fresh ids: the same mix of hits / new keys as the timed steps
        step(ids)
    torch.cuda.synchronize()
    launches = N.LAUNCHES[0] // n_prof
    prof = {k: float(np.mean([a.elapsed_time(b) for a, b in v])) for k, v in N.PROFILE.items()}
    buf = (torch.zeros(3, dtype=torch.float32)).numpy()
    import ctypes
    N.lib.demb_profile_read(buf.ctypes.data_as(ctypes.c_void_p))
    bwd_stage_ms = buf.copy()
    N.PROFILE = None
    N.lib.demb_profile_enable(0)
    # the timed region lasts tens of ms — shorter than one nvidia-smi sample — so clocks / throttle reasons are sampled over a
    # ~2 s continuation of exactly the same steps (not part of any reported time); same step count on every rank
    n_cont = max(8, int(2000.0 / max(ms / args.steps, 1e-3)))
    clocks = ClockSampler(local)
    clocks.start()
    for i in range(n_cont):
        timed_step(batches[args.warmup + (i % args.steps)])
        if i % 8 == 7:
            torch.cuda.synchronize()
    torch.cuda.synchronize()
    clk = clocks.stop()
    clk["note"] = "sampled every 100 ms over a 2 s continuation of the timed steps"
    step_mode = ("cuda-graph step (make_graphed_step)" if graphed_noloss is not None
                 else ("eager step" + (f" (graph capture failed: {graph_err})" if graph_err else "")))

    # ---- e2e: the public API with HOST inputs: pinned ids -> H2D, step (prefetch + forward + fused backward), loss stand-in
    # (sum of the looked-up rows) -> D2H, every step.
    host_batches = [power_law_ids(n_ids, gen, dev).cpu().pin_memory() for _ in range(args.steps + 3)]      # fresh ids (3 for the warm-up pass)
    host_res = torch.zeros(len(host_batches), dtype=torch.float32).pin_memory()
    # Input pipeline as a trainer runs it: the H2D copy of step i+1 goes through a copy stream into one of two staging buffers while step
    # i computes; the 4-byte result of step i is read back asynchronously and consumed (host sync) one step later.  Every byte still
    # crosses PCIe inside the timed region, once per step, in both directions.
    copy_stream = torch.cuda.Stream(dev)
    cur = torch.cuda.current_stream(dev)
    staging = [torch.empty_like(dev_ids) for _ in range(2)]
    h2d_done = [torch.cuda.Event() for _ in range(2)]
    consumed = [torch.cuda.Event() for _ in range(2)]
    res_done = [torch.cuda.Event() for _ in range(len(host_batches))]
    for ev in consumed:
        ev.record(cur)

    def e2e_pass(hbs):
        for i, hb in enumerate(hbs):
            sl = i % 2
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(consumed[sl])
                staging[sl].copy_(hb, non_blocking=True)
                h2d_done[sl].record(copy_stream)
            cur.wait_event(h2d_done[sl])
            dev_ids.copy_(staging[sl], non_blocking=True)
            consumed[sl].record(cur)
            if graphed is not None:
                graphed[0].replay()
                host_res[i : i + 1].copy_(graphed[2].reshape(1), non_blocking=True)
            else:
                out = call(dev_ids)
                loss = out.detach().sum()
                out.backward(grad)
                host_res[i : i + 1].copy_(loss.reshape(1), non_blocking=True)
            res_done[i].record(cur)
            if i > 0:
                res_done[i - 1].synchronize()          # the host consumes step i-1's result here
        res_done[len(hbs) - 1].synchronize()

    e2e_pass(host_batches[:3])                       # untimed warm-up of exactly this path (first launch of this graph, copy stream, staging)
    barrier()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    e2e_pass(host_batches[3:])
    e1.record()
    barrier()
    e2e_value = world * n_ids * args.steps / (max_over_ranks(e0.elapsed_time(e1)) / 1e3)
    e2e_mode = step_mode + "; H2D of step i+1 on a copy stream (2 staging buffers) overlaps step i, result of step i read back async and consumed at step i+1"

    # ---- what one rank's local module sees per step (N>1: the ids it received, not the batch it fed)
    with torch.no_grad():
        probe_ids = batches[-1]
        if world > 1:
            mine = probe_ids.new_empty(0)
            allb = [torch.empty_like(probe_ids) for _ in range(world)]
            dist.all_gather(allb, probe_ids)
            got = torch.cat([torch.unique(b) for b in allb])         # per-rank dedup before the exchange
            got = got[(fmix64_torch(got) % world) == rank]
            n_local, nu_local = int(got.numel()), int(torch.unique(got).numel())
            nu_fed = int(torch.unique(probe_ids).numel())
        else:
            n_local, nu_local = n_ids, int(torch.unique(probe_ids).numel())
            nu_fed = nu_local

    extra = {}
    if rank == 0 or world > 1:
        hbm, tfl, which = peaks()
    # ---- roofline.  Algorithmic bytes, per launch, for the LOCAL module of this rank:
    #   fused lookup forward (eval: probe + gather in one kernel)  N_t*(8 id + 512 out) + N_u*(24 probe + 512 row)
    #   training forward = prefetch (unique + probe + insert/evict + init) + gather: same bytes + N_new*(1024 row write + 17 map)
    #   backward tiles kernel : N_t*(512 grad + 8 sorted pair) + N_u*(2*1024 row RW + 8)
    roof = None
    if rank == 0:
        b_fwd = n_local * (8 + 512) + nu_local * (24 + 512)
        b_bwd = n_local * (512 + 8) + nu_local * (2 * 1024 + 8)
        traffic = {}
        try:   # dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` capture of this command
            with open(os.path.join(ROOT, "profiles", "kernel_traffic.json")) as f:
                traffic = json.load(f)
        except (OSError, ValueError):
            pass
        other = {**{k: round(v, 4) for k, v in prof.items()}, "backward.pairs+sort": float(bwd_stage_ms[0]),
                 "backward.tiles": float(bwd_stage_ms[1]), "backward.spans": float(bwd_stage_ms[2])}
        roof = {"bound": "hbm", "peak": hbm, "unit": "GB/s", "peak_source": which, "other_kernels_ms": other,
                "local_ids_per_step": n_local, "local_unique_per_step": nu_local}
        if world == 1:
            # the north_star's headline kernel: fused probe + gather forward (eval path) on a batch of the same stream
            m.eval()
            ev = []
            for i in range(7):
                a, b = torch.cuda.Event(True), torch.cuda.Event(True)
                a.record(); m(batches[-1 - i], offsets); b.record()
                torch.cuda.synchronize()
                ev.append(a.elapsed_time(b))
            # the pooled (EmbeddingBagCollection) form of the same lookup: 2 features x 52428 bags x hotness 10, SUM, fused probe + pool
            from dynamicemb import dynamicemb_extensions as ext
            Fp, Bp = 2, n_ids // 20
            np_ids = Fp * Bp * 10
            poff = torch.arange(0, np_ids + 1, 10, dtype=torch.int64, device=dev)
            tb = m.tables
            pool_call = lambda k: ext.lookup_forward(tb.table_storage_, tb.table_bucket_offsets_, tb.bucket_capacity_, m._values, D, batches[-1 - k][:np_ids],
                                                     row_base=tb.row_base_, offsets=poff, batch_size=Bp, num_features=Fp, combiner=0, num_scores=tb.num_scores_)
            pev = []
            for i in range(7):
                a, b = torch.cuda.Event(True), torch.cuda.Event(True)
                a.record(); pool_call(i); b.record()
                torch.cuda.synchronize()
                pev.append(a.elapsed_time(b))
            pms = sorted(pev)[len(pev) // 2]
            nu_p = int(torch.unique(batches[-1][:np_ids]).numel())
            b_pool = np_ids * 8 + nu_p * (24 + 512) + Fp * Bp * 512
            b_pool_nt = np_ids * (8 + 512) + Fp * Bp * 512            # SURVEY 8(a) accounting of gather_embedding_pooled: every id's row is read
            roof["pooled_lookup_forward"] = {"kernel": "forward_pool_kernel (fused tile probe + SUM pooling, hotness 10)", "ms": pms, "algorithmic_bytes": b_pool,
                                             "achieved_GBps": b_pool / pms / 1e6, "frac": b_pool / pms / 1e6 / hbm, "ids": np_ids, "bags": Fp * Bp,
                                             "frac_counting_every_id_row": b_pool_nt / pms / 1e6 / hbm,
                                             "note": "algorithmic_bytes counts each UNIQUE row once (hot rows are re-read from L2); the second fraction counts "
                                                     "N_t x 512 B of row reads as SURVEY 8(a) does for gather_embedding_pooled"}
            # A/B of the three fused-lookup kernels on the same batches (option 0 of demb_set_option; 1 is the shipped default)
            variants = {}
            for opt_v, nm in ((0, "round-1 thread-per-key probe (forward_seq_tma_kernel)"), (2, "specialised probe / copy warps (forward_seq_probe2_kernel)"),
                              (3, "12 copy + 20 small probe warps (forward_seq_probe3_kernel<20>)"), (4, "12 copy + 12 small probe warps (forward_seq_probe3_kernel<12>)"),
                              (5, "per-warp pipeline, second-generation tile probe (forward_seq_probe_kernel<2>)"),
                              (1, "one probe+copy pipeline per warp (forward_seq_probe_kernel, default)")):
                N.lib.demb_set_option(0, opt_v)
                tv = []
                for i in range(7):
                    a, b = torch.cuda.Event(True), torch.cuda.Event(True)
                    a.record(); m(batches[-1 - i], offsets); b.record()
                    torch.cuda.synchronize()
                    tv.append(a.elapsed_time(b))
                variants[nm] = sorted(tv)[len(tv) // 2]
            N.lib.demb_set_option(0, 1)
            roof["fused_lookup_kernel_variants_ms"] = variants
            m.train()
            fms = sorted(ev)[len(ev) // 2]
            roof.update({"kernel": "forward_seq_probe_kernel (demb_lookup_forward: fused hash probe + row gather, the 128-d lookup path)",
                         "achieved": b_fwd / fms / 1e6, "frac": b_fwd / fms / 1e6 / hbm, "ms_per_launch": fms, "algorithmic_bytes_per_launch": b_fwd,
                         "traffic": traffic.get("forward_seq_probe_kernel"), "lookups_per_s": n_ids / fms * 1e3})
            tf_ms = prof.get("train_prefetch", 0.0) + prof.get("gather_forward", 0.0)
            roof["train_forward"] = {"kernels": "demb_train_prefetch (unique, probe, insert/evict, row init) + forward_seq_tma_kernel (gather)",
                                     "ms": tf_ms, "prefetch_ms": prof.get("train_prefetch"), "gather_ms": prof.get("gather_forward"),
                                     "algorithmic_bytes": b_fwd, "achieved_GBps": b_fwd / tf_ms / 1e6 if tf_ms > 0 else None,
                                     "frac": b_fwd / tf_ms / 1e6 / hbm if tf_ms > 0 else None,
                                     "note": "bytes exclude the rows written for new keys (insert is not on the steady-state roofline)"}
        else:
            roof.update({"kernel": "backward_tiles_kernel (local module of rank 0; per-rank byte counts)", "achieved": None, "frac": None,
                         "ms_per_launch": float(bwd_stage_ms[1]), "algorithmic_bytes_per_launch": b_bwd, "traffic": None})
            if bwd_stage_ms[1] > 0:
                roof["achieved"] = b_bwd / float(bwd_stage_ms[1]) / 1e6
                roof["frac"] = roof["achieved"] / hbm
        bt = float(bwd_stage_ms[1])
        roof["backward_tiles_kernel"] = {"ms": bt, "algorithmic_bytes": b_bwd, "achieved_GBps": b_bwd / bt / 1e6 if bt > 0 else None,
                                         "frac": b_bwd / bt / 1e6 / hbm if bt > 0 else None, "traffic": traffic.get("backward_tiles_kernel"),
                                         "note": "dominant kernel of the step by time"}
        if world > 1:
            # wire bytes per rank per step (SURVEY cfg 5): (W-1)/W of the unique ids out (8 B) and their rows back (512 B), and the same
            # number of gradient rows in backward
            sent = nu_fed * (world - 1) / world
            roof["wire"] = {"ids_out_bytes": sent * 8, "rows_back_bytes": sent * 512, "grad_rows_out_bytes": sent * 512,
                            "GBps_per_direction_at_this_step_time": sent * 512 * 2 / (ms / args.steps) / 1e6,
                            "nvlink5_unidirectional_GBps": 900.0}
    table_load = m.tables.size() / args.capacity

    # ---- variants (N=1): the same graph step without eviction (SURVEY §8(d) cfg 2: "separately report a no-eviction run, key space 64 Mi")
    if world == 1 and not args.no_variants:
        try:
            graphed = graphed_noloss = model = None
            graphed = graphed_noloss = None
            m = None
            torch.cuda.empty_cache()
            ks2 = 64 * 1024 * 1024
            m2 = make_module(args.capacity, dev)
            gen2 = torch.Generator(device=dev).manual_seed(99)
            fill2 = prefill(m2, gen2, dev, 1, 0, ks2, 1 << 28, False, args.capacity)
            b2 = [power_law_ids(n_ids, gen2, dev, ks2) for _ in range(total)]
            dev_ids.copy_(b2[0])
            g2 = m2.make_graphed_step(dev_ids, offsets, grad, with_loss=False)

            def s2(ids):
                dev_ids.copy_(ids, non_blocking=True)
                g2[0].replay()

            for i in range(args.warmup):
                s2(b2[i])
            ms2 = time_steps(s2, b2[args.warmup:], barrier)
            m2.eval()
            ev = []
            for i in range(5):
                a, b = torch.cuda.Event(True), torch.cuda.Event(True)
                a.record(); m2(b2[-1 - i], offsets); b.record(); torch.cuda.synchronize()
                ev.append(a.elapsed_time(b))
            f2 = sorted(ev)[2]
            nu2 = int(torch.unique(b2[-1]).numel())
            bf2 = n_ids * (8 + 512) + nu2 * (24 + 512)
            extra["no_eviction"] = {"key_space": ks2, "table_load": m2.tables.size() / args.capacity, "value": n_ids * args.steps / (ms2 / 1e3),
                                    "ms_per_step": ms2 / args.steps, "unique_per_step": nu2,
                                    "fused_lookup_forward": {"ms": f2, "frac": bf2 / f2 / 1e6 / hbm, "achieved_GBps": bf2 / f2 / 1e6}, "prefill": fill2}
            del g2, m2
            torch.cuda.empty_cache()
        except Exception as e:  # noqa: BLE001
            extra["no_eviction"] = {"error": repr(e)[:300]}
    # ---- SURVEY cfg 5 (N>1): batch sweep 2^14..2^20 ids per GPU per step on a 1e10 key space, through the sharded wrapper (eager)
    if world > 1 and not args.no_variants:
        sweep = []
        try:
            for lg in range(14, 21):
                nb = 1 << lg
                smp = min(4096, nb // 4)
                lens = torch.full((smp,), nb // smp, dtype=torch.int64, device=dev)
                gb = torch.randn(nb, D, device=dev)
                bs = [power_law_ids(nb, gen, dev, 10_000_000_000) for _ in range(8)]

                def s5(ids):
                    o = model(ids, lens)
                    o.backward(gb)

                for b in bs[:3]:
                    s5(b)
                t5 = max_over_ranks(time_steps(s5, bs[3:], barrier)) / 5
                nu5 = int(torch.unique(bs[-1]).numel())
                sent = nu5 * (world - 1) / world
                sweep.append({"ids_per_gpu": nb, "ms_per_step": t5, "lookups_per_s": world * nb / t5 * 1e3,
                              "wire_GBps_per_direction": sent * 512 * 2 / t5 / 1e6, "unique_per_rank": nu5})
            extra["cfg5_batch_sweep_keyspace1e10"] = sweep
        except Exception as e:  # noqa: BLE001
            extra["cfg5_batch_sweep_keyspace1e10"] = {"error": repr(e)[:300], "partial": sweep}

    # ---- SURVEY cfg 4: HSTU-large + DynamicEmb end-to-end step, our kernels and the reference's own GPU kernels in the same harness
    # (tools/e2e_harness.py); every rank takes part (the embedding is row-wise sharded at N > 1)
    if not args.no_e2e:
        try:
            graphed = graphed_noloss = None
            if model is not None and getattr(model, "_buf", None) is not None:
                torch.cuda.synchronize()
                dist.barrier()
                model._buf.close()
            model = m = None
            batches = host_batches = eager_batches = None
            torch.cuda.empty_cache()
            from tools import e2e_harness
            extra["cfg4_e2e_hstu_large"] = e2e_harness.run_both(dev, world, rank)
        except Exception as e:  # noqa: BLE001
            extra["cfg4_e2e_hstu_large"] = {"error": repr(e)[:300]}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic", "config": {**workload_config(args, n_ids), "step_mode": step_mode}, "clocks": clk,
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": n_ids * 8, "d2h_bytes_per_step": 4, "mode": e2e_mode},
                "gpu_launches": launches, "roofline": roof, "table_load": table_load, "prefill": fill,
                "value_eager": {"value": value_eager, "ms_per_step": ms_eager / args.steps,
                                "mode": "module.forward(ids, offsets); out.backward(grad) per step, no CUDA graph (the reference-shaped call sequence)"},
                **extra}
        if world == 1 and not args.no_cpu:
            line["cpu_baseline"] = cpu_baseline()
            line["cpu_cfg1_embeddingbag"] = cpu_cfg1_embeddingbag()
            line["cpu_eager_hstu"] = cpu_eager_hstu()
        if world == 1 and not args.no_hstu:
            batches = host_batches = None
            torch.cuda.empty_cache()
            line["hstu_attn"] = bench_hstu(dev, tfl)
            try:                                   # HSTU layer glue kernels at the HSTU-large shape, eager torch beside them
                torch.cuda.empty_cache()
                from tools import bench_hstu_glue
                line["hstu_layer_glue"] = bench_hstu_glue.run(dev)
            except Exception as e:  # noqa: BLE001
                line["hstu_layer_glue"] = {"error": repr(e)[:300]}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
