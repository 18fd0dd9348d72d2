"""CPU suite: the oracle against fixtures produced by the reference's own Python (tests/golden/gen_golden_cpu.py), host logic, and that
the C-ABI library loads and exports every symbol include/*.h declares (no compute calls: there is no GPU here)."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "tests", "golden")


def test_hash_matches_reference_python():
    from oracle import dynamicemb as orc
    from dynamicemb.scored_hashtable import murmur3_hash_64bits
    z = np.load(os.path.join(G, "hash_vectors.npz"))
    for k, want in zip(z["keys"].view(np.uint64).tolist(), z["fmix64"].tolist()):
        assert orc.fmix64(k) == want
        assert orc.hash63(k) == want & 0x7FFFFFFFFFFFFFFF          # types.cuh:123-131
        assert murmur3_hash_64bits(k) == want
    assert orc.lib().orc_empty_digest() == (orc.fmix64(0xFFFFFFFFFFFFFFFF) & 0x7FFFFFFFFFFFFFFF) >> 32 & 0xFF


def test_hstu_mask_matches_reference_construct_mask():
    from oracle.hstu_attn import build_mask
    z = np.load(os.path.join(G, "hstu_mask.npz"))
    for c in range(int(z["ncases"])):
        B, seqlen, seqlen_c, seqlen_t, Gs, w0, w1, N = z[f"c{c}_meta"].tolist()
        hist, nc, nt = z[f"c{c}_hist"], z[f"c{c}_nc"], z[f"c{c}_nt"]
        lens = (hist + nc + nt).tolist()
        causal = (w0 < 0 and w1 == 0)
        m = build_mask(lens, N, nc if (seqlen_c and causal) else None, nt if (seqlen_t and causal) else None, Gs, (w0, w1)).numpy()
        ref = z[f"c{c}_mask"]
        for b in range(B):
            L = lens[b]
            assert np.array_equal(m[b, :L, :L], ref[b, :L, :L]), f"case {c} batch {b}"
            assert not m[b, L:, :].any() and not m[b, :, L:].any()


def test_hstu_oracle_matches_reference_eager():
    from oracle.hstu_attn import hstu_attention
    z = np.load(os.path.join(G, "hstu_eager.npz"))
    for c in range(int(z["ncases"])):
        H, D, Gs, scaling = z[f"c{c}_meta"].tolist()
        lens = z[f"c{c}_lens"].tolist()
        nt = None if z[f"c{c}_nt"][0] < 0 else z[f"c{c}_nt"]
        nc = None if z[f"c{c}_nc"][0] < 0 else z[f"c{c}_nc"]
        cu = np.concatenate([[0], np.cumsum(lens)])
        q, k, v = (torch.from_numpy(z[f"c{c}_{n}"]) for n in "qkv")
        out = hstu_attention(q, k, v, cu, max(lens), 1.0 / D ** 0.5, scaling, nc, nt, Gs, (-1, 0))
        torch.testing.assert_close(out, torch.from_numpy(z[f"c{c}_out"]), rtol=1e-5, atol=1e-6)


def test_c_abi_exports_every_declared_symbol():
    lib = ctypes.CDLL(os.path.join(ROOT, "recsys-examples_b200", "lib", "librecsys_b200.so"))
    names = set()
    for h in ("dynamicemb_b200.h", "hstu_b200.h"):
        src = open(os.path.join(ROOT, "include", h)).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        names |= set(re.findall(r"\b(?:int|int64_t)\s+((?:demb|hstu|sm100)_\w+)\s*\(", src))
    assert len(names) >= 24
    for n in sorted(names):
        assert hasattr(lib, n), f"{n} declared in include/ but not exported"


def test_no_cpu_fallback_in_product():
    """The product must not import the oracle (tier rule); grep the package."""
    pkg = os.path.join(ROOT, "recsys-examples_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith(".py"):
                s = open(os.path.join(dp, f)).read()
                assert "oracle" not in s.replace("# oracle", ""), f"{f} mentions the oracle"


def test_oracle_table_invariants():
    """test_table_operation.py:277-526 invariants on the oracle itself."""
    from oracle.dynamicemb import OracleTable
    rng = np.random.default_rng(0)
    o = OracleTable([128 * 4], 128)
    keys = np.unique(rng.integers(0, 1 << 60, size=400, dtype=np.int64))
    idx, res, _, _ = o.insert(keys, None, policy=1, score_in=np.ones(keys.size, dtype=np.int64))
    assert (res == 0).all() and len(set(idx.tolist())) == keys.size
    idx2, res2, _, _ = o.insert(keys, None, policy=1, score_in=np.full(keys.size, 2, dtype=np.int64))
    assert (res2 == 2).all() and np.array_equal(idx, idx2)
    _, f, i = o.lookup(keys, None, policy=0)
    assert f.all() and np.array_equal(i, idx)
    o.erase(keys[:100])
    _, f, _ = o.lookup(keys[:100], None, policy=0)
    assert not f.any()
    # fill to capacity then overflow => EVICT of the minimum-score slot; erased slots are reclaimed first only when no empty slot is left
    more = np.unique(rng.integers(1 << 60, 1 << 61, size=700, dtype=np.int64))
    _, res3, _, (ek, ei, es, _) = o.insert(more, None, policy=1, score_in=np.full(more.size, 5, dtype=np.int64))
    assert set(res3.tolist()) <= {0, 1, 3}
    assert (res3 == 1).sum() > 0 and (res3 == 3).sum() > 0
    assert o.bucket_sizes.sum() <= 512


def test_get_sharded_table_capacity():
    from dynamicemb import get_sharded_table_capacity
    assert get_sharded_table_capacity(1000, 8, 128) == 128
    assert get_sharded_table_capacity(10_000_000_000, 8, 128) == 1_250_000_000 // 128 * 128 + (128 if 1_250_000_000 % 128 else 0)
    with pytest.raises(ValueError):
        get_sharded_table_capacity(10, 0, 128)


def test_bench_reference_arm_prints_contract_line():
    """`bench.py --impl reference` (the CPU port of the reference path) runs without a GPU and prints ONE JSON line with the contract keys."""
    import json
    import subprocess
    import sys
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "3"],
                         capture_output=True, text=True, timeout=600, env={**os.environ, "RANK": "0", "WORLD_SIZE": "1"})
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.strip().splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "dtype", "data", "config",
              "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["value"] > 0 and d["cpu_baseline"]["kind"] in ("port", "reference")
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0


def test_oracle_table_random_op_sequences():
    """Property check of the oracle table on random insert / erase / lookup sequences (small table, so buckets fill up and evict).
    After every operation: the key image has no duplicates; lookup finds exactly the keys of the image, at their slots; bucket_sizes
    counts them; a successfully inserted key that was not evicted later is there; an evicted key that was not re-inserted is gone; the
    four reserved key values are rejected as ILLEGAL; erased keys are gone."""
    from hypothesis import given, settings, strategies as st
    from oracle.dynamicemb import OracleTable

    @settings(max_examples=25, deadline=None)
    @given(st.integers(0, 2 ** 31), st.lists(st.sampled_from(["ins", "ins", "ins", "erase", "look"]), min_size=3, max_size=12))
    def run(seed, ops):
        rng = np.random.default_rng(seed)
        C = 16
        o = OracleTable([C * 3], C)
        step = 0

        def image():
            img = np.ascontiguousarray(o.keys_view()).reshape(-1).view(np.int64)
            pos = np.nonzero((img != -1) & (img != -2) & (img != -3))[0]          # EmptyKey / ReclaimKey / LockedKey as int64
            return {int(img[p]): int(p) for p in pos}, pos.size

        for op in ops:
            step += 1
            if op == "ins":
                keys = np.unique(rng.integers(-400, 400, size=int(rng.integers(1, 40)), dtype=np.int64))
                idx, res, _, (ek, ei, es, _) = o.insert(keys, None, policy=1, score_in=np.full(keys.size, step, dtype=np.int64), use_counter=False)
                ok = {int(k) for k, r in zip(keys.tolist(), res.tolist()) if r <= 3}
                evicted = {int(k) for k in np.ascontiguousarray(ek).view(np.int64).tolist()}
                for k, i, r in zip(keys.tolist(), idx.tolist(), res.tolist()):
                    if -4 <= k <= -1:
                        assert r == 6 and i < 0               # reserved values (types.cuh:117-121)
                    else:
                        assert r <= 3 or (r == 5 and i < 0)   # BUSY is the only failure of a legal key
                now, _ = image()
                assert all(k in now for k in ok - evicted)
                assert all(k not in now for k in evicted - ok)
            elif op == "erase":
                now, _ = image()
                if now:
                    drop = rng.choice(np.array(sorted(now), dtype=np.int64), size=max(1, len(now) // 2), replace=False)
                    o.erase(drop)
                    after, _ = image()
                    assert all(int(k) not in after for k in drop.tolist())
            now, n_live = image()
            assert n_live == len(now)                                             # no duplicate keys
            assert int(o.bucket_sizes.sum()) == n_live
            probe = np.unique(np.concatenate([np.array(sorted(now), dtype=np.int64), rng.integers(-400, 400, size=20, dtype=np.int64)]))
            _, found, slots = o.lookup(probe, None, policy=0)
            for k, f, s in zip(probe.tolist(), found.tolist(), slots.tolist()):
                assert bool(f) == (int(k) in now), (k, f)
                if f:
                    assert s == now[int(k)]

    run()


def test_ctypes_call_sites_match_the_declared_signatures():
    """Every call of a C-ABI entry point in the host package passes exactly as many arguments as its ctypes signature (and the header)
    declare — a mismatch only shows up on a GPU box otherwise."""
    from dynamicemb import _native as N
    pkg = os.path.join(ROOT, "recsys-examples_b200")
    src = "".join(open(os.path.join(pkg, "dynamicemb", f)).read() for f in ("dynamicemb_extensions.py", "shard.py", "batched_dynamicemb_tables.py"))

    def count_args(s, start):
        depth, n, i = 0, 1, start
        while True:
            c = s[i]
            if c in "([{":
                depth += 1
            elif c in ")]}":
                if depth == 0:
                    return n
                depth -= 1
            elif c == "," and depth == 0:
                n += 1
            i += 1

    seen = 0
    for pat in (r"N\.lib\.(demb_\w+)\(", r"N\.launch\(\"[\w_]+\", \d+, N\.lib\.(demb_\w+), "):
        for m in re.finditer(pat, src):
            name = m.group(1)
            assert count_args(src, m.end()) == len(N._SIGS[name][1]), f"call of {name} does not match its ctypes signature"
            seen += 1
    assert seen >= 30
    # and the header agrees with the ctypes table on the parameter count
    hdr = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "dynamicemb_b200.h")).read(), flags=re.S)
    for name, (_, args) in N._SIGS.items():
        m = re.search(r"\b" + name + r"\s*\(([^;]*?)\)\s*;", hdr, flags=re.S)
        assert m, f"{name} not declared in include/dynamicemb_b200.h"
        params = m.group(1).strip()
        n_decl = 0 if params in ("", "void") else params.count(",") + 1
        assert n_decl == len(args), f"{name}: header declares {n_decl} parameters, ctypes table {len(args)}"
