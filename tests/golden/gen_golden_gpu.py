"""Golden-vector generator, GPU side (run on the B200 box through gpurun; NOT part of the test suite).

Drives the UNMODIFIED reference kernels — `dynamicemb_extensions`, built from
/root/reference/corelib/dynamicemb/src by baseline/build_ref_dynamicemb.py into baseline/_ref/ — exactly the
way the reference's own Python does (scored_hashtable.py:378-425 storage layout, :1451-1557 deterministic
wave insert) and stores inputs + outputs as small .npz fixtures under gpurun_out/golden/, which are then
committed under tests/golden/.  tests/test_golden.py replays them against the CPU oracle (no GPU) and
against the CUDA path (GPU).

  gpurun -- python tests/golden/gen_golden_gpu.py
"""
import importlib.util
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
SO = os.path.join(ROOT, "baseline", "_ref", "dynamicemb_ext", "dynamicemb_extensions.so")
OUT = os.path.join(ROOT, "gpurun_out", "golden")


def load_ref():
    spec = importlib.util.spec_from_file_location("dynamicemb_extensions", SO)
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def murmur3_fmix64(k):
    k &= 0xFFFFFFFFFFFFFFFF
    k ^= k >> 33
    k = (k * 0xFF51AFD7ED558CCD) & 0xFFFFFFFFFFFFFFFF
    k ^= k >> 33
    k = (k * 0xC4CEB9FE1A85EC53) & 0xFFFFFFFFFFFFFFFF
    k ^= k >> 33
    return k


class RefTable:
    """scored_hashtable.py LinearBucketTable.__init__/_init_table restated with the reference's own ops."""

    def __init__(self, ref, caps, C, dev):
        self.ref, self.C, self.dev = ref, C, dev
        nbs = [(c + C - 1) // C for c in caps]
        off = [0]
        for nb in nbs:
            off.append(off[-1] + nb)
        self.nb = off[-1]
        self.off = torch.tensor(off, dtype=torch.int64, device=dev)
        self.storage = torch.empty(17 * C * self.nb, dtype=torch.uint8, device=dev)
        self.keys_, self.digests_, self.scores_ = ref.table_partition(self.storage, [torch.int64, torch.uint8, torch.uint64], C, self.nb)
        self.keys_.fill_(-1)
        self.scores_.fill_(0)
        self.digests_.fill_((murmur3_fmix64(0xFFFFFFFFFFFFFFFF) >> 32) & 0xFF)
        self.bucket_sizes = torch.zeros(self.nb, dtype=torch.int32, device=dev)
        self.counter = torch.zeros(self.nb * C, dtype=torch.int32, device=dev)

    def _waves(self, keys, tids, scores):
        """_bucketize_and_pad (scored_hashtable.py:1451-1510)"""
        bkt_keys, offsets, inverse = self.ref.bucketize_keys(keys, tids, self.off, self.nb, self.C)
        bkt_tids = tids[inverse]
        bkt_scores = scores[inverse]
        lengths = offsets[1:] - offsets[:-1]
        nbk = offsets.numel() - 1
        max_len = int(lengths.max().item())
        bucket_idx = torch.repeat_interleave(torch.arange(nbk, device=self.dev), lengths)
        starts = torch.repeat_interleave(offsets[:-1], lengths)
        positions = torch.arange(bkt_keys.numel(), device=self.dev) - starts
        pk = torch.full((nbk, max_len), -1, dtype=keys.dtype, device=self.dev)
        pt = torch.zeros((nbk, max_len), dtype=torch.int64, device=self.dev)
        ps = torch.zeros((nbk, max_len), dtype=torch.int64, device=self.dev)
        pk[bucket_idx, positions] = bkt_keys
        pt[bucket_idx, positions] = bkt_tids
        ps[bucket_idx, positions] = bkt_scores
        return pk.t().contiguous(), pt.t().contiguous(), ps.t().contiguous()

    def det_insert(self, keys, tids, scores, policy, evict=False):
        """_deterministic_insert[_and_evict] (:1512-1640).  Returns (indices, evicted triples sorted)."""
        ev = []
        kt, tt, st = self._waves(keys, tids, scores)
        for i in range(kt.size(0)):
            valid = kt[i] != -1
            if not valid.any():
                continue
            vk, vt, vs = kt[i][valid].contiguous(), tt[i][valid].contiguous(), st[i][valid].contiguous().view(torch.uint64)
            if evict:
                _idx, nev, ek, ei, es, et = self.ref.table_insert_and_evict(self.storage, self.off, self.C, self.bucket_sizes, vk, vt, vs, policy, self.counter)
                h = int(nev.cpu().item())
                if h:
                    ev.append((ek[:h].cpu().numpy(), ei[:h].cpu().numpy(), es[:h].cpu().numpy().astype(np.int64)))
            else:
                self.ref.table_insert(self.storage, self.off, self.C, self.bucket_sizes, vk, vt, vs, policy, self.counter)
        _, founds, idx = self.ref.table_lookup(self.storage, self.off, self.C, keys, tids, None, self.ref.ScorePolicy.CONST)
        return idx, ev


def uniq_keys(rng, n):
    k = np.unique(rng.integers(-(1 << 62), 1 << 62, size=3 * n, dtype=np.int64))
    rng.shuffle(k)
    return k[:n]


def main():
    os.makedirs(OUT, exist_ok=True)
    ref = load_ref()
    dev = torch.device("cuda", 0)
    P = ref.ScorePolicy
    rng = np.random.default_rng(20260924)
    report = []

    # ---------------- case A: single table, ASSIGN scores, fill + overflow (evict) + erase/reclaim ----------------
    for name, caps, C, nsteps, n in [("single_c128", [128 * 32], 128, 6, 1500), ("multi_c64", [64 * 9, 64 * 4, 64 * 30], 64, 8, 900),
                                      ("one_bucket_c16", [16], 16, 4, 40), ("c1024", [1024 * 2], 1024, 3, 1200)]:
        t = RefTable(ref, caps, C, dev)
        rec = {"caps": np.array(caps), "C": np.array(C)}
        for s in range(nsteps):
            keys = uniq_keys(rng, n)
            tids = rng.integers(0, len(caps), size=n).astype(np.int64)
            scores = rng.integers(1, 1000, size=n, dtype=np.int64) if s % 2 == 0 else np.full(n, 1000 + s, dtype=np.int64)
            kt, tt, st = (torch.from_numpy(a).to(dev) for a in (keys, tids, scores))
            idx, ev = t.det_insert(kt, tt, st, P.ASSIGN, evict=True)
            rec[f"s{s}_keys"], rec[f"s{s}_tids"], rec[f"s{s}_scores"] = keys, tids, scores
            rec[f"s{s}_indices"] = idx.cpu().numpy()
            rec[f"s{s}_image"] = t.storage.cpu().numpy().copy()
            rec[f"s{s}_bucket_sizes"] = t.bucket_sizes.cpu().numpy().copy()
            if ev:
                rec[f"s{s}_ev_keys"] = np.concatenate([e[0] for e in ev])
                rec[f"s{s}_ev_idx"] = np.concatenate([e[1] for e in ev])
                rec[f"s{s}_ev_scores"] = np.concatenate([e[2] for e in ev])
            if s % 3 == 1:   # erase some present keys
                present = idx.cpu().numpy() >= 0
                ek = keys[present][: n // 5]
                et = tids[present][: n // 5]
                ref.table_erase(t.storage, t.off, C, t.bucket_sizes, torch.from_numpy(ek).to(dev), torch.from_numpy(et).to(dev))
                rec[f"s{s}_erase_keys"], rec[f"s{s}_erase_tids"] = ek, et
                rec[f"s{s}_image_after_erase"] = t.storage.cpu().numpy().copy()
            # lookup with score mutation (ASSIGN) of a subset + unknown keys
            q = np.concatenate([keys[: n // 3], uniq_keys(rng, 20)])
            qt = np.concatenate([tids[: n // 3], rng.integers(0, len(caps), size=20)]).astype(np.int64)
            qs = np.full(q.size, 5000 + s, dtype=np.int64)
            so, fo, io = ref.table_lookup(t.storage, t.off, C, torch.from_numpy(q).to(dev), torch.from_numpy(qt).to(dev),
                                          torch.from_numpy(qs).to(dev).view(torch.uint64), P.ASSIGN)
            rec[f"s{s}_q_keys"], rec[f"s{s}_q_tids"], rec[f"s{s}_q_scores"] = q, qt, qs
            rec[f"s{s}_q_founds"], rec[f"s{s}_q_indices"], rec[f"s{s}_q_score_out"] = fo.cpu().numpy(), io.cpu().numpy(), so.cpu().numpy()
            rec[f"s{s}_image_after_lookup"] = t.storage.cpu().numpy().copy()
        rec["nsteps"] = np.array(nsteps)
        np.savez_compressed(os.path.join(OUT, f"table_{name}.npz"), **rec)
        report.append(f"table_{name}: {nsteps} steps, image {t.storage.numel()} B")

    # ---------------- case B: row ops (fp32): pooled gather, reduce_grads, adagrad/adam/sgd/rowwise flat-table update ----------------
    D, B, F = 128, 16, 3
    nu = 300
    lens = rng.integers(0, 40, size=B * F)
    offsets = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    n = int(offsets[-1])
    inverse = np.minimum(rng.zipf(1.3, size=n) - 1, nu - 1).astype(np.int64)
    inverse[:nu] = np.arange(nu)
    uemb = torch.randn(nu, D, device=dev)
    rec = {"D": np.array(D), "B": np.array(B), "F": np.array(F), "offsets": offsets, "inverse": inverse, "unique_embs": uemb.cpu().numpy()}
    for comb in (0, 1):
        out = torch.empty(B, F * D, device=dev)
        ref.gather_embedding_pooled(uemb, out, torch.from_numpy(inverse).to(dev), torch.from_numpy(offsets).to(dev), comb, F * D, B)
        rec[f"pooled_{comb}"] = out.cpu().numpy()
        g = torch.randn(B, F * D, device=dev)
        ug = ref.reduce_grads(torch.from_numpy(inverse).to(dev), g, nu, B, D, torch.from_numpy(offsets).to(dev), None, comb, F * D)
        rec[f"grads_{comb}"], rec[f"ugrads_{comb}"] = g.cpu().numpy(), ug.cpu().numpy()
    out = torch.empty(n, D, device=dev)
    ref.gather_embedding(uemb, out, torch.from_numpy(inverse).to(dev))
    rec["seq"] = out.cpu().numpy()
    # (sequence-mode reduce_grads is exercised through the pooled variants: same LocalReduce kernels, dynamic_emb_op.cu:160-283)
    # optimizers on a flat table
    ug = torch.randn(nu, D, device=dev)
    rows = torch.from_numpy(rng.permutation(1000)[:nu].astype(np.int64)).to(dev)
    rec["opt_grads"], rec["opt_rows"] = ug.cpu().numpy(), rows.cpu().numpy()
    tids = torch.zeros(nu, dtype=torch.int64, device=dev)
    for nm, state in [("sgd", 0), ("adagrad", D), ("adam", 2 * D), ("rowwise", 4)]:
        vdim = D + state
        table = (torch.rand(1000, vdim, device=dev) + 0.1).contiguous()
        rec[f"opt_{nm}_before"] = table.cpu().numpy().copy()
        ptrs = torch.tensor([table.data_ptr()], dtype=torch.int64, device=dev)
        vd = torch.tensor([vdim], dtype=torch.int64, device=dev)
        ed = torch.tensor([D], dtype=torch.int64, device=dev)
        if nm == "sgd":
            ref.sgd_update_for_flat_table(ug, rows, ptrs, tids, vd, ed, D, True, 0.05, 0)
        elif nm == "adagrad":
            ref.adagrad_update_for_flat_table(ug, rows, ptrs, tids, vd, ed, 0.05, 1e-8, D, True, 0)
        elif nm == "adam":
            ref.adam_update_for_flat_table(ug, rows, ptrs, tids, vd, ed, 0.05, 0.9, 0.999, 1e-8, 0.01, 3, D, True, 0)
        else:
            ref.rowwise_adagrad_for_flat_table(ug, rows, ptrs, tids, vd, ed, 0.05, 1e-8, D, True, 0)
        torch.cuda.synchronize()
        rec[f"opt_{nm}_after"] = table.cpu().numpy()[rows.cpu().numpy()]          # keep fixtures small: touched rows only
        rec[f"opt_{nm}_before"] = rec[f"opt_{nm}_before"][rows.cpu().numpy()]
    np.savez_compressed(os.path.join(OUT, "rows.npz"), **rec)
    report.append("rows: pooled/seq gather, reduce_grads, 4 optimizers")

    # ---------------- case C: segmented unique (order-free) + block bucketize ----------------
    n, T = 20000, 3
    keys = (rng.zipf(1.1, size=n) % 3000).astype(np.int64)
    trange = np.array([0, 5000, 5000, n], dtype=np.int64)
    nu_t, uk, rev, toffs, _ = ref.segmented_unique_cuda(torch.from_numpy(keys).to(dev), torch.from_numpy(trange).to(dev), T, None)
    k = int(nu_t.item())
    rec = {"keys": keys, "trange": trange, "num_unique": np.array(k), "unique_keys": uk[:k].cpu().numpy(), "reverse": rev.cpu().numpy(),
           "table_offsets": toffs.cpu().numpy()}
    Bb, Fb, W = 8, 3, 8
    lens = rng.integers(0, 100, size=Bb * Fb).astype(np.int64)
    ids = rng.integers(0, 1 << 50, size=int(lens.sum()), dtype=np.int64)
    ids[:40] = rng.integers(0, 4000, size=40)
    blk = np.array([1000, 1 << 47, 77], dtype=np.int64)
    rec.update({"bk_lengths": lens, "bk_ids": ids, "bk_blk": blk, "bk_B": np.array(Bb), "bk_W": np.array(W)})
    for tag, dts in [("cont", [0, 0, 0]), ("rr", [1, 1, 1]), ("hash", [2, 2, 2]), ("mixed", [0, 1, 2])]:
        r = ref.block_bucketize_sparse_features(torch.from_numpy(lens).to(dev), torch.from_numpy(ids).to(dev), False, True,
                                                torch.tensor(dts, dtype=torch.int32, device=dev), torch.from_numpy(blk).to(dev), W, None, None, Bb, None)
        rec[f"bk_{tag}_dts"] = np.array(dts)
        rec[f"bk_{tag}_new_lengths"] = r[0].cpu().numpy()
        rec[f"bk_{tag}_new_ids"] = r[1].cpu().numpy()
        rec[f"bk_{tag}_perm"] = r[4].cpu().numpy() if r[4] is not None else np.zeros(0)
    np.savez_compressed(os.path.join(OUT, "unique_bucketize.npz"), **rec)
    report.append("unique + bucketize")
    print("\n".join(report))
    print("GOLDEN_OK")


if __name__ == "__main__":
    try:
        main()
    except Exception as e:   # keep the rest of the gpurun command going
        import traceback
        traceback.print_exc()
        print("GOLDEN_FAILED", e)
        sys.exit(0)
