"""TEST INFRASTRUCTURE ONLY: a CPU stand-in for the native op layer (`dynamicemb.dynamicemb_extensions`), backed by the oracle.

The product has no CPU path — every op of `dynamicemb_extensions` launches a kernel of librecsys_b200.so.  The HOST logic above the op
layer (the module's op-by-op prefetch, admission, checkpoint plumbing) is plain Python over tensors, though, and this shim lets the CPU
suite run it end to end: each native op used by that logic is restated here on the oracle (`oracle/dynamicemb.py`, the C restatement of
the reference's table) or on torch CPU ops, with the SAME signatures and return conventions as the real op layer.  `patched_module()`
swaps it into the module / table namespaces for the duration of a test and makes the module allocate on the CPU.

It is only meaningful for small inputs (python loops).  The random initializer modes are deterministic per (seed, key, column) like the
product's Philox, but are not the kernels' numbers.
"""
import contextlib
import ctypes

import numpy as np
import torch

from oracle import dynamicemb as orc

_SENTINEL_MIN = np.uint64(0xFFFFFFFFFFFFFFFD)     # Locked / Reclaim / Empty (types.cuh:117-121)


def _np(t, dtype=None):
    if t is None:
        return None
    a = t.detach().contiguous().numpy()
    return a if dtype is None else np.ascontiguousarray(a, dtype=dtype)


def _table(storage, bkt_off, C, ns, bucket_sizes=None, counter=None):
    """An OracleTable VIEW over the tensors of a LinearBucketTable (the oracle mutates them in place)."""
    o = orc.OracleTable.__new__(orc.OracleTable)
    o.C, o.ns = int(C), int(ns)
    o.bkt_off = np.ascontiguousarray(bkt_off.numpy(), dtype=np.int64)
    o.num_buckets = int(o.bkt_off[-1])
    assert storage.is_contiguous()
    o.storage = storage.numpy()
    o.bucket_sizes = bucket_sizes.numpy() if bucket_sizes is not None else np.zeros(o.num_buckets, dtype=np.int32)
    o.counter = counter.numpy() if counter is not None else np.zeros(o.num_buckets * o.C, dtype=np.int32)
    return o


class CpuExt:
    """Attribute lookups fall through to the real op layer (enums, pure-torch helpers); native ops are overridden below."""

    def __init__(self, real):
        self._real = real
        self._clock = 1000

    def __getattr__(self, name):
        return getattr(self._real, name)

    # ------------------------------------------------------------------ table
    def device_timestamp(self):
        self._clock += 1
        return self._clock

    def table_init(self, table_storage, bucket_capacity, num_scores=1):
        nb = table_storage.numel() // ((9 + 8 * num_scores) * bucket_capacity)
        orc.lib().orc_table_init(ctypes.c_void_p(table_storage.data_ptr()), ctypes.c_int64(nb), ctypes.c_int64(bucket_capacity), ctypes.c_int64(num_scores))

    def fill_i32(self, t, v):
        t.fill_(v)

    def table_lookup(self, table_storage, table_bucket_offsets, bucket_capacity, keys, table_ids, score_input, policy_type, num_scores=1, timestamp=0):
        o = _table(table_storage, table_bucket_offsets, bucket_capacity, num_scores)
        so, f, idx = o.lookup(_np(keys), _np(table_ids, np.int64), policy=int(policy_type), score_in=_np(score_input), timer=int(timestamp))
        return torch.from_numpy(so), torch.from_numpy(f), torch.from_numpy(idx)

    def table_insert(self, table_storage, table_bucket_offsets, bucket_capacity, bucket_sizes, keys, table_ids, score_input, policy_type, counter,
                     insert_results=None, score_output=None, num_scores=1, timestamp=0):
        if keys.numel() == 0:
            return torch.empty(0, dtype=torch.int64)
        o = _table(table_storage, table_bucket_offsets, bucket_capacity, num_scores, bucket_sizes, counter)
        idx, res, so, _ = o.insert(_np(keys), _np(table_ids, np.int64), policy=int(policy_type), score_in=_np(score_input), timer=int(timestamp),
                                   use_counter=counter is not None)
        if insert_results is not None:
            insert_results.copy_(torch.from_numpy(res))
        if score_output is not None:
            score_output.copy_(torch.from_numpy(so))
        return torch.from_numpy(idx)

    def table_insert_and_evict(self, table_storage, table_bucket_offsets, bucket_capacity, bucket_sizes, keys, table_ids, score_input, policy_type,
                               counter, insert_results=None, score_output=None, num_scores=1, timestamp=0):
        n = keys.numel()
        z = lambda: torch.zeros(n, dtype=torch.int64)      # noqa: E731
        if n == 0:
            return torch.empty(0, dtype=torch.int64), torch.zeros(1, dtype=torch.int64), torch.zeros(0, dtype=keys.dtype), z(), z(), z()
        o = _table(table_storage, table_bucket_offsets, bucket_capacity, num_scores, bucket_sizes, counter)
        idx, res, so, (ek, ei, es, et) = o.insert(_np(keys), _np(table_ids, np.int64), policy=int(policy_type), score_in=_np(score_input),
                                                  timer=int(timestamp), use_counter=counter is not None)
        if insert_results is not None:
            insert_results.copy_(torch.from_numpy(res))
        if score_output is not None:
            score_output.copy_(torch.from_numpy(so))
        outk, outi, outs, outt = torch.zeros(n, dtype=keys.dtype), z(), z(), z()
        m = ek.size
        outk[:m] = torch.from_numpy(ek.view(np.int64).copy())
        outi[:m], outs[:m], outt[:m] = torch.from_numpy(ei.copy()), torch.from_numpy(es.copy()), torch.from_numpy(et.copy())
        return torch.from_numpy(idx), torch.tensor([m], dtype=torch.int64), outk, outi, outs, outt

    def host_values(self, rows, width):
        return torch.zeros(rows, width, dtype=torch.float32)

    def table_erase(self, table_storage, table_bucket_offsets, bucket_capacity, bucket_sizes, keys, table_ids, indices=None, num_scores=1):
        if keys.numel():
            _table(table_storage, table_bucket_offsets, bucket_capacity, num_scores, bucket_sizes).erase(_np(keys), _np(table_ids, np.int64))

    def table_update_counter_with_layout(self, counter, slot_indices, delta, table_bucket_offsets, bucket_capacity, total_capacity=None,
                                         num_tables=None, table_ids=None, overflow_output_offsets=None, overflow_bucket_capacity=0):
        s = slot_indices.to(torch.int64)
        ok = s >= 0
        base = table_bucket_offsets[:-1] * bucket_capacity
        g = s + (base[table_ids.to(torch.int64)] if table_ids is not None else 0)
        counter.index_add_(0, g[ok], torch.full((int(ok.sum()),), int(delta), dtype=counter.dtype))

    def table_export_batch(self, table_storage, bucket_capacity, batch, offset, key_dtype, threshold=None, table_begin=0, num_scores=1, score_index=0):
        C = bucket_capacity
        nb = table_storage.numel() // ((9 + 8 * num_scores) * C)
        o = _table(table_storage, torch.tensor([0, nb]), C, num_scores)
        keys = o.keys_view().reshape(-1)[offset:offset + batch]
        scores = o.scores_view()[:, :, score_index].reshape(-1)[offset:offset + batch]
        live = keys < _SENTINEL_MIN
        if threshold is not None:
            live &= scores >= np.uint64(threshold)
        sel = np.nonzero(live)[0]
        kout = torch.zeros(batch, dtype=key_dtype)
        sout = torch.zeros(batch, dtype=torch.int64)
        iout = torch.zeros(batch, dtype=torch.int64)
        kout[:sel.size] = torch.from_numpy(keys[sel].copy().view(np.int64)).view(key_dtype) if key_dtype != torch.int64 else torch.from_numpy(keys[sel].copy().view(np.int64))
        sout[:sel.size] = torch.from_numpy(scores[sel].copy().view(np.int64))
        iout[:sel.size] = torch.from_numpy(sel.astype(np.int64) + offset - table_begin)
        return torch.tensor([sel.size], dtype=torch.int64), kout, sout, iout

    # ------------------------------------------------------------------ fused training prefetch (one C call in the product)
    def unique_scratch(self, n_max, num_tables, device):
        return torch.empty(0, dtype=torch.uint8)

    def table_update_counter_n(self, counter, slot_indices, delta, table_bucket_offsets, bucket_capacity, n_device, table_ids=None):
        n = int(n_device.item())
        self.table_update_counter_with_layout(counter, slot_indices[:n], delta, table_bucket_offsets, bucket_capacity,
                                              table_ids=table_ids[:n] if table_ids is not None else None)

    def train_prefetch(self, table_storage, table_bucket_offsets, bucket_capacity, bucket_sizes, ref_counter, bucket_heads, values, emb_dim, row_base,
                       keys, table_range, num_tables, policy, table_scores, timestamp, init_mode, init_params, seed, state_init,
                       freq_in=None, num_scores=1, table_init=None, n_dev=None, unique_scratch=None):
        """unique -> lookup (hits scored + pinned) -> insert the misses -> init their rows -> pin; outputs padded to n with -1 / 0."""
        n = keys.numel()
        n_real = int(n_dev.item()) if n_dev is not None else n
        want_freq = int(policy) in (2, 4)
        fin = (freq_in if freq_in is not None else torch.empty(0, dtype=torch.int64)) if want_freq else None
        num_u, uk, rev, _, freq, utids = self.segmented_unique_cuda(keys[:n_real], table_range, num_tables, fin, want_table_ids=True)
        nu = int(num_u.item())
        k, t = uk[:nu].contiguous(), utids[:nu].contiguous()
        if int(policy) == 1:
            score = table_scores[t]
        elif want_freq:
            score = freq[:nu]
        else:
            score = None
        args = (table_storage, table_bucket_offsets, bucket_capacity)
        _, founds, slots = self.table_lookup(*args, k, t, score, policy, num_scores=num_scores, timestamp=timestamp)
        self.table_update_counter_with_layout(ref_counter, slots, 1, table_bucket_offsets, bucket_capacity, table_ids=t)
        miss = (~founds).nonzero(as_tuple=True)[0]
        if miss.numel():
            mk, mt = k[miss].contiguous(), t[miss].contiguous()
            new = self.table_insert(*args, bucket_sizes, mk, mt, score[miss].contiguous() if score is not None else None, policy, ref_counter,
                                    num_scores=num_scores, timestamp=timestamp)
            p0, p1, p2, p3 = init_params
            self.init_rows(values, emb_dim, self.rows_from_slots(new, mt, row_base), mk, init_mode, p0, p1, p2, p3, seed=seed, state_init=state_init,
                           table_ids=mt if table_init is not None else None, table_init=table_init)
            self.table_update_counter_with_layout(ref_counter, new, 1, table_bucket_offsets, bucket_capacity, table_ids=mt)
            slots[miss] = new
        pad = lambda x, fill: torch.cat([x, torch.full((n - x.numel(),), fill, dtype=x.dtype)])      # noqa: E731
        rows = self.rows_from_slots(slots, t, row_base)
        rev_full = torch.zeros(n, dtype=torch.int64)
        rev_full[:n_real] = rev[:n_real]
        return pad(k, 0), rev_full, pad(t, 0), pad(slots, -1), pad(rows, -1), torch.tensor([nu], dtype=torch.int64)

    # ------------------------------------------------------------------ dedup
    def get_table_range(self, offsets, feature_offsets, num_features=None):
        if num_features is None:
            num_features = int(feature_offsets[-1])
        B = (offsets.numel() - 1) // num_features if num_features > 0 else 0
        return offsets.to(torch.int64)[feature_offsets.to(torch.int64) * B].contiguous()

    def segmented_unique_cuda(self, keys, segment_range, num_tables, input_frequencies=None, want_table_ids=False, n_dev=None, scratch=None):
        n = keys.numel()
        rng = _np(segment_range, np.int64) if (segment_range is not None and num_tables > 1) else np.array([0, n], dtype=np.int64)
        uk, inv, offs = orc.segmented_unique(_np(keys), rng)
        nu = uk.size
        ukeys = torch.zeros(n, dtype=keys.dtype)
        ukeys[:nu] = torch.from_numpy(uk)
        freq = None
        if input_frequencies is not None:
            w = _np(input_frequencies, np.int64) if input_frequencies.numel() == n and n > 0 else np.ones(n, dtype=np.int64)
            f = np.zeros(n, dtype=np.int64)
            np.add.at(f, inv, w)
            freq = torch.from_numpy(f)
        utids = torch.zeros(n, dtype=torch.int64)
        for t in range(len(offs) - 1):
            utids[int(offs[t]):int(offs[t + 1])] = t
        out = (torch.tensor([nu], dtype=torch.int64), ukeys, torch.from_numpy(inv), torch.from_numpy(offs), freq)
        return out + (utids,) if want_table_ids else out

    # ------------------------------------------------------------------ rows
    def rows_from_slots(self, slots, table_ids, row_base):
        base = row_base[table_ids.to(torch.int64)] if (row_base is not None and table_ids is not None) else (row_base[0] if row_base is not None else 0)
        return torch.where(slots >= 0, slots + base, torch.full_like(slots, -1))

    def init_rows(self, values, emb_dim, rows, keys, mode, p0=0.0, p1=1.0, p2=0.0, p3=0.0, seed=0, state_init=0.0, only_if=None, emb_out=None,
                  table_ids=None, table_init=None):
        M = self._real.InitializerMode
        nk = keys.numel()
        modes, params, seeds = [int(mode)] * nk, [(float(p0), float(p1), float(p2), float(p3))] * nk, [int(seed)] * nk
        if table_init is not None:
            img = table_init.numpy().view(np.dtype([("mode", "<i4"), ("p", "<f4", (4,)), ("reserved", "<u4"), ("seed", "<u8")]))
            tid = table_ids.tolist() if table_ids is not None else [0] * nk
            modes, params = [int(img["mode"][t]) for t in tid], [tuple(float(x) for x in img["p"][t]) for t in tid]
            seeds = [int(img["seed"][t]) for t in tid]
        ku = _np(keys).view(np.uint64)
        for i in range(nk):
            if only_if is not None and not bool(only_if[i]):
                continue
            a, b, lo, up = params[i]
            if modes[i] == int(M.DEBUG):
                v = float(int(ku[i]) % 100000)
            elif modes[i] == int(M.CONSTANT):
                v = a
            else:
                # random modes: like the product's Philox, a function of (seed, key, column) only — NOT the same numbers as the kernels
                g = np.random.default_rng([seeds[i] & 0xFFFFFFFF, seeds[i] >> 32, int(ku[i]) & 0xFFFFFFFF, int(ku[i]) >> 32])
                if modes[i] == int(M.UNIFORM):
                    v = torch.from_numpy(g.uniform(a, b, emb_dim).astype(np.float32))
                else:
                    z = g.standard_normal(emb_dim) * b + a
                    v = torch.from_numpy((np.clip(z, lo, up) if modes[i] == int(M.TRUNCATED_NORMAL) else z).astype(np.float32))
            r = int(rows[i]) if rows is not None else -1
            if r >= 0:
                values[r, :emb_dim] = v
                values[r, emb_dim:] = state_init
            if emb_out is not None:
                emb_out[i, :emb_dim] = v

    def lookup_forward(self, table_storage, table_bucket_offsets, bucket_capacity, values, emb_dim, keys, *, row_base=None, table_range=None,
                       num_tables=1, offsets=None, batch_size=0, num_features=0, combiner=-1, out_dtype=torch.float32, absent_value=0.0,
                       want_founds=False, num_scores=1):
        """Read-only probe + gather (the fused eval kernel): ids in table-major order, absent ids read `absent_value`."""
        n = keys.numel()
        if table_range is not None and num_tables > 1:
            tr = table_range.to(torch.int64)
            tids = torch.repeat_interleave(torch.arange(num_tables), tr[1:] - tr[:-1])
        else:
            tids = torch.zeros(n, dtype=torch.int64)
        _, founds, slots = self.table_lookup(table_storage, table_bucket_offsets, bucket_capacity, keys, tids, None, 0, num_scores=num_scores)
        rows = self.rows_from_slots(slots, tids, row_base)
        kw = dict(offsets=offsets, batch_size=batch_size, num_features=num_features, combiner=combiner, out_dtype=torch.float32)
        out = self.gather_forward(values, emb_dim, rows, None, n, **kw)
        if absent_value != 0.0:
            fill = torch.full((1, values.shape[1]), float(absent_value))
            out = out + self.gather_forward(fill, emb_dim, torch.where(founds, torch.full_like(slots, -1), torch.zeros_like(slots)), None, n, **kw)
        out = out.to(out_dtype)
        return (out, founds, slots) if want_founds else out

    def copy_rows(self, values, width, rows, dense, to_table):
        ok = rows >= 0
        if to_table:
            values[rows[ok], :width] = dense[ok, :width]
        else:
            dense[:, :width] = 0
            dense[ok, :width] = values[rows[ok], :width]

    def gather_forward(self, values, emb_dim, rows, inverse, n, *, offsets=None, batch_size=0, num_features=0, combiner=-1, out_dtype=torch.float32,
                       n_dev=None, out=None):
        per_id = rows[inverse] if inverse is not None else rows
        if combiner < 0:
            res = torch.from_numpy(orc.gather_rows(values.numpy(), emb_dim, _np(per_id, np.int64))) if n else torch.zeros(0, emb_dim)
        else:
            res = torch.from_numpy(orc.pool_rows(values.numpy(), emb_dim, _np(offsets, np.int64), _np(per_id, np.int64), combiner, batch_size, num_features))
        res = res.to(out_dtype)
        if out is not None:
            out.copy_(res)
            return out
        return res

    def backward(self, values, emb_dim, inverse, num_unique_bound, rows, grads, *, offsets=None, batch_size=0, num_features=0, combiner=-1,
                 opt_type=0, lr=0.0, eps=1e-8, beta1=0.9, beta2=0.999, weight_decay=0.0, bc1=1.0, bc2=1.0, want_unique_grads=False, prepared=None,
                 n_dev=None, grad_row_of=None, unique_grad_addr=None, grad_stride=None):
        """Per-unique gradient sums (torch index_add in fp32; the GPU kernel's summation order is covered by the GPU tests) + the
        oracle's optimizer update on rows >= 0."""
        n = inverse.numel()
        D = emb_dim
        g = grads.to(torch.float32)
        if combiner >= 0:       # pooled: every id of bag (f, b) receives grads[b, f*D:(f+1)*D] (scaled by 1/len for MEAN)
            off = offsets.to(torch.int64)
            lens = off[1:] - off[:-1]
            bag = torch.repeat_interleave(torch.arange(lens.numel()), lens)
            f, b = bag // batch_size, bag % batch_size
            per_id = g.view(batch_size, num_features, D)[b, f]
            if combiner == 1:
                per_id = per_id / lens[bag].clamp(min=1).to(torch.float32).unsqueeze(1)
        else:
            per_id = g.view(n, D)
        ug = torch.zeros(num_unique_bound, D, dtype=torch.float32)
        ug.index_add_(0, inverse, per_id)
        if values is not None and int(opt_type) != 0:
            ok = (rows[:num_unique_bound] >= 0).nonzero(as_tuple=True)[0]
            name = {1: "sgd", 2: "adam", 3: "adagrad", 4: "rowwise_adagrad"}[int(opt_type)]
            step = 1
            if int(opt_type) == 2:      # bc1 = 1 - beta1^step
                import math
                step = max(1, round(math.log(max(1.0 - bc1, 1e-300)) / math.log(beta1)))
            orc.optimizer_update(values.numpy(), D, _np(rows[ok], np.int64), ug[ok].numpy(), name, lr, eps=eps, beta1=beta1, beta2=beta2,
                                 weight_decay=weight_decay, step=step)
        return ug if want_unique_grads else None


    class BackwardPrep:
        """The product's side stream for the gradient-independent half of the backward; nothing to prepare on the CPU."""

        def __init__(self, device=None):
            pass

    def backward_prepare(self, prep, emb_dim, inverse, num_unique_bound, n_dev=None, grad_row_of=None):
        return None

    def reduce_grads(self, reverse_indices, grads, num_unique, batch_size, out_dim, offsets=None, D_offsets=None, combiner=-1, total_D=0):
        F = (total_D // out_dim) if offsets is not None else 0
        return self.backward(None, out_dim, reverse_indices, num_unique, None, grads, offsets=offsets, batch_size=batch_size if offsets is not None else 0,
                             num_features=F, combiner=combiner if offsets is not None else -1, want_unique_grads=True)

    # ------------------------------------------------------------------ row-wise input dist
    def block_bucketize_sparse_features(self, lengths, indices, batch_size, world_size, block_sizes, dist_type_per_feature=None, weights=None,
                                        sequence=True):
        dts = dist_type_per_feature.tolist() if dist_type_per_feature is not None else [0] * (lengths.numel() // max(batch_size, 1))
        nl, ni, perm = orc.block_bucketize(_np(lengths, np.int64), _np(indices, np.int64), batch_size, world_size, _np(block_sizes, np.int64), dts)
        return torch.from_numpy(nl), torch.from_numpy(ni), None, (torch.from_numpy(perm) if sequence else None)


class _CudaStub:
    def current_device(self):
        return 0

    def is_current_stream_capturing(self):
        return False

    def current_stream(self, *a, **k):
        class _S:
            def synchronize(self):
                pass
        return _S()


class _TorchProxy:
    """`torch` as the module sees it during a CPU test: every device is the CPU, `torch.cuda` is a stub."""

    def __init__(self):
        self.cuda = _CudaStub()

    def __getattr__(self, name):
        return getattr(torch, name)

    def device(self, *a, **k):
        return torch.device("cpu")


@contextlib.contextmanager
def patched_module():
    """Swap the CPU shim into the module, table and admission namespaces; yields the shim."""
    import dynamicemb.batched_dynamicemb_tables as btm
    import dynamicemb.scored_hashtable as sht
    import dynamicemb.shard as shd
    real = btm.ext
    shim = CpuExt(real)
    saved = (btm.ext, sht.ext, btm.torch, shd.ext)
    btm.ext, sht.ext, btm.torch, shd.ext = shim, shim, _TorchProxy(), shim
    try:
        yield shim
    finally:
        btm.ext, sht.ext, btm.torch, shd.ext = saved
