"""Golden-vector generator, CPU side: executes the reference's OWN Python (unmodified source text, imported / exec'd from
/root/reference) in this container and commits inputs + outputs as fixtures.  Not run by the test suite (the GPU box has no
/root/reference); rerun by hand:  python tests/golden/gen_golden_cpu.py

  hash_vectors.npz : `murmur3_hash_64bits` (corelib/dynamicemb/dynamicemb/scored_hashtable.py:279-291)
  hstu_mask.npz    : `construct_mask` (third_party/FBGEMM/.../hstu/test/hstu_test.py:86-171).  Local-window cases use left,right > 0:
                     the helper maps a window bound of 0 to "unbounded" (:158-159) whereas the kernels treat 0 as a real bound
                     (hstu_blackwell/mask.py limit_right = row + 1 + window_size_right); the product follows the kernels.
  hstu_eager.npz   : `pytorch_hstu_mha` (examples/hstu/ops/pt_ops/pt_hstu_attention.py:150-196) with the two fbgemm jagged ops
                     it needs (jagged_to_padded_dense / dense_to_jagged — pure data movement) provided by a torch.library shim
"""
import ast
import math
import os
import sys

import numpy as np
import torch

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))


def extract(path, names, extra_globals=None):
    src = open(path).read()
    tree = ast.parse(src)
    g = {"torch": torch, "math": math, "np": np}
    g.update(extra_globals or {})
    for node in tree.body:
        if isinstance(node, (ast.FunctionDef,)) and node.name in names:
            code = compile(ast.Module(body=[node], type_ignores=[]), path, "exec")
            exec(code, g)
    return [g[n] for n in names]


def gen_hash():
    (h,) = extract(f"{REF}/corelib/dynamicemb/dynamicemb/scored_hashtable.py", ["murmur3_hash_64bits"])
    rng = np.random.default_rng(1)
    keys = np.concatenate([np.array([0, 1, 2, 12345, -1, -2, -3, -4, (1 << 63) - 1, -(1 << 63)], dtype=np.int64), rng.integers(-(1 << 63), (1 << 63) - 1, size=1000, dtype=np.int64)])
    out = np.array([h(int(k)) for k in keys.view(np.uint64)], dtype=np.uint64)
    np.savez_compressed(os.path.join(OUT, "hash_vectors.npz"), keys=keys, fmix64=out)
    print("hash_vectors", keys.size)


def gen_mask():
    (cm,) = extract(f"{REF}/third_party/FBGEMM/fbgemm_gpu/experimental/hstu/test/hstu_test.py", ["construct_mask"])
    rec = {}
    cases = []
    rng = np.random.default_rng(2)
    for ci, (B, seqlen, seqlen_c, seqlen_t, G, win) in enumerate([(3, 40, 0, 0, 1, (-1, 0)), (3, 40, 0, 9, 1, (-1, 0)), (3, 40, 0, 9, 3, (-1, 0)),
                                                                     (2, 33, 5, 8, 2, (-1, 0)), (2, 50, 0, 0, 1, (7, 2)), (2, 50, 0, 0, 1, (5, 3)), (2, 20, 0, 0, 1, (-1, -1))]):
        N = seqlen_c + seqlen + seqlen_t
        hist = rng.integers(1, seqlen + 1, size=B)
        nc = rng.integers(0, seqlen_c + 1, size=B) if seqlen_c else np.zeros(B, dtype=np.int64)
        nt = rng.integers(0, seqlen_t + 1, size=B) if seqlen_t else np.zeros(B, dtype=np.int64)
        lens = hist + nc + nt
        cu = torch.tensor(np.concatenate([[0], np.cumsum(lens)]), dtype=torch.int32)
        # construct_mask takes history length from cu_seqlens when seqused_k is None => pass cu of (context+history) lengths as the test does
        cu_hist = torch.tensor(np.concatenate([[0], np.cumsum(hist)]), dtype=torch.int32)
        m = cm(seqlen_c, seqlen, seqlen_t, G, win, None, cu_hist, cu_hist, None, None, torch.tensor(nc), torch.tensor(nt), torch.device("cpu"))
        rec[f"c{ci}_mask"] = m[:, 0].numpy()
        rec[f"c{ci}_hist"], rec[f"c{ci}_nc"], rec[f"c{ci}_nt"] = hist, nc, nt
        rec[f"c{ci}_meta"] = np.array([B, seqlen, seqlen_c, seqlen_t, G, win[0], win[1], N])
    rec["ncases"] = np.array(7)
    np.savez_compressed(os.path.join(OUT, "hstu_mask.npz"), **rec)
    print("hstu_mask ok")


def gen_eager():
    lib = torch.library.Library("fbgemm", "DEF")
    lib.define("jagged_to_padded_dense(Tensor values, Tensor[] offsets, int[] max_lengths, float padding_value=0.0) -> Tensor")
    lib.define("dense_to_jagged(Tensor dense, Tensor[] offsets, int? total_L=None) -> (Tensor, Tensor[])")

    def j2pd(values, offsets, max_lengths, padding_value=0.0):
        off = offsets[0].tolist(); N = max_lengths[0]
        out = values.new_full((len(off) - 1, N, values.shape[1]), padding_value)
        for b in range(len(off) - 1):
            n = min(off[b + 1] - off[b], N)
            out[b, :n] = values[off[b]: off[b] + n]
        return out

    def d2j(dense, offsets, total_L=None):
        off = offsets[0].tolist()
        return torch.cat([dense[b, : off[b + 1] - off[b]] for b in range(len(off) - 1)], dim=0), offsets

    lib.impl("jagged_to_padded_dense", j2pd, "CompositeExplicitAutograd")
    lib.impl("dense_to_jagged", d2j, "CompositeExplicitAutograd")
    import torch.nn.functional as F
    from typing import Optional, Union, Tuple
    fns = extract(f"{REF}/examples/hstu/ops/pt_ops/pt_hstu_attention.py", ["_get_valid_attn_mask", "_pad_qkv", "pytorch_hstu_mha"],
                  {"F": F, "Optional": Optional, "Union": Union, "Tuple": Tuple})
    g = fns[2].__globals__
    mha = fns[2]
    rec = {}
    torch.manual_seed(3)
    cases = [dict(lens=[37, 5, 64], H=2, D=32, nt=None, nc=None, G=1, scaling=-1), dict(lens=[50, 31], H=2, D=32, nt=[7, 0], nc=None, G=1, scaling=-1),
             dict(lens=[45, 60], H=1, D=64, nt=[10, 13], nc=None, G=4, scaling=200), dict(lens=[40, 22], H=2, D=32, nt=[6, 3], nc=[3, 2], G=2, scaling=-1)]
    for ci, c in enumerate(cases):
        lens = c["lens"]; T = sum(lens); N = max(lens)
        off = torch.tensor(np.concatenate([[0], np.cumsum(lens)]), dtype=torch.int64)
        q, k, v = (torch.randn(T, c["H"], c["D"]) for _ in range(3))
        alpha = 1.0 / math.sqrt(c["D"])
        out = mha(N, alpha, q, k, v, off, causal=True, dropout_pr=0.0, training=False,
                  num_targets=None if c["nt"] is None else torch.tensor(c["nt"]), num_contextuals=None if c["nc"] is None else torch.tensor(c["nc"]),
                  max_attn_len=None, target_group_size=c["G"], scaling_seqlen=c["scaling"])
        rec[f"c{ci}_q"], rec[f"c{ci}_k"], rec[f"c{ci}_v"], rec[f"c{ci}_out"] = q.numpy(), k.numpy(), v.numpy(), out.numpy()
        rec[f"c{ci}_lens"] = np.array(lens)
        rec[f"c{ci}_nt"] = np.array(c["nt"] if c["nt"] is not None else [-1])
        rec[f"c{ci}_nc"] = np.array(c["nc"] if c["nc"] is not None else [-1])
        rec[f"c{ci}_meta"] = np.array([c["H"], c["D"], c["G"], c["scaling"]])
    rec["ncases"] = np.array(len(cases))
    np.savez_compressed(os.path.join(OUT, "hstu_eager.npz"), **rec)
    print("hstu_eager ok")


if __name__ == "__main__":
    gen_hash()
    gen_mask()
    gen_eager()
