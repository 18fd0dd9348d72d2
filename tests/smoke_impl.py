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
