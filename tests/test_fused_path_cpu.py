"""The DEFAULT module path (fused prefetch, the one bench.py and the GPU suite run) through its Python plumbing on the CPU shim:
`tests/cpu_ext_shim.py` restates `train_prefetch` / `table_update_counter_n` as the op sequence they fuse, so the module code that wraps
them — prefetch state with device-side counts, forward / fused backward, unpin, eval lookup, multi-table initializers — executes here.
Same check as the GPU test `test_fused_prefetch_matches_op_by_op`: the fused and the op-by-op module must stay identical step by step."""
import numpy as np
import pytest
import torch

from tests.cpu_ext_shim import patched_module


def _pair(pooling, T, opt):
    from dynamicemb import (BatchedDynamicEmbeddingTablesV2, DynamicEmbInitializerArgs, DynamicEmbInitializerMode, DynamicEmbScoreStrategy,
                            DynamicEmbTableOptions)
    mods = []
    for fused in (True, False):
        opts = [DynamicEmbTableOptions(dim=32, max_capacity=512, local_hbm_for_values=1 << 40, score_strategy=DynamicEmbScoreStrategy.STEP,
                                       initializer_args=DynamicEmbInitializerArgs(mode=DynamicEmbInitializerMode.DEBUG)) for _ in range(T)]
        mods.append(BatchedDynamicEmbeddingTablesV2(opts, table_names=[f"t{i}" for i in range(T)], feature_table_map=list(range(T)), pooling_mode=pooling,
                                                    optimizer=opt, learning_rate=0.05, fused_prefetch=fused))
    return mods


@pytest.mark.parametrize("pooling_name,T", [("sum", 1), ("mean", 2), ("none_nograd", 2)])
def test_fused_module_matches_op_by_op_cpu_shim(pooling_name, T):
    from dynamicemb import DynamicEmbPoolingMode as P, EmbOptimType
    pooling = {"sum": P.SUM, "mean": P.MEAN, "none_nograd": P.NONE}[pooling_name]
    rng = np.random.default_rng(11)
    with patched_module():
        ma, mb = _pair(pooling, T, EmbOptimType.EXACT_ADAGRAD)
        ma.train(), mb.train()
        B = 16
        for step in range(5):                                             # 512-row tables, key space 700: eviction starts after a few steps
            lens = rng.integers(0, 6, size=T * B)
            ids = torch.from_numpy(rng.integers(1, 700, size=int(lens.sum())).astype(np.int64))
            off = torch.from_numpy(np.concatenate([[0], np.cumsum(lens)]).astype(np.int64))
            if pooling == P.NONE:                                         # sequence mode with gradients prepares the backward on a CUDA side stream
                with torch.no_grad():
                    oa, ob = ma(ids, off), mb(ids, off)
                assert torch.equal(oa, ob), step
                continue
            oa, ob = ma(ids, off), mb(ids, off)
            assert torch.equal(oa, ob), step
            g = torch.from_numpy(rng.standard_normal(tuple(oa.shape)).astype(np.float32))
            oa.backward(g), ob.backward(g)
            assert torch.equal(ma.tables.table_storage_, mb.tables.table_storage_) and torch.equal(ma._values, mb._values), step
            assert int(ma.tables._ref_counter.sum()) == 0 and int(mb.tables._ref_counter.sum()) == 0, "pins released"
        assert ma.get_score() == mb.get_score()
        ma.eval(), mb.eval()
        q = torch.from_numpy(rng.integers(1, 900, size=T * B * 2).astype(np.int64))
        qo = torch.arange(0, T * B * 2 + 1, 2, dtype=torch.int64)
        assert torch.equal(ma(q, qo), mb(q, qo))
