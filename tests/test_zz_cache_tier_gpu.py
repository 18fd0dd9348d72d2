"""Cache tier on the kernels: the scenarios of tests/test_cache_tier_cpu.py with the real op layer on a B200 (the backing table's value
rows in pinned host memory, read and written by the row-copy kernels).

STATUS: written after this round's GPU budget was spent — these tests have NEVER run on hardware, so they are opt-in
(`RECSYS_B200_UNVERIFIED_GPU_TESTS=1`) and cannot turn the validated GPU suite red; the same host code runs in the CPU suite on the
oracle-backed shim.  What only the GPU run can show: that the kernels dereference the pinned host rows (`dynamicemb_extensions.host_values`)
through unified addressing.
"""
import os

import pytest

from tests.test_cache_tier_cpu import (scenario_cache_pooled_two_tables, scenario_cache_train_evict_refetch, scenario_cache_with_admission,
                                       scenario_hybrid_pooled_checkpoint, scenario_hybrid_train)

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get("RECSYS_B200_UNVERIFIED_GPU_TESTS") != "1",
                                 reason="never run on hardware (GPU budget of the round was spent); set RECSYS_B200_UNVERIFIED_GPU_TESTS=1")]


@pytest.mark.parametrize("strategy", ["step", "lfu", "timestamp", "no_eviction"])
def test_cache_train_evict_refetch(cuda, strategy):
    from dynamicemb import DynamicEmbScoreStrategy as S
    scenario_cache_train_evict_refetch(cuda, {"step": S.STEP, "lfu": S.LFU, "timestamp": S.TIMESTAMP, "no_eviction": S.NO_EVICTION}[strategy],
                                       key_space=900 if strategy == "lfu" else 3000)


@pytest.mark.parametrize("mean", [False, True])
def test_cache_pooled_two_tables(cuda, mean):
    scenario_cache_pooled_two_tables(cuda, mean)


def test_cache_with_admission(cuda):
    scenario_cache_with_admission(cuda)


@pytest.mark.parametrize("strategy", ["step", "timestamp"])
def test_hybrid_train(cuda, strategy):
    from dynamicemb import DynamicEmbScoreStrategy as S
    scenario_hybrid_train(cuda, {"step": S.STEP, "timestamp": S.TIMESTAMP}[strategy])


def test_hybrid_pooled_checkpoint(cuda, tmp_path):
    scenario_hybrid_pooled_checkpoint(cuda, tmp_path)
