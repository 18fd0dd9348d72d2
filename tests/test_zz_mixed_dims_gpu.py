"""Mixed embedding dims inside one pooled module on the kernels: the scenarios of tests/test_mixed_dims_cpu.py with the real op layer.

STATUS: written after this round's GPU budget was spent — NEVER run on hardware, hence opt-in (`RECSYS_B200_UNVERIFIED_GPU_TESTS=1`).
The kernels run exactly as in a uniform-dim module of width max_D (validated); what is new is the torch column slice / scatter around
them and the state-block addressing of the checkpoint, both covered by the CPU suite on the shim.
"""
import os

import pytest

from tests.test_mixed_dims_cpu import scenario_mixed_dims

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get("RECSYS_B200_UNVERIFIED_GPU_TESTS") != "1",
                                 reason="never run on hardware (GPU budget of the round was spent); set RECSYS_B200_UNVERIFIED_GPU_TESTS=1")]


@pytest.mark.parametrize("optimizer_name,mean", [("sgd", False), ("sgd", True), ("adagrad", False), ("adam", True)])
def test_mixed_dims_pooled(cuda, optimizer_name, mean, tmp_path):
    scenario_mixed_dims(cuda, optimizer_name, mean, tmp_path)
