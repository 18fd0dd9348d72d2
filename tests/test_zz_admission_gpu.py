"""Admission on the kernels: the scenarios of tests/test_admission_cpu.py with the real op layer on a B200.

STATUS: written after this round's GPU budget was spent — these tests have NEVER run on hardware.  They are therefore opt-in
(`RECSYS_B200_UNVERIFIED_GPU_TESTS=1`), so that an unverified add-on cannot turn the validated GPU suite red; the same host code runs in
the CPU suite on the oracle-backed shim.  Every device step they reach is an op the validated GPU tests already cover (table insert
with ACCUMULATE, read-only lookup, erase, init_rows, gather_forward, backward with rows = -1, table export).
"""
import os

import pytest
import torch

from tests.test_admission_cpu import (scenario_counter_checkpoint, scenario_counter_dictionary, scenario_lfu, scenario_pooled_two_tables, scenario_sequence)

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get("RECSYS_B200_UNVERIFIED_GPU_TESTS") != "1",
                                 reason="never run on hardware (GPU budget of the round was spent); set RECSYS_B200_UNVERIFIED_GPU_TESTS=1")]


def test_admission_sequence(cuda):
    scenario_sequence(cuda)


@pytest.mark.parametrize("mean", [False, True])
def test_admission_pooled_two_tables(cuda, mean):
    scenario_pooled_two_tables(cuda, mean)


def test_admission_counter_checkpoint(cuda, tmp_path):
    scenario_counter_checkpoint(cuda, tmp_path)


def test_admission_lfu(cuda):
    scenario_lfu(cuda)


def test_counter_matches_dictionary(cuda):
    scenario_counter_dictionary(cuda)
