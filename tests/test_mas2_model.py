"""The index arithmetic of isp_mas2.cu's backtrack (tools/mas2_model.py: chain words -> transposed rows ->
bit-sliced group maps -> hops -> per-group expansion) against the C oracle, on the CPU."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
import mas2_model as mm
from oracle import mas as omas

CASES = [(1, 1), (1, 3), (5, 1), (3, 6), (10, 4), (33, 7), (64, 65), (97, 33), (130, 64), (75, 129), (200, 37)]


@pytest.mark.parametrize("n,m", CASES)
@pytest.mark.parametrize("quant", [0.0, 0.5])
def test_model_path_equals_oracle(n, m, quant):
    rs = np.random.RandomState(n * 1000 + m)
    x = rs.standard_normal((n, m)).astype(np.float32)
    if quant:
        x = (np.rint(x / quant) * quant).astype(np.float32)
    hard, _ = omas.b_mas_with_durations(x[None], [m], [n])
    ref = hard[0].argmax(1)
    assert np.array_equal(mm.mas_path(x), ref)


def test_model_all_equal_input():
    x = np.zeros((10, 4), np.float32)
    assert np.bincount(mm.mas_path(x), minlength=4).tolist() == [7, 1, 1, 1]
