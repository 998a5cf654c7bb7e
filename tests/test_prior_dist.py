"""PriorDist (klerg.py:27-50) of the B200 mirror - closed form for the diagonal covariance, evaluated where the
samples live - against the oracle's MultivariateNormal restatement (the oracle's is checked against the live reference
class in test_oracle_live.py)."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "embodied-active-learning-vision_b200")]
from oracle import klerg_oracle as ko  # noqa: E402


@pytest.mark.parametrize("states", ["xy", "xyz", "xyzrpw", "xyXY", "xyb"])
def test_prior_dist_matches_oracle(states):
    from control_torch.klerg import PriorDist
    g = torch.Generator().manual_seed(len(states))
    x = torch.rand(500, len(states), generator=g) * 4 - 2
    if "r" in states:
        x[:, 3] += 3.0
    want = ko.OraclePrior(states).pdf_torch(x)
    got = PriorDist(states).pdf_torch(x)
    np.testing.assert_allclose(got.numpy(), want.numpy(), rtol=2e-6, atol=0)
    assert want.min() >= 1e-5
