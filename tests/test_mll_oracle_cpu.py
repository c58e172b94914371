"""pin the oracle restatement of get_marginal_likelihood (utils/training_evaluation.py:177-206)
against results recorded from the real reference (tests/golden/make_golden_mll.py).  CPU only."""
import os

import numpy as np
import pytest

from golden_util import GOLDEN_DIR, MLL_CASES, orc


@pytest.mark.parametrize("name", MLL_CASES)
def test_oracle_marginal_ll_matches_reference(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    ml64, rows = orc.marginal_ll_bernoulli(z["logits"], z["data"], clamp_abs=20.0, dtype=np.float64)
    ref = float(z["ml"])
    # the reference works in fp32 (sums of 784 terms, exp / log): 2e-6 relative on a value of O(10..1000)
    assert abs(ml64 - ref) <= 2e-6 * abs(ref) + 1e-6, (ml64, ref)
    ml32, _ = orc.marginal_ll_bernoulli(z["logits"], z["data"], clamp_abs=20.0, dtype=np.float32)
    assert abs(float(ml32) - ref) <= 5e-6 * abs(ref) + 1e-6
    assert rows.shape == (z["data"].shape[0],)


def test_oracle_marginal_ll_known_answer():
    """One sample, logits 0: every pixel costs log 2 whatever the data."""
    ml, rows = orc.marginal_ll_bernoulli(np.zeros((1, 10)), np.random.default_rng(0).random((5, 10)), clamp_abs=20.0)
    assert np.allclose(rows, -10 * np.log(2.0))
    assert abs(ml + 10 * np.log(2.0)) < 1e-12
    # clamp: logits beyond +-20 are cut (training_evaluation.py:180)
    a, _ = orc.marginal_ll_bernoulli(np.full((2, 3), 50.0), np.zeros((1, 3)), clamp_abs=20.0)
    b, _ = orc.marginal_ll_bernoulli(np.full((2, 3), 20.0), np.zeros((1, 3)), clamp_abs=20.0)
    assert a == b
