"""SURVEY §8(f) N1 on the GPU: mcpc_marginal_ll_bernoulli (tcgen05 GEMM + streaming min / sum-exp epilogue) against
the oracle, through the C ABI (montecarlopredictivecoding_b200.mcpc_utils.bernoulli_marginal_ll).

Tolerance: the operands are split into bf16 hi + lo (three products accumulate in fp32), which leaves ~2^-16
relative error per product; on per-row log-likelihoods of magnitude 10..1000 that is far below 1e-5 relative.
The tests allow 2e-5 * |value| + 2e-4."""
import os

import numpy as np
import pytest
import torch

from golden_util import GOLDEN_DIR, MLL_CASES, orc
from montecarlopredictivecoding_b200 import mcpc_utils as mu

pytestmark = pytest.mark.gpu


def _check(logits, data, clamp=20.0):
    dev = torch.device("cuda:0")
    ml, rows = mu.bernoulli_marginal_ll(torch.from_numpy(logits).to(dev), torch.from_numpy(data).to(dev),
                                        clamp_abs=clamp, return_rows=True)
    ml_ref, rows_ref = orc.marginal_ll_bernoulli(logits, data, clamp_abs=clamp, dtype=np.float64)
    rows = rows.cpu().numpy().astype(np.float64)
    err = np.abs(rows - rows_ref)
    assert np.all(err <= 2e-5 * np.abs(rows_ref) + 2e-4), (float(err.max()), float(np.abs(rows_ref).max()))
    assert abs(float(ml) - ml_ref) <= 2e-5 * abs(ml_ref) + 2e-4
    return float(ml), ml_ref


@pytest.mark.parametrize("name", MLL_CASES)
def test_marginal_ll_golden(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    ml, _ = _check(z["logits"], z["data"])
    ref = float(z["ml"])                                  # what the real reference returned
    assert abs(ml - ref) <= 2e-5 * abs(ref) + 2e-4


@pytest.mark.parametrize("N,S,D", [(1, 1, 1), (5, 3, 7), (130, 257, 64), (300, 700, 784), (129, 513, 100)])
def test_marginal_ll_shapes(N, S, D):
    """Ragged sizes: partial row tiles, partial sample tiles (padding must contribute exactly 0), K padding."""
    rng = np.random.default_rng(N * 1000 + S)
    logits = (rng.standard_normal((S, D)) * 6.0).astype(np.float32)          # some beyond the clamp
    logits[rng.random((S, D)) < 0.01] = 35.0
    data = rng.random((N, D)).astype(np.float32)
    data[rng.random((N, D)) < 0.5] = 0.0                                      # MNIST-like: mostly exact zeros
    _check(logits, data)


def test_marginal_ll_properties_full_size():
    """table_1.py sizes (5000 samples x 784) on a 2,048-row slice: permutation invariance over samples and rows,
    and duplicating the sample set leaves every row unchanged (mean over samples)."""
    dev = torch.device("cuda:0")
    g = torch.Generator(device="cpu").manual_seed(3)
    logits = (torch.randn(5000, 784, generator=g) * 4.0).to(dev)
    data = (torch.rand(2048, 784, generator=g) < 0.2).float().to(dev)
    ml, rows = mu.bernoulli_marginal_ll(logits, data, return_rows=True)
    perm_s = torch.randperm(5000, generator=g).to(dev)
    perm_r = torch.randperm(2048, generator=g).to(dev)
    ml2, rows2 = mu.bernoulli_marginal_ll(logits[perm_s], data[perm_r], return_rows=True)
    assert torch.allclose(rows[perm_r], rows2, rtol=1e-5, atol=1e-3)
    assert abs(float(ml) - float(ml2)) <= 1e-5 * abs(float(ml))
    ml3, rows3 = mu.bernoulli_marginal_ll(torch.cat([logits, logits]), data, return_rows=True)
    assert torch.allclose(rows, rows3, rtol=1e-5, atol=1e-3)
    # spot-check 16 rows against the oracle
    idx = np.arange(0, 2048, 128)
    _, ref = orc.marginal_ll_bernoulli(logits.cpu().numpy(), data[idx].cpu().numpy(), dtype=np.float64)
    got = rows.cpu().numpy()[idx]
    assert np.all(np.abs(got - ref) <= 2e-5 * np.abs(ref) + 2e-4)


def test_marginal_ll_requires_cuda():
    with pytest.raises(RuntimeError):
        mu.bernoulli_marginal_ll(torch.zeros(2, 3), torch.zeros(2, 3))
