"""Sinkhorn association: the oracle and the CUDA kernel against golden outputs of the reference's
own `log_optimal_transport` / `Track4D.sinkhorn_module` (tests/golden/sinkhorn.npz, generator oracle/gen_golden_assoc.py)."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle import association_oracle

G = np.load(os.path.join(GOLDEN, "sinkhorn.npz"))
CASES = [tuple(int(x) for x in c) for c in G["cases"]]


@pytest.mark.parametrize("m,n", CASES)
def test_oracle_matches_reference_sinkhorn(m, n):
    aff = torch.from_numpy(G[f"aff_{m}_{n}"])
    idx1, scores = association_oracle.sinkhorn_module(aff)
    assert np.array_equal(idx1.numpy(), G[f"idx1_{m}_{n}"])
    assert np.abs(scores.numpy() - G[f"scores_{m}_{n}"]).max() <= 1e-5


@pytest.mark.gpu
@pytest.mark.parametrize("m,n", CASES)
def test_kernel_matches_reference_sinkhorn(m, n):
    """indices bit-exact; log-couplings to 2e-4 abs (values span [-30, 1]; 500 iterations of fp32 exp/log on a different device)."""
    from ratrack_b200 import association

    aff = torch.from_numpy(G[f"aff_{m}_{n}"]).cuda()
    idx1 = association.sinkhorn_module(aff, None)
    scores = association.log_optimal_transport(aff, 0.9, 500)
    torch.cuda.synchronize()
    assert idx1.dtype == torch.int64 and tuple(idx1.shape) == (1, n)
    assert np.array_equal(idx1.cpu().numpy(), G[f"idx1_{m}_{n}"])
    assert np.abs(scores.cpu().numpy() - G[f"scores_{m}_{n}"]).max() <= 2e-4


@pytest.mark.gpu
def test_kernel_batched_and_limits():
    from ratrack_b200 import _cabi, association

    a = torch.from_numpy(np.concatenate([G["aff_20_20"]] * 3)).cuda()
    a[1] = a[1].flip(1)
    idx = association.sinkhorn_module(a, None).cpu().numpy()
    ref = G["idx1_20_20"][0]
    assert np.array_equal(idx[0], ref) and np.array_equal(idx[2], ref) and np.array_equal(idx[1], ref[::-1])
    with pytest.raises(_cabi.RatrackError):
        association.sinkhorn_module(torch.rand(1, 200, 3, device="cuda"), None)      # > 127 objects
    with pytest.raises(_cabi.RatrackError):
        association.sinkhorn_module(torch.rand(1, 0, 3, device="cuda"), None)        # empty: the reference skips association
