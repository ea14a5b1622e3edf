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
        association.sinkhorn_module(torch.rand(1, 0, 3, device="cuda"), None)        # empty: the reference skips association


@pytest.mark.gpu
@pytest.mark.parametrize("m,n", [(150, 140), (200, 3), (5, 300)])
def test_kernel_beyond_127_objects(m, n):
    """The reference has no size limit (track4d.py:166-180): beyond 127 objects on a side the coupling matrix moves from
    shared memory to a stream-ordered global scratch, same kernel.  Checked against the CPU oracle."""
    from ratrack_b200 import association

    g = torch.Generator().manual_seed(m * 1000 + n)
    aff = torch.rand(2, m, n, generator=g)
    want_idx, want_scores = association_oracle.sinkhorn_module(aff)
    got_idx = association.sinkhorn_module(aff.cuda(), None)
    got_scores = association.log_optimal_transport(aff.cuda(), 0.9, 500)
    torch.cuda.synchronize()
    assert np.abs(got_scores.cpu().numpy() - want_scores.numpy()).max() <= 5e-4
    # uniform random affinities produce near-ties between couplings; an index may differ only where the two candidates'
    # scores agree to within the fp32 noise of 500 exp/log iterations
    gi, wi = got_idx.cpu().numpy(), want_idx.numpy()
    sc = want_scores.numpy()[:, :-1, :-1]
    for b, j in zip(*np.nonzero(gi != wi)):
        col = sc[b, :, j]
        cand = [col[i] for i in (gi[b, j], wi[b, j]) if i >= 0]
        assert cand and col.max() - min(cand) <= 1e-3, (b, j, gi[b, j], wi[b, j])


def _cluster_sets():
    """(name, X (n,8) float32, min_samples): blobs + background, a long chain (large graph diameter), duplicates, tiny sets"""
    rng = np.random.default_rng(7)
    sets = []
    for n, k, ms in ((300, 12, 2), (1024, 40, 2), (700, 25, 3), (512, 10, 5), (37, 3, 1)):
        centres = rng.uniform(-40, 40, (k, 8))
        pts = np.concatenate([centres[rng.integers(0, k, n // 2)] + rng.normal(0, 0.45, (n // 2, 8)),
                              rng.uniform(-45, 45, (n - n // 2, 8))])
        sets.append((f"blobs{n}", pts[rng.permutation(n)].astype(np.float32), ms))
    chain = np.zeros((400, 8), np.float32)
    chain[:, 0] = np.arange(400) * 1.4                       # every link within eps = 1.5: one component of diameter 399
    sets.append(("chain", chain[rng.permutation(400)], 2))
    dup = rng.normal(0, 3.0, (64, 8)).astype(np.float32)
    dup[10:20] = dup[0]
    sets.append(("duplicates", dup, 2))
    sets.append(("single", np.zeros((1, 8), np.float32), 2))
    sets.append(("pair", np.array([[0] * 8, [1] + [0] * 7], np.float32), 2))
    return sets


@pytest.mark.parametrize("name,x,ms", _cluster_sets(), ids=[s[0] for s in _cluster_sets()])
def test_dbscan_oracle_matches_sklearn(name, x, ms):
    from sklearn.cluster import DBSCAN   # the reference's own clustering call (track4d.py:36,118)

    want = DBSCAN(eps=1.5, min_samples=ms).fit_predict(x)
    assert np.array_equal(association_oracle.dbscan_labels(x, 1.5, ms), want)


@pytest.mark.gpu
@pytest.mark.parametrize("name,x,ms", _cluster_sets(), ids=[s[0] for s in _cluster_sets()])
def test_dbscan_kernel_matches_sklearn(name, x, ms):
    from sklearn.cluster import DBSCAN

    from ratrack_b200 import association

    want = DBSCAN(eps=1.5, min_samples=ms).fit_predict(x)
    got = association.dbscan_labels(torch.from_numpy(x).cuda(), 1.5, ms)
    torch.cuda.synchronize()
    assert got.dtype == torch.int64 and np.array_equal(got.cpu().numpy(), want)


@pytest.mark.gpu
def test_dbscan_kernel_batched_and_limits():
    from ratrack_b200 import _cabi, association

    sets = [s for s in _cluster_sets() if s[0] == "blobs300"]
    x = torch.from_numpy(np.stack([sets[0][1], sets[0][1][::-1].copy()])).cuda()
    lab = association.dbscan_labels(x, 1.5, 2).cpu().numpy()
    ref = [association_oracle.dbscan_labels(x[i].cpu().numpy(), 1.5, 2) for i in range(2)]
    assert np.array_equal(lab[0], ref[0]) and np.array_equal(lab[1], ref[1])
    with pytest.raises(_cabi.RatrackError):
        association.dbscan_labels(torch.zeros(20000, 8, device="cuda"))             # > 16384 points per set


@pytest.mark.gpu
@pytest.mark.parametrize("n,ms", [(1025, 2), (3000, 2), (2500, 3)])
def test_dbscan_kernel_beyond_1024_points(n, ms):
    """config-5 geometry (N ~ 3000) can leave more than 1024 moving points: the adjacency then lives in a stream-ordered
    global scratch instead of shared memory.  Labels still identical to sklearn."""
    from sklearn.cluster import DBSCAN

    from ratrack_b200 import association

    rng = np.random.default_rng(n)
    k = n // 25
    centres = rng.uniform(-60, 60, (k, 8))
    pts = np.concatenate([centres[rng.integers(0, k, n // 2)] + rng.normal(0, 0.45, (n // 2, 8)),
                          rng.uniform(-65, 65, (n - n // 2, 8))])
    x = pts[rng.permutation(n)].astype(np.float32)
    want = DBSCAN(eps=1.5, min_samples=ms).fit_predict(x)
    got = association.dbscan_labels(torch.from_numpy(np.stack([x, x[::-1].copy()])).cuda(), 1.5, ms)
    torch.cuda.synchronize()
    assert np.array_equal(got[0].cpu().numpy(), want)
    assert np.array_equal(got[1].cpu().numpy(), DBSCAN(eps=1.5, min_samples=ms).fit_predict(x[::-1].copy()))
