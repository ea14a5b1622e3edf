"""CPU restatement of the reference's Sinkhorn association.  TEST INFRASTRUCTURE ONLY.

log_optimal_transport: src/models/utils/track4d_utils.py:405-434; matching: Track4D.sinkhorn_module,
src/models/track4d.py:166-180.  Pinned by tests/golden/sinkhorn.npz, which oracle/gen_golden_assoc.py produced by
calling the reference's own functions."""
import torch


def log_sinkhorn_iterations(Z, log_mu, log_nu, iters):
    u, v = torch.zeros_like(log_mu), torch.zeros_like(log_nu)
    for _ in range(iters):
        u = log_mu - torch.logsumexp(Z + v.unsqueeze(1), dim=2)
        v = log_nu - torch.logsumexp(Z + u.unsqueeze(2), dim=1)
    return Z + u.unsqueeze(2) + v.unsqueeze(1)


def log_optimal_transport(scores, alpha, iters):
    b, m, n = scores.shape
    alpha = torch.as_tensor(alpha, dtype=scores.dtype)
    ms, ns = torch.tensor(float(m)), torch.tensor(float(n))
    couplings = torch.cat([torch.cat([scores, alpha.expand(b, m, 1)], -1),
                           torch.cat([alpha.expand(b, 1, n), alpha.expand(b, 1, 1)], -1)], 1)
    norm = -(ms + ns).log()
    log_mu = torch.cat([norm.expand(m), ns.log()[None] + norm])[None].expand(b, -1)
    log_nu = torch.cat([norm.expand(n), ms.log()[None] + norm])[None].expand(b, -1)
    return log_sinkhorn_iterations(couplings, log_mu, log_nu, iters) - norm


def sinkhorn_module(aff, alpha=0.9, iters=500):
    scores = log_optimal_transport(aff, alpha, iters)
    max0, max1 = scores[:, :-1, :-1].max(2), scores[:, :-1, :-1].max(1)
    indices0, indices1 = max0.indices, max1.indices
    ar0 = torch.arange(indices0.shape[1])[None]
    ar1 = torch.arange(indices1.shape[1])[None]
    mutual0 = ar0 == indices1.gather(1, indices0)
    mutual1 = ar1 == indices0.gather(1, indices1)
    mscores0 = torch.where(mutual0, max0.values.exp(), torch.zeros(()))
    valid0 = mutual0 & (mscores0 > 0)
    valid1 = mutual1 & valid0.gather(1, indices1)
    return torch.where(valid1, indices1, torch.full_like(indices1, -1)), scores


def dbscan_labels(x, eps=1.5, min_samples=2):
    """numpy restatement of sklearn.cluster.DBSCAN(eps, min_samples).fit_predict as the reference uses it
    (src/models/track4d.py:36,118): core points, components in order of their first core point, border points to the
    first cluster that reaches them, noise -1.  Pinned against sklearn itself in tests/test_association.py."""
    import numpy as np

    x = np.asarray(x, dtype=np.float64)
    n = x.shape[0]
    if n == 0:
        return np.zeros(0, dtype=np.int64)
    d2 = ((x[:, None, :] - x[None, :, :]) ** 2).sum(-1)
    adj = d2 <= float(eps) ** 2
    core = adj.sum(1) >= min_samples
    labels = np.full(n, -1, dtype=np.int64)
    cur = 0
    for i in range(n):
        if labels[i] != -1 or not core[i]:
            continue
        stack = [i]
        while stack:
            v = stack.pop()
            if labels[v] == -1:
                labels[v] = cur
                if core[v]:
                    stack.extend(int(j) for j in np.nonzero(adj[v] & (labels == -1))[0])
        cur += 1
    return labels
