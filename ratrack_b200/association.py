"""Association step of Track4D on the device: the reference's `log_optimal_transport`
(src/models/utils/track4d_utils.py:405-434) and `Track4D.sinkhorn_module` (src/models/track4d.py:166-180) with their names
and signatures, served by ONE kernel (`rt_sinkhorn_match`, csrc/sinkhorn.cu) instead of ~2000 torch launches per frame
(500 iterations x two logsumexp + broadcasts).  No CPU / torch fallback."""
import torch

from . import _cabi


def _run(scores, alpha, iters, want_scores):
    if not scores.is_cuda:
        raise _cabi.RatrackError("sinkhorn: needs a CUDA tensor (no CPU / torch fallback)")
    aff = scores.detach().contiguous().float()
    b, m, n = aff.shape
    if m == 0 or n == 0:
        raise _cabi.RatrackError("sinkhorn: empty affinity matrix (the reference skips the association in that case, track4d.py:141)")
    out = torch.empty(b, m + 1, n + 1, device=aff.device) if want_scores else None
    i0 = torch.empty(b, m, dtype=torch.int64, device=aff.device)
    i1 = torch.empty(b, n, dtype=torch.int64, device=aff.device)
    with torch.cuda.device_of(aff):
        _cabi.call("rt_sinkhorn_match", b, m, n, aff.data_ptr(), float(alpha), int(iters), out.data_ptr() if want_scores else None,
                   i0.data_ptr(), i1.data_ptr(), torch.cuda.current_stream(aff.device).cuda_stream)
    return out, i0, i1


def log_optimal_transport(scores, alpha, iters):
    """(b,m,n) scores, dustbin score alpha, iteration count -> (b,m+1,n+1) log-couplings.  reference: track4d_utils.py:414-434.
    Inference-only here (the reference never back-propagates through it either: track4d.py:217 builds aff_mat from
    detached python floats)."""
    return _run(scores, alpha, iters, True)[0]


def sinkhorn_module(aff_mat_tensor, indices1=None, alpha=0.9, iters=500):
    """aff_mat (1,m,n) -> indices1 (1,n): index of the mutually-best previous object of every current object, -1 = none.
    reference: Track4D.sinkhorn_module, track4d.py:166-180 (alpha 0.9, 500 iterations)."""
    return _run(aff_mat_tensor, alpha, iters, False)[2]


def dbscan_labels(features, eps=1.5, min_samples=2):
    """(n,d) or (b,n,d) fp32 CUDA features -> int64 labels of the same leading shape, identical to
    sklearn.cluster.DBSCAN(eps, min_samples).fit_predict (the reference clusters on the host: Track4D.clustering,
    src/models/track4d.py:108-126, `self.dbscan = DBSCAN(eps=1.5, min_samples=args.min_obj_points)`, :36)."""
    if not features.is_cuda:
        raise _cabi.RatrackError("dbscan: needs a CUDA tensor (no CPU / sklearn fallback)")
    x = features.detach().contiguous().float()
    single = x.dim() == 2
    if single:
        x = x.unsqueeze(0)
    b, n, d = x.shape
    labels = torch.empty(b, n, dtype=torch.int32, device=x.device)
    with torch.cuda.device_of(x):
        _cabi.call("rt_dbscan", b, n, d, x.data_ptr(), float(eps), int(min_samples), labels.data_ptr(),
                   torch.cuda.current_stream(x.device).cuda_stream)
    labels = labels.long()
    return labels[0] if single else labels
