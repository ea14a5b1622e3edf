"""`Track4D` with the reference's constructor, forward signature, return tuple and state_dict keys
(reference: src/models/track4d.py:13-245), every device-side step served by this package:

    backbone            -> the fused engine / modular path (model_utils.Track4DBackbone)
    clustering          -> rt_dbscan on the device (the reference: device->host copy + scikit-learn, :108-126)
    affinity_module     -> ONE batched evaluation of the Affinity MLP over all m x n object pairs, object embeddings by
                           segment reductions (the reference: a Python double loop, ~15 torch launches per pair, :182-223)
    sinkhorn_module     -> rt_sinkhorn_match (500 log-space iterations + mutual matching in one kernel, :166-180)

What stays on the host is what the reference's API makes host-side by construction: the dicts / lists of per-object
tensors it returns and the track-id bookkeeping (one small device->host read of labels and matches per frame).
`forward` takes the reference's batch-1 call unchanged; `forward_batch` runs B independent sequences per call (one batched
backbone pass, per-sequence association).  SURVEY.md section 8(f) rows 1-3.
"""
import numpy as np
import torch
import torch.nn as nn

from . import _cabi, association
from .model_utils import Track4DBackbone


class Affinity(nn.Module):
    """MLP emb -> 4 emb -> 2 emb -> emb/2 -> emb/4 -> 1 + sigmoid on the DIFFERENCE of two object embeddings.  reference :226-245"""

    def __init__(self, emb_dims=137):
        super().__init__()
        self.emb_dims = emb_dims
        self.affinity = nn.Sequential(nn.Linear(emb_dims, emb_dims * 4), nn.ReLU(),
                                      nn.Linear(emb_dims * 4, emb_dims * 2), nn.ReLU(),
                                      nn.Linear(emb_dims * 2, emb_dims // 2), nn.ReLU(),
                                      nn.Linear(emb_dims // 2, emb_dims // 4), nn.ReLU(),
                                      nn.Linear(emb_dims // 4, 1), nn.Sigmoid())

    def forward(self, *input):
        return self.affinity((input[0] - input[1])[0])


def object_embeddings(points, seg, nseg):
    """points (139, P) feature columns of all objects side by side, seg (P,) object index of every column ->
    (nseg, 141) embeddings [centre(3), position variance(3), max feature(128), mean flow(3), mean rrv(2), rrv variance(2)]
    exactly as the reference assembles them per object (track4d.py:202-216; variances are population variances,
    computed two-pass like torch.var)."""
    if points.is_cuda and not (torch.is_grad_enabled() and points.requires_grad):
        return object_embeddings_packed(points, torch.bincount(seg, minlength=nseg).tolist())
    cnt = torch.zeros(nseg, device=points.device).index_add_(0, seg, torch.ones_like(seg, dtype=torch.float32))

    def seg_mean(x):                       # (C,P) -> (nseg,C)
        return torch.zeros(nseg, x.shape[0], device=x.device).index_add_(0, seg, x.t()) / cnt[:, None]

    def seg_var(x, mean):
        d = x.t() - mean[seg]
        return torch.zeros(nseg, x.shape[0], device=x.device).index_add_(0, seg, d * d) / cnt[:, None]

    pos, flow, rrv, feat = points[3:6], points[6:9], points[9:11], points[11:139]
    pos_m, rrv_m = seg_mean(pos), seg_mean(rrv)
    fmax = torch.full((nseg, feat.shape[0]), float("-inf"), device=points.device)
    fmax = fmax.scatter_reduce(0, seg[:, None].expand(-1, feat.shape[0]), feat.t(), reduce="amax", include_self=True)
    return torch.cat([pos_m, seg_var(pos, pos_m), fmax, seg_mean(flow), rrv_m, seg_var(rrv, rrv_m)], dim=1)


def object_embeddings_packed(points, counts):
    """The same embeddings for objects whose columns sit side by side in `points` (139, P) -- object i owns the next counts[i]
    columns -- in ONE kernel launch (rt_object_embeddings); inference only (no autograd)."""
    nseg = len(counts)
    off = torch.tensor([0] + list(np.cumsum(counts)), dtype=torch.int32).to(points.device, non_blocking=True)
    pts = points.detach().contiguous()
    out = torch.empty(nseg, 141, dtype=torch.float32, device=points.device)
    with torch.cuda.device_of(pts):
        _cabi.call("rt_object_embeddings", nseg, pts.shape[1], pts.data_ptr(), off.data_ptr(), out.data_ptr(),
                   torch.cuda.current_stream(pts.device).cuda_stream)
    return out


class Track4D(Track4DBackbone):
    """reference: src/models/track4d.py:13-223.  `args` needs npoints (FPS samples) and min_obj_points (DBSCAN min_samples)."""

    def __init__(self, args):
        super().__init__(args)
        self.min_samples = int(getattr(args, "min_obj_points", 2))
        self.eps = 1.5                                     # reference :36
        self.affinity = Affinity(141)
        self.register_parameter("bin_score", nn.Parameter(torch.tensor(1.0)))
        self.max_id = 0

    # -- reference-shaped methods ----------------------------------------------------------------------------------
    def forward(self, pc1, pc2, feature1, feature2, h, objects_prev, npts1=None, npts2=None):
        """Batch 1 with `objects_prev` a dict: the reference's call (track4d.py:49-65), same 10-tuple.
        Batch B with `objects_prev` a list of B dicts (B independent sequences at the same time step): see forward_batch.
        npts1 / npts2 (batch 1, beyond the reference's signature): point counts of a zero-padded pair whose two clouds differ
        in size (main_utils.pair_tensors); the padded columns are dropped again before clustering, so the 10-tuple is that of
        the unpadded pair."""
        if isinstance(objects_prev, (list, tuple)):
            if npts1 is not None:
                raise ValueError("forward_batch takes clouds of one size; variable-size pairs go one pair per call")
            return self.forward_batch(pc1, pc2, feature1, feature2, h, objects_prev)
        if npts1 is None:
            out = self.backbone(pc1, pc2, feature1, feature2, h)
            return self.track(pc1, feature1, out, objects_prev)
        if pc1.shape[0] != 1:
            raise ValueError("npts1 / npts2 are for a single padded pair (batch 1)")
        out = self.backbone(pc1, pc2, feature1, feature2, h, npts1=npts1, npts2=npts2)
        n1 = int(npts1[0])
        out = tuple(o if (o is None or i == 1) else o[..., :(int(npts2[0]) if i == 5 else n1)] for i, o in enumerate(out))
        return self.track(pc1[..., :n1], feature1[..., :n1], out, objects_prev)

    def forward_batch(self, pc1, pc2, feature1, feature2, h, objects_prev, max_ids=None):
        """B independent sequences per call: ONE batched backbone pass (h is (5,B,128), one recurrent state per sequence),
        then clustering / affinity / Sinkhorn / id bookkeeping per sequence -- the reference hard-codes batch 1 here
        (`squeeze(0)`, track4d.py:56).  `max_ids` = one id counter per sequence (default: the module-wide counter, ids then
        unique across sequences).  -> (h (5,B,128), pc1_warp (B,3,N), cls (B,N), frames, max_ids) with frames[b] = the
        per-sequence tail of the reference's tuple: (aff_list, aff_mat, indices1, confs, objects, timeout_obj_curr, objects_curr)."""
        B = pc1.shape[0]
        if len(objects_prev) != B:
            raise ValueError(f"forward_batch: {len(objects_prev)} object dicts for a batch of {B}")
        out = self.backbone(pc1, pc2, feature1, feature2, h)
        frames, new_ids = [], []
        shared = self.max_id
        for b in range(B):
            ob = tuple(None if o is None else (o[:, b:b + 1] if i == 1 else o[b:b + 1]) for i, o in enumerate(out))
            if max_ids is not None:
                self.max_id = int(max_ids[b])
            r = self.track(pc1[b:b + 1], feature1[b:b + 1], ob, objects_prev[b])
            frames.append(r[3:])
            if max_ids is not None:
                new_ids.append(self.max_id)
                self.max_id = shared
        return out[1], pc1 + out[0], out[2], frames, (new_ids if max_ids is not None else None)

    def track(self, pc1, feature1, backbone_out, objects_prev):
        """Everything of `forward` after the backbone (reference :52-65), given the backbone's 7-tuple."""
        output, h, cls, _, _, _, prop_features = backbone_out
        pc1_warp = pc1 + output
        pc1_features_warp = torch.cat((pc1_warp, pc1, output, feature1, prop_features), dim=1)
        mov_mask = (cls > 0.5).squeeze(0)
        objects_curr = self.clustering(pc1_features_warp[:, :, mov_mask])
        objects = dict()
        aff_list, aff_mat, indices1, confs = self.association_module(objects, objects_curr, objects_prev)
        return h, pc1_warp, cls, aff_list, aff_mat, indices1, confs, objects, dict(), objects_curr

    def clustering(self, pc1_features):
        """(1,139,Nm) features of the moving points -> list of (1,139,n_k) tensors, one per DBSCAN cluster, in the order
        the clusters first appear along the point index (the reference's defaultdict insertion order, :120-126)."""
        if pc1_features.shape[2] == 0:
            return []
        f = torch.cat((pc1_features[0, 3:9, :], pc1_features[0, 10:12, :]), dim=0).t().contiguous()
        labels = association.dbscan_labels(f.detach(), self.eps, self.min_samples).cpu().numpy()      # the frame's first host read
        # clusters in the order they first appear along the point index, points inside a cluster in index order: one stable
        # sort on the host, ONE gather on the device, the per-cluster tensors are views of its result
        keep = np.nonzero(labels != -1)[0]
        if keep.size == 0:
            return []
        labs = labels[keep]
        uniq, first = np.unique(labs, return_index=True)
        rank = np.empty(int(uniq.max()) + 1, dtype=np.int64)
        rank[uniq[np.argsort(first, kind="stable")]] = np.arange(uniq.size)
        order = np.argsort(rank[labs], kind="stable")
        counts = np.bincount(rank[labs], minlength=uniq.size).tolist()
        perm = torch.from_numpy(keep[order]).to(pc1_features.device, non_blocking=True)
        packed = pc1_features.index_select(2, perm)
        objs = list(torch.split(packed, counts, dim=2))
        self._packed = (objs, packed, counts)        # affinity_module reuses the side-by-side layout (no cat, no segment ids)
        return objs

    def affinity_module(self, objects_curr, objects_prev):
        """-> (aff_list (m*n,), aff_mat (1,m,n), m, n); empty inputs give ([], (1,0)) as the reference does (:218-222)."""
        m, n = len(objects_prev), len(objects_curr)
        dev = next(self.parameters()).device
        if m == 0 or n == 0:
            return [], torch.zeros(1, 0, device=dev), m, n

        def embed(objs):
            packed = getattr(self, "_packed", None)
            fast = objs[0].is_cuda and not (torch.is_grad_enabled() and any(o.requires_grad for o in objs))
            if fast and packed is not None and packed[0] is objs:                   # this frame's clusters: already side by side
                return object_embeddings_packed(packed[1][0], packed[2])
            pts = torch.cat([o[0] for o in objs], dim=1)
            if fast:
                return object_embeddings_packed(pts, [o.shape[2] for o in objs])
            seg = torch.cat([torch.full((o.shape[2],), i, dtype=torch.long, device=pts.device) for i, o in enumerate(objs)])
            return object_embeddings(pts, seg, len(objs))

        e_curr, e_prev = embed(objects_curr), embed(list(objects_prev.values()))
        diff = (e_curr[None, :, :] - e_prev[:, None, :]).reshape(m * n, -1)      # row-major (prev, curr), as the reference loops
        aff_list = self.affinity.affinity(diff).reshape(m * n)
        return aff_list, aff_list.reshape(m, n).unsqueeze(0), m, n

    def sinkhorn_module(self, aff_mat_tensor, indices1=None):
        return association.sinkhorn_module(aff_mat_tensor, indices1)

    def association_module(self, objects, objects_curr, objects_prev):
        """reference :135-164: matched objects keep the previous id, the others get fresh ids from self.max_id."""
        aff_list, aff_mat, m, n = self.affinity_module(objects_curr, objects_prev)
        indices1 = None
        confs = []
        if m > 0 and n > 0:
            try:
                indices1 = self.sinkhorn_module(aff_mat, indices1)
            except _cabi.RatrackError:
                # the reference wraps the association in try/except and hands out fresh ids on any failure (:143-160)
                for obj in objects_curr:
                    objects[self.max_id] = obj
                    self.max_id += 1
                    confs.append(0)
                return aff_list, aff_mat, None, confs
            idx = indices1[0]
            conf_all = aff_mat[0, idx.clamp(min=0), torch.arange(n, device=aff_mat.device)]
            both = torch.stack((idx.to(torch.float64), conf_all.detach().to(torch.float64))).cpu()   # the frame's second (last) host read
            idx_h, conf_h = both[0].long().tolist(), both[1].tolist()
            # (an unmatched object has index -1: the reference then reads aff_mat[0, -1, i]; only the sign of the test matters)
            prev_keys = list(objects_prev.keys())
            for i in range(n):
                if idx_h[i] == -1 or idx_h[i] >= m or conf_h[i] < 0.01:
                    objects[self.max_id] = objects_curr[i]
                    self.max_id += 1
                    confs.append(0)
                else:
                    objects[prev_keys[idx_h[i]]] = objects_curr[i]
                    confs.append(conf_all[i])
        else:
            for obj in objects_curr:
                objects[self.max_id] = obj
                self.max_id += 1
                confs.append(0)
        return aff_list, aff_mat, indices1, confs
