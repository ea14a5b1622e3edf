"""GPU parity of Track4D.backbone (PNHead x3 + cost volume + flow decoder) against the oracle and the
golden outputs of the unmodified reference Python (tests/golden/backbone_*.npz)."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle import backbone_oracle
from ratrack_b200 import synthetic
from ratrack_b200.model_utils import Track4DBackbone

pytestmark = pytest.mark.gpu

# fp32 everywhere: torch's default cudnn.allow_tf32=True silently runs the 1x1 convs of the modular path in
# TF32 (~1e-3 relative); the reference never sets these flags (SURVEY.md hard part 3), parity needs them off.
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False

# Tolerances, as a fraction of the tensor's scale: |d| <= TOL * max(1, max|ref|)   (DESIGN.md "Parity").
# north_star asks 1e-5 abs on the fp32 scene flow.  Measured on B200 (tests/test_gpu_parity_floor.py measures it on
# every run; one run is committed as profiles/r2_parity_floor.txt):
#   * the reference's OWN GPU arithmetic (its unmodified CUDA kernels + torch fp32 cuDNN/cuBLAS layers, TF32 off)
#     differs from the same layers on the CPU by 0.8-1.7e-4 abs on the flow (|flow| <= 11), 2-4e-5 on h and
#     1.1e-5 abs on the per-point features: the GRU / global-max-pool / BatchNorm chain amplifies 1e-6-level
#     differences, so 1e-5 abs on the flow is below the reference's own cross-device noise;
#   * every mode of the product -- modular, fp32 SIMT engine and the default tcgen05 engine (split-fp16 products with
#     the correction terms accumulated first, tools/tc_precision.cu) -- is held to ONE table, set from that floor.
TOL_FP32 = {"f1": 1e-5, "f2": 1e-5, "cor": 1e-5, "prop": 1.5e-5, "flow": 3e-5, "h": 1.5e-4, "cls": 3e-5}
TOL_TC = TOL_FP32      # no separate, looser table for the tensor-core engine
TOL = TOL_TC


class Args:
    npoints = 512


def _net(fused):
    net = Track4DBackbone(Args())
    sd = synthetic.make_state_dict(net, seed=1234)
    net.load_state_dict(sd, strict=False)
    net.use_fused = fused
    net.capture_knn = True
    return net.cuda().eval(), sd


def _run(net, batch, n, seed=1234):
    d = synthetic.make_batch(batch, n, seed=seed)
    t = {k: torch.from_numpy(v).cuda() for k, v in d.items()}
    with torch.no_grad():
        out = net.backbone(t["pc1"], t["pc2"], t["ft1"], t["ft2"], torch.zeros(5, batch, 128, device="cuda"))
        knn = net.cost_volume_neighbours(t["pc1"], t["pc2"])
    if fused_engine := getattr(net, "_engine", None):
        torch.cuda.synchronize()
        fused_engine.check_status()
    return d, [o.cpu() for o in out], [k.cpu().long() for k in knn]


def knn_sets_equivalent(pc_q, pc_s, got, want, slack=8.0):
    """Tie-aware kNN comparison.  Rows whose index sets differ are accepted only when every index in
    the symmetric difference sits within `slack` ulps (of the expanded-form terms) of the k-th distance.
    Returns (#rows that differ, #rows that differ without a tie excuse)."""
    q = pc_q.transpose(0, 2, 1).astype(np.float64)
    s = pc_s.transpose(0, 2, 1).astype(np.float64)
    differ = bad = 0
    for b in range(got.shape[0]):
        for i in range(got.shape[1]):
            a, w = set(got[b, i].tolist()), set(want[b, i].tolist())
            if a == w:
                continue
            differ += 1
            d2 = ((s[b] - q[b, i]) ** 2).sum(-1)
            kth = np.sort(d2)[got.shape[2] - 1]
            mag = (q[b, i] ** 2).sum() + (s[b] ** 2).sum(-1).max()
            eps = slack * np.finfo(np.float32).eps * mag
            if any(abs(d2[j] - kth) > eps for j in a ^ w):
                bad += 1
    return differ, bad


@pytest.mark.parametrize("fused", [False, True])
@pytest.mark.parametrize("name,batch,n", [("backbone_n256_b1.npz", 1, 256), ("backbone_n1024_b2.npz", 2, 1024), ("backbone_n3000_b1.npz", 1, 3000)])
def test_backbone_vs_reference_golden(fused, name, batch, n):
    net, sd = _net(fused)
    if fused and not net.fused_available():
        pytest.skip("fused engine not built")
    g = np.load(os.path.join(GOLDEN, name))
    d, out, knn = _run(net, batch, n)
    # 1. neighbour sets: equal to what the reference's torch.topk chose, up to genuine distance ties
    diff12, bad12 = knn_sets_equivalent(d["pc1"], d["pc2"], knn[0].numpy(), g["knn12"])
    diff11, bad11 = knn_sets_equivalent(d["pc1"], d["pc1"], knn[1].numpy(), g["knn11"])
    assert bad12 == 0 and bad11 == 0, (diff12, bad12, diff11, bad11)
    # 2. values: against the oracle replaying the product's neighbour choice
    c = {k: torch.from_numpy(v) for k, v in d.items()}
    ref = backbone_oracle.backbone(sd, c["pc1"], c["pc2"], c["ft1"], c["ft2"], torch.zeros(5, batch, 128),
                                   knn_override=tuple(knn))
    names = ["flow", "h", "cls", "cor", "f1", "f2", "prop"]
    for nm, a, r in zip(names, out, ref):
        scale = max(1.0, float(r.abs().max()))
        err = float((a - r).abs().max())
        assert err <= (TOL_TC if fused else TOL_FP32)[nm] * scale, (nm, err, scale)
    # 3. and directly against the reference-generated golden flow when no neighbour set differed
    if diff12 == 0 and diff11 == 0:
        err = np.abs(out[0].numpy() - g["flow"]).max()
        assert err <= (TOL_TC if fused else TOL_FP32)["flow"] * max(1.0, np.abs(g["flow"]).max()), err


def test_modules_keep_reference_state_dict_surface():
    net, sd = _net(False)
    keys = set(net.state_dict().keys())
    for k in ["pn_head.sa1.mlps.0.layer0.conv.weight", "pn_head.sa1.mlps.1.layer2.bn.bn.running_var",
              "pn_head.fp2.mlp.layer0.conv.weight", "pn_head.linear3.bias", "fc_layer.mlp_convs.0.weight",
              "fc_layer.weightnet2.mlp_convs.2.bias", "fd_layer.mse.sa1.mlps.0.layer0.conv.weight",
              "fd_layer.fp.sf_mlp.2.1.running_mean", "fd_layer.cp.linear.weight", "fd_layer.torchGRU.weight_hh_l4"]:
        assert k in keys, k
    assert net.state_dict()["fd_layer.mse.sa1.mlps.0.layer0.conv.weight"].shape == (16, 517, 1, 1)
    assert net.state_dict()["fc_layer.mlp_convs.0.weight"].shape == (256, 515, 1, 1)


def test_train_mode_backward_runs():
    """forward+backward through the modular path (autograd over the CUDA grad kernels)."""
    net, _ = _net(False)
    net.train()
    d = synthetic.make_batch(2, 256, seed=5)
    t = {k: torch.from_numpy(v).cuda() for k, v in d.items()}
    out = net.backbone(t["pc1"], t["pc2"], t["ft1"], t["ft2"], torch.zeros(5, 2, 128, device="cuda"))
    loss = out[0].square().mean() + out[2].mean()
    loss.backward()
    g = net.pn_head.sa1.mlps[0][0].conv.weight.grad
    assert g is not None and torch.isfinite(g).all() and float(g.abs().sum()) > 0


def test_tensor_core_kernels_match_simt_chain():
    """A/B inside the engine: tcgen05 cost volume (fp16 hi/lo split, 3 MMAs per product) vs the fp32 SIMT chain of
    the same dataflow -- isolates the split arithmetic from everything else."""
    from ratrack_b200.engine import FusedBackbone

    net, _ = _net(True)
    d = synthetic.make_batch(3, 1000, seed=77)    # 3000 points: not a multiple of the 8-point tile x grid
    t = {k: torch.from_numpy(v).cuda() for k, v in d.items()}
    eng = FusedBackbone(net)
    outs = {}
    for mode in (False, True):
        eng.set_flags(costvol_tc=mode, mlp_tc=mode)
        with torch.no_grad():
            outs[mode] = [o.clone() for o in eng(t["pc1"], t["pc2"], t["ft1"], t["ft2"], None)]
        torch.cuda.synchronize()
        eng.check_status()
    cor_s, cor_t = outs[False][3], outs[True][3]
    scale = float(cor_s.abs().max())
    assert float((cor_s - cor_t).abs().max()) <= 1e-5 * scale
    for i, nm in ((4, "f1"), (5, "f2"), (6, "prop")):
        assert float((outs[False][i] - outs[True][i]).abs().max()) <= TOL[nm] * float(outs[False][i].abs().max()), nm
    assert float((outs[False][0] - outs[True][0]).abs().max()) <= TOL["flow"] * float(outs[False][0].abs().max())


def test_fused_handles_odd_sizes():
    """N not a multiple of 4/8/16, N < npoints, batch 1 and 5: the engine vs the modular path on the same device."""
    for batch, n in ((1, 333), (5, 700), (2, 256), (2, 3000)):
        net, _ = _net(True)
        d = synthetic.make_batch(batch, n, seed=n)
        t = {k: torch.from_numpy(v).cuda() for k, v in d.items()}
        with torch.no_grad():
            net.use_fused = True
            a = net.backbone(t["pc1"], t["pc2"], t["ft1"], t["ft2"], None)
            net.use_fused = False
            b = net.backbone(t["pc1"], t["pc2"], t["ft1"], t["ft2"], None)
        for nm, x, y in zip(["flow", "h", "cls", "cor", "f1", "f2", "prop"], a, b):
            scale = max(1.0, float(y.abs().max()))
            assert float((x - y).abs().max()) <= 2 * TOL[nm] * scale, (batch, n, nm)


def test_two_lanes_equal_one_lane():
    """Splitting a batch over two concurrent lanes must not change a single bit of any output (every pair is
    independent; h is (5, B, 128) so the lanes write strided slices of it)."""
    from ratrack_b200.engine import FusedBackbone

    net, _ = _net(True)
    d = synthetic.make_batch(9, 512, seed=3)     # odd batch: lanes of 5 and 4 pairs
    t = {k: torch.from_numpy(v).cuda() for k, v in d.items()}
    h = torch.randn(5, 9, 128, device="cuda") * 0.1
    eng = FusedBackbone(net)
    outs = {}
    for lanes in (False, True):
        eng.set_flags(two_lanes=lanes)
        assert eng.num_lanes(9) == (2 if lanes else 1)
        with torch.no_grad():
            outs[lanes] = [o.clone() for o in eng(t["pc1"], t["pc2"], t["ft1"], t["ft2"], h, want_knn=True)]
            outs[lanes] += [k.clone() for k in eng.last_knn]
        torch.cuda.synchronize()
        eng.check_status()
    for a, b in zip(outs[False], outs[True]):
        assert torch.equal(a, b)


def test_fp16_range_guard_is_loud_and_falls_back():
    """an activation beyond the fp16 hi/lo range must never yield silently saturated outputs on the DEFAULT
    path.  The step's outputs become NaN on the device, the status arrives without a host sync, and the engine continues
    on its fp32 SIMT kernels; checked_forward=True re-runs the step at once."""
    import warnings

    d = synthetic.make_batch(2, 256, seed=3)
    t = {k: torch.from_numpy(v).cuda() for k, v in d.items()}
    for checked in (False, True):
        net, _ = _net(True)
        with torch.no_grad():
            net.fc_layer.mlp_convs[1].bias.add_(1.0e5)      # cost-volume layer-2 activations ~1e5 > 65504
        net.refresh_engine()
        net.checked_forward = checked
        with torch.no_grad(), warnings.catch_warnings(record=True) as w:
            warnings.simplefilter("always")
            first = net.backbone(t["pc1"], t["pc2"], t["ft1"], t["ft2"], None)
            torch.cuda.synchronize()
            second = net.backbone(t["pc1"], t["pc2"], t["ft1"], t["ft2"], None)
            torch.cuda.synchronize()
            net.use_fused = False
            want = net.backbone(t["pc1"], t["pc2"], t["ft1"], t["ft2"], None)
        if checked:
            assert torch.isfinite(first[0]).all() and torch.isfinite(first[2]).all()
        else:
            assert torch.isnan(first[0]).all() and torch.isnan(first[2]).all() and torch.isnan(first[1]).all()
        assert not net._engine.tensor_cores and any("fp16" in str(x.message) for x in w)
        assert torch.isfinite(second[0]).all()
        assert float((second[0] - want[0]).abs().max()) <= 1e-4 * max(1.0, float(want[0].abs().max()))


def test_out_of_range_weights_select_the_simt_kernels():
    import warnings

    net, _ = _net(True)
    with torch.no_grad():
        net.fd_layer.cp.sf_mlp[0][0].weight[0, 0] = 1.0e6      # folded: 2^10 * W far beyond 65504
    net.refresh_engine()
    d = synthetic.make_batch(1, 256, seed=4)
    t = {k: torch.from_numpy(v).cuda() for k, v in d.items()}
    with torch.no_grad(), warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        a = net.backbone(t["pc1"], t["pc2"], t["ft1"], t["ft2"], None)
        net.use_fused = False
        b = net.backbone(t["pc1"], t["pc2"], t["ft1"], t["ft2"], None)
    assert not net._engine.tensor_cores and any("fp16" in str(x.message) for x in w)
    assert torch.isfinite(a[2]).all() and float((a[2] - b[2]).abs().max()) <= 1e-4


def test_engine_snapshot_follows_the_parameters():
    """load_state_dict / .to() after an eval forward must not leave the folded-weight snapshot behind."""
    net, sd = _net(True)
    d = synthetic.make_batch(1, 256, seed=5)
    t = {k: torch.from_numpy(v).cuda() for k, v in d.items()}
    with torch.no_grad():
        a = net.backbone(t["pc1"], t["pc2"], t["ft1"], t["ft2"], None)[0].clone()
        sd2 = {k: (v * 1.25 if k.endswith("pn_head.linear3.weight") else v) for k, v in net.state_dict().items()}
        net.load_state_dict(sd2)
        b = net.backbone(t["pc1"], t["pc2"], t["ft1"], t["ft2"], None)[0].clone()
        net.use_fused = False
        want = net.backbone(t["pc1"], t["pc2"], t["ft1"], t["ft2"], None)[0]
    assert float((a - b).abs().max()) > 1e-3
    assert float((b - want).abs().max()) <= TOL["flow"] * max(1.0, float(want.abs().max()))


def _vod_like_frames():
    """The three real VoD radar frames the reference ships (xyz from the fixture, synthetic RCS / radial velocity)."""
    g = np.load(os.path.join(GOLDEN, "pointnet2_vod_frames.npz"))
    rng = np.random.default_rng(3)
    out = []
    for tag in sorted(k[4:] for k in g.files if k.startswith("xyz_")):
        xyz = g[f"xyz_{tag}"]
        rec = np.zeros((xyz.shape[0], 7), np.float32)
        rec[:, 0:3] = xyz
        rec[:, 3] = rng.normal(-13, 12, xyz.shape[0])
        rec[:, 4] = rng.normal(-2, 1.6, xyz.shape[0])
        out.append(rec)
    return out


def test_variable_size_batch_equals_every_pair_alone():
    """SURVEY 8f row 4 / VERDICT r1 'missing' 4: real frames have 242..352 points each.  One padded batch with per-cloud
    counts (rt_backbone_forward_varlen) must give, for every pair, bit for bit what that pair gives alone at its own size:
    FPS with the tie-break of the cloud's own size (two size classes here: 242 -> CTA size 128, 322 / 352 -> 256), ball query /
    kNN / global max-pool restricted to the cloud's own points; columns of padded points are zero."""
    from ratrack_b200 import data_io

    f = _vod_like_frames()
    syn = synthetic.make_batch(2, 300, seed=9)
    extra = np.concatenate([syn["pc1"][0].T, syn["ft1"][0].T, np.zeros((300, 2), np.float32)], axis=1)      # a 300-point frame
    pairs = [(f[0], f[1]), (f[1], f[2]), (f[2], f[0]), (f[2], f[2]), (extra, f[0])]
    net, _ = _net(True)
    batcher = data_io.PaddedBatcher(batch=len(pairs), pin=False)
    for i, (a, b) in enumerate(pairs):
        batcher.add(i, a, b)
    (keys, pc1, pc2, ft1, ft2, n1, n2), = list(batcher.flush())
    h = torch.randn(5, len(pairs), 128, device="cuda") * 0.1
    with torch.no_grad():
        out = [o.clone() for o in net.backbone(pc1.cuda(), pc2.cuda(), ft1.cuda(), ft2.cuda(), h, npts1=n1, npts2=n2)]
    torch.cuda.synchronize()
    net._engine.check_status()
    names = ["flow", "h", "cls", "cor", "f1", "f2", "prop"]
    for i, (a, b) in enumerate(pairs):
        na, nb = a.shape[0], b.shape[0]
        if na != nb:
            # alone, a pair must have one N for both clouds (the reference's loader resamples): check the pc1-side outputs of
            # unequal pairs through a second variable-size call of batch 1 instead
            with torch.no_grad():
                solo = net.backbone(pc1[i:i + 1].cuda(), pc2[i:i + 1].cuda(), ft1[i:i + 1].cuda(), ft2[i:i + 1].cuda(), h[:, i:i + 1].contiguous(),
                                    npts1=n1[i:i + 1], npts2=n2[i:i + 1])
            for nm, x, y in zip(names, out, solo):
                xs = x[:, i:i + 1] if nm == "h" else x[i:i + 1]
                assert torch.equal(xs, y), (i, nm)
            continue
        p1, p2, q1, q2 = data_io.frame_pair_inputs(a, b)
        with torch.no_grad():
            solo = net.backbone(torch.from_numpy(p1).cuda(), torch.from_numpy(p2).cuda(), torch.from_numpy(q1).cuda(),
                                torch.from_numpy(q2).cuda(), h[:, i:i + 1].contiguous())
        for nm, x, y in zip(names, out, solo):
            if nm == "h":
                assert torch.equal(x[:, i:i + 1], y), (i, nm)
            else:
                assert torch.equal(x[i:i + 1, ..., :na], y), (i, nm, float((x[i:i + 1, ..., :na] - y).abs().max()))
                assert float(x[i, ..., (nb if nm == "f2" else na):].abs().max() if x.shape[-1] > max(na, nb) else 0.0) == 0.0
    # equal-size pairs exist in the batch (pair 3: 242 / 242), unequal ones too
    assert any(a.shape[0] == b.shape[0] for a, b in pairs) and any(a.shape[0] != b.shape[0] for a, b in pairs)
