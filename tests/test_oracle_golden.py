"""CPU: the oracle restatement against the committed golden vectors.

tests/golden/backbone_*.npz are outputs of the UNMODIFIED reference Python (Track4D.backbone,
models/track4d.py:67-86) produced in the dev container by oracle/gen_golden.py;
tests/golden/pointnet2_vod_frames.npz are C-oracle outputs on the reference's real radar frames.
"""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from conftest import GOLDEN, ROOT
from oracle import backbone_oracle, pointnet2_oracle as P
from ratrack_b200 import synthetic


def _oracle_backbone(batch, n):
    from ratrack_b200.model_utils import Track4DBackbone

    class A:
        npoints = 512

    net = Track4DBackbone(A())
    sd = synthetic.make_state_dict(net, seed=1234)
    d = synthetic.make_batch(batch, n, seed=1234)
    t = {k: torch.from_numpy(v) for k, v in d.items()}
    return backbone_oracle.backbone(sd, t["pc1"], t["pc2"], t["ft1"], t["ft2"], torch.zeros(5, batch, 128))


@pytest.mark.parametrize("name,batch,n", [("backbone_n256_b1.npz", 1, 256), ("backbone_n1024_b2.npz", 2, 1024), ("backbone_n3000_b1.npz", 1, 3000)])
def test_backbone_oracle_matches_reference_golden(name, batch, n):
    g = np.load(os.path.join(GOLDEN, name))
    out, h, cls, cor, f1, f2, prop = _oracle_backbone(batch, n)
    # Tolerance (DESIGN.md "Parity"): 1e-5 of the tensor's scale, i.e. |d| <= 1e-5 * max(1, max|ref|).
    # north_star asks 1e-5 abs on the fp32 scene flow; flows here reach |7.5| and even this restatement,
    # built from the same torch-CPU ops as the reference modules, differs from them by 1.4e-5 abs at
    # N=1024 (fp32 re-association in the GRU / predictor), so a pure abs 1e-5 is below fp32 noise.
    assert np.abs(out.numpy() - g["flow"]).max() <= 1e-5 * max(1.0, np.abs(g["flow"]).max())
    assert np.abs(cls.numpy() - g["cls"]).max() <= 1e-5
    assert np.abs(h.numpy() - g["h"]).max() <= 1e-5
    assert abs(float(cor.double().sum()) - float(g["sum_cor"])) <= 1e-6 * float(g["abs_cor"])
    assert abs(float(prop.double().sum()) - float(g["sum_prop"])) <= 1e-6 * float(g["abs_prop"])
    if "cor" in g.files:
        assert np.abs(cor.numpy() - g["cor"]).max() <= 1e-5
        assert np.abs(f1.numpy() - g["f1"]).max() <= 1e-5
        assert np.abs(prop.numpy() - g["prop"]).max() <= 1e-5


def test_c_oracle_on_real_vod_frames():
    g = np.load(os.path.join(GOLDEN, "pointnet2_vod_frames.npz"))
    tags = sorted(k[4:] for k in g.files if k.startswith("xyz_"))
    assert len(tags) == 3
    for tag in tags:
        xyz = g[f"xyz_{tag}"][None]
        fps = P.furthest_point_sample(xyz, 512)
        assert np.array_equal(fps[0], g[f"fps_{tag}"])
        new_xyz = np.ascontiguousarray(xyz[0][fps[0]][None])
        for r, ns in ((2.0, 4), (4.0, 8), (16.0, 32)):
            assert np.array_equal(P.ball_query(r, ns, xyz, new_xyz)[0], g[f"bq_{tag}_{int(r)}_{ns}"])
        d2, i3 = P.three_nn_raw(xyz, new_xyz)
        assert np.array_equal(i3[0], g[f"nn_idx_{tag}"]) and np.array_equal(d2[0], g[f"nn_d2_{tag}"])


def test_c_oracle_matches_reference_kernels_golden():
    """tests/golden/ref_gpu_ops.npz = outputs of the reference's OWN CUDA kernels (unmodified
    /root/reference/src/lib/src, compiled for sm_100 into oracle/_ref) run on a B200 by
    oracle/gen_golden_ref_gpu.py.  The C restatement must reproduce them bit for bit."""
    g = np.load(os.path.join(GOLDEN, "ref_gpu_ops.npz"))
    cases = sorted({k.split("/")[0] for k in g.files})
    assert len(cases) == 4
    for case in cases:
        B, N, S = (int(x[1:]) for x in case.split("_"))
        d = synthetic.make_batch(B, N, seed=1234)
        xyz = np.ascontiguousarray(d["pc1"].transpose(0, 2, 1))
        fps = P.furthest_point_sample(xyz, S)
        assert np.array_equal(fps, g[f"{case}/fps"]), case
        new_xyz = np.ascontiguousarray(np.take_along_axis(xyz, fps[..., None].astype(np.int64), 1))
        for r, ns in ((2.0, 4), (4.0, 8), (8.0, 16), (16.0, 32)):
            assert np.array_equal(P.ball_query(r, ns, xyz, new_xyz), g[f"{case}/bq_{int(r)}_{ns}"]), (case, r)
        d2, i3 = P.three_nn_raw(xyz, new_xyz)
        assert np.array_equal(i3, g[f"{case}/nn_idx"]) and np.array_equal(d2, g[f"{case}/nn_d2"])
        kd, ki = P.knn_raw(16, new_xyz, xyz)
        assert np.array_equal(ki, g[f"{case}/knn_idx"]) and np.array_equal(kd, g[f"{case}/knn_d2"])
        feats = torch.randn((B, 8, S), generator=torch.Generator().manual_seed(1234)).numpy()
        assert np.array_equal(P.three_interpolate(feats, i3, g[f"{case}/interp_w"]), g[f"{case}/interp"])


def _bitrev(v, bits):
    r = 0
    for i in range(bits):
        r |= ((v >> i) & 1) << (bits - 1 - i)
    return r


@pytest.mark.parametrize("n", [5, 64, 100, 256, 300, 1024, 1500])
def test_fps_tie_rule(n):
    """The literal tree simulation in the C oracle equals the closed-form rule the CUDA kernel uses:
    among maxima, minimise (bitrev(k mod bs), k div bs)  (sampling_gpu.cu:86-91,143-203)."""
    rng = np.random.default_rng(n)
    # few distinct locations => massive ties
    xyz = rng.integers(0, 3, size=(2, n, 3)).astype(np.float32)
    m = min(n + 3, 40)
    got = P.furthest_point_sample(xyz, m)
    bs = P.opt_n_threads(n)
    logbs = bs.bit_length() - 1
    for b in range(2):
        temp = np.full(n, 1e10, np.float32)
        old = 0
        for j in range(1, m):
            d = ((xyz[b] - xyz[b, old]) ** 2).astype(np.float32)
            d2 = (d[:, 1] + d[:, 0] + d[:, 2]).astype(np.float32)  # small ints: exact in any order
            temp = np.minimum(temp, d2)
            mx = temp.max()
            cands = np.nonzero(temp == mx)[0]
            old = int(min(cands, key=lambda k: (_bitrev(int(k) % bs, logbs), int(k) // bs)))
            assert got[b, j] == old, (b, j)


def test_ball_query_semantics():
    xyz = np.array([[[0, 0, 0], [1, 0, 0], [0.5, 0, 0], [5, 5, 5], [0, 0.1, 0]]], np.float32)
    q = np.array([[[0, 0, 0], [100, 100, 100]]], np.float32)
    idx = P.ball_query(1.0, 4, xyz, q)
    assert idx[0, 0].tolist() == [0, 2, 4, 0]      # d2 < r2 strict: point 1 (d2 == 1) excluded; pad with first hit
    assert idx[0, 1].tolist() == [0, 0, 0, 0]      # no hit: caller's zero fill


def test_three_nn_fewer_than_three_known():
    u = np.zeros((1, 2, 3), np.float32)
    k = np.ones((1, 2, 3), np.float32)
    d2, idx = P.three_nn_raw(u, k)
    assert np.isinf(d2[0, 0, 2]) and idx[0, 0].tolist() == [0, 1, 0]


def test_knn_sorted_stable_and_capped():
    rng = np.random.default_rng(0)
    u = rng.normal(size=(1, 7, 3)).astype(np.float32)
    k = np.repeat(rng.normal(size=(1, 10, 3)).astype(np.float32), 2, axis=1)   # every point twice
    d2, idx = P.knn_raw(6, u, k)
    assert (np.diff(d2, axis=-1) >= 0).all()
    assert (idx[0, :, 0::2] + 1 == idx[0, :, 1::2]).all()   # duplicates: lower index first
    with pytest.raises(ValueError):
        P.knn_raw(201, u, k)


@pytest.mark.skipif(not os.path.isdir("/root/reference/src"), reason="reference tree only exists in the dev container")
def test_oracle_vs_live_reference_python():
    """Run the unmodified reference PNHead + FeatureCorrelator live (subprocess: the harness monkey-patches
    torch.Tensor.cuda) on a small random case and compare with the restatement."""
    code = r'''
import sys, numpy as np, torch
sys.path.insert(0, %r)
from oracle import ref_harness, backbone_oracle
from ratrack_b200 import synthetic
net = ref_harness.make_track4d(npoints=64)
net.load_state_dict(synthetic.make_state_dict(net, seed=1234)); net.eval()
d = synthetic.make_batch(2, 200, seed=99)
t = {k: torch.from_numpy(v) for k, v in d.items()}
with torch.no_grad():
    ref = net.backbone(t["pc1"], t["pc2"], t["ft1"], t["ft2"], torch.zeros(5, 2, 128))
got = backbone_oracle.backbone(net.state_dict(), t["pc1"], t["pc2"], t["ft1"], t["ft2"], torch.zeros(5, 2, 128), npoint=64)
for a, b in zip(ref, got):
    scale = max(1.0, float(a.abs().max()))
    assert float((a - b).abs().max()) <= 1e-5 * scale, float((a - b).abs().max())
print("OK")
''' % ROOT
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "OK" in r.stdout, r.stdout + r.stderr


@pytest.mark.parametrize("name,batch,n", [("train_step_n256_b1.npz", 1, 256), ("train_step_n512_b2.npz", 2, 512)])
def test_backbone_oracle_train_mode_matches_reference_train_forward(name, batch, n):
    """training=True (BatchNorm on batch statistics) against the forward outputs of the reference's own training step
    (tests/golden/train_step_*.npz, oracle/gen_golden_train.py)."""
    from ratrack_b200.model_utils import Track4DBackbone

    class A:
        npoints = 512

    g = np.load(os.path.join(GOLDEN, name))
    sd = synthetic.make_state_dict(Track4DBackbone(A()), seed=1234)
    d = synthetic.make_batch(batch, n, seed=1234)
    t = {k: torch.from_numpy(v) for k, v in d.items()}
    out = backbone_oracle.backbone(sd, t["pc1"], t["pc2"], t["ft1"], t["ft2"], torch.zeros(5, batch, 128), training=True,
                                   knn_override=(torch.from_numpy(g["knn12"]).long(), torch.from_numpy(g["knn11"]).long()))
    assert np.abs(out[0].numpy() - g["flow"]).max() <= 2e-5 * max(1.0, np.abs(g["flow"]).max())
    assert np.abs(out[2].numpy() - g["cls"]).max() <= 2e-5
