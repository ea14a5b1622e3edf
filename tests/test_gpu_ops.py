"""GPU parity of the ten pointnet2 ops: sm_100a kernels (through the C ABI) vs the CPU oracle,
vs the committed golden vectors and -- when oracle/_ref is present -- vs the reference's own
kernels on the same device.  Indices and copies bit-exact; atomics-based grads to 1e-5 relative."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle import pointnet2_oracle as P
from ratrack_b200 import synthetic
from ratrack_b200.lib import pointnet2_utils as U

pytestmark = pytest.mark.gpu


def _cloud(B, N, seed=1234):
    d = synthetic.make_batch(B, N, seed=seed)
    return np.ascontiguousarray(d["pc1"].transpose(0, 2, 1))


def _cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.mark.parametrize("B,N,S", [(1, 256, 512), (3, 1024, 512), (2, 322, 512), (2, 3000, 512), (2, 5, 9),
                                   (1, 64, 64), (2, 100, 37), (1, 2049, 128), (1, 4097, 64), (150, 512, 512)])
def test_fps_bit_exact(B, N, S):
    xyz = _cloud(B, N, seed=N)
    got = U.furthest_point_sample(_cu(xyz), S).cpu().numpy()
    assert np.array_equal(got, P.furthest_point_sample(xyz, S))


def test_fps_massive_ties_and_temp():
    rng = np.random.default_rng(3)
    for N in (256, 300, 1024, 700):
        xyz = rng.integers(0, 3, size=(2, N, 3)).astype(np.float32)
        got = U.furthest_point_sample(_cu(xyz), 64).cpu().numpy()
        assert np.array_equal(got, P.furthest_point_sample(xyz, 64)), N


@pytest.mark.parametrize("N,S", [(1024, 512), (700, 512), (256, 512), (322, 512), (1024, 256), (2000, 1024), (512, 128)])
def test_fps_of_an_fps_ordered_cloud_identity_shortcut(N, S):
    """Levels 2 and 3 of PNHead sample S of the S points level 1 selected.  fps_identity_kernel answers that with the identity
    permutation where it can PROVE the serial sampler would (csrc/fps.cu); everything else falls through to the serial
    kernel.  Against the oracle: FPS-ordered clouds (shortcut taken), N < S (level 1 pads with repeats of one point: ties,
    shortcut refused), exact duplicates inside the cloud, a shuffled cloud (not in FPS order) and a mixed batch."""
    xyz = _cloud(4, N, seed=N + S)
    lvl1 = P.furthest_point_sample(xyz, S)
    q = np.ascontiguousarray(np.stack([xyz[b, lvl1[b]] for b in range(4)]))            # what level 2 sees
    q[1, S // 2] = q[1, 3]                                                              # a duplicate inside an ordered cloud
    q[2] = q[2, np.random.default_rng(1).permutation(S)]                                # not in FPS order at all
    want = P.furthest_point_sample(q, S)
    got = U.furthest_point_sample(_cu(q), S).cpu().numpy()
    assert np.array_equal(got, want)
    if N >= 2 * S:
        assert np.array_equal(want[0], np.arange(S)) and not np.array_equal(want[2], np.arange(S))
    # the running-minimum buffer is left as the serial kernel leaves it
    from ratrack_b200 import pointnet2_cuda as C
    temp = torch.full((4, S), 1e10, device="cuda")
    idx = torch.empty(4, S, dtype=torch.int32, device="cuda")
    C.furthest_point_sampling_wrapper(4, S, S, _cu(q), temp, idx)
    _, temp2 = P.furthest_point_sample(q, S, return_temp=True)
    assert np.array_equal(idx.cpu().numpy(), want)
    assert np.array_equal(temp.cpu().numpy(), temp2)


def test_vod_frames_golden():
    g = np.load(os.path.join(GOLDEN, "pointnet2_vod_frames.npz"))
    for tag in sorted(k[4:] for k in g.files if k.startswith("xyz_")):
        xyz = _cu(g[f"xyz_{tag}"][None])
        fps = U.furthest_point_sample(xyz, 512)
        assert np.array_equal(fps.cpu().numpy()[0], g[f"fps_{tag}"])
        new_xyz = U.gather_operation(xyz.transpose(1, 2).contiguous(), fps).transpose(1, 2).contiguous()
        for r, ns in ((2.0, 4), (4.0, 8), (16.0, 32)):
            assert np.array_equal(U.ball_query(r, ns, xyz, new_xyz).cpu().numpy()[0], g[f"bq_{tag}_{int(r)}_{ns}"])
        d, i3 = U.three_nn(xyz, new_xyz)
        assert np.array_equal(i3.cpu().numpy()[0], g[f"nn_idx_{tag}"])
        assert np.array_equal(d.cpu().numpy()[0], np.sqrt(g[f"nn_d2_{tag}"]))


@pytest.mark.parametrize("N,S", [(1024, 512), (512, 512), (333, 77), (3000, 512), (2050, 100)])
@pytest.mark.parametrize("r,ns", [(2.0, 4), (4.0, 8), (8.0, 16), (16.0, 32), (0.01, 8), (1000.0, 64)])
def test_ball_query_bit_exact(N, S, r, ns):
    xyz = _cloud(2, N, seed=N + 1)
    new_xyz = np.ascontiguousarray(xyz[:, np.random.default_rng(0).permutation(N)[:S] if S <= N else np.arange(S) % N])
    got = U.ball_query(r, ns, _cu(xyz), _cu(new_xyz)).cpu().numpy()
    assert np.array_equal(got, P.ball_query(r, ns, xyz, new_xyz))


@pytest.mark.parametrize("n,m", [(512, 512), (1024, 512), (300, 2), (77, 1), (3000, 2500)])
def test_three_nn_bit_exact(n, m):
    u, k = _cloud(2, n, seed=5), _cloud(2, m, seed=6)
    d, i = U.three_nn(_cu(u), _cu(k))
    od, oi = P.three_nn(u, k)
    assert np.array_equal(i.cpu().numpy(), oi)
    assert np.array_equal(d.cpu().numpy(), od)


@pytest.mark.parametrize("k", [1, 3, 16, 17, 32, 64, 200])
def test_knn_bit_exact(k):
    u, kn = _cloud(2, 300, seed=7), _cloud(2, 2100, seed=8)
    d, i = U.knn(k, _cu(u), _cu(kn))
    od, oi = P.knn(k, u, kn)
    assert np.array_equal(i.cpu().numpy(), oi) and np.array_equal(d.cpu().numpy(), od)


def test_knn_k_over_200_raises():
    u = _cu(_cloud(1, 32))
    with pytest.raises(RuntimeError):
        U.knn(201, u, u)


@pytest.mark.parametrize("C,N,S,ns", [(3, 1024, 512, 8), (514, 1024, 512, 8), (64, 512, 512, 32), (5, 100, 33, 3)])
def test_group_and_gather_bit_exact_and_grads(C, N, S, ns):
    rng = np.random.default_rng(C)
    feats = rng.normal(size=(2, C, N)).astype(np.float32)
    idx = rng.integers(0, N, size=(2, S, ns)).astype(np.int32)
    f = _cu(feats).requires_grad_(True)
    out = U.grouping_operation(f, _cu(idx))
    assert np.array_equal(out.detach().cpu().numpy(), P.grouping_operation(feats, idx))
    go = rng.normal(size=out.shape).astype(np.float32)
    out.backward(_cu(go))
    ref = P.grouping_operation_grad(go, idx, N)
    assert np.allclose(f.grad.cpu().numpy(), ref, rtol=1e-5, atol=1e-5)
    # gather = group with nsample 1
    f2 = _cu(feats).requires_grad_(True)
    o2 = U.gather_operation(f2, _cu(idx[:, :, 0].copy()))
    assert np.array_equal(o2.detach().cpu().numpy(), P.gather_operation(feats, idx[:, :, 0]))
    g2 = rng.normal(size=o2.shape).astype(np.float32)
    o2.backward(_cu(g2))
    assert np.allclose(f2.grad.cpu().numpy(), P.gather_operation_grad(g2, idx[:, :, 0], N), rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("C,m,n", [(64, 512, 512), (128, 512, 1024), (7, 33, 101)])
def test_three_interpolate_bit_exact_and_grad(C, m, n):
    rng = np.random.default_rng(n)
    feats = rng.normal(size=(2, C, m)).astype(np.float32)
    idx = rng.integers(0, m, size=(2, n, 3)).astype(np.int32)
    w = rng.uniform(size=(2, n, 3)).astype(np.float32)
    w /= w.sum(-1, keepdims=True)
    f = _cu(feats).requires_grad_(True)
    out = U.three_interpolate(f, _cu(idx), _cu(w))
    assert np.array_equal(out.detach().cpu().numpy(), P.three_interpolate(feats, idx, w))
    go = rng.normal(size=out.shape).astype(np.float32)
    out.backward(_cu(go))
    assert np.allclose(f.grad.cpu().numpy(), P.three_interpolate_grad(go, idx, w, m), rtol=1e-5, atol=1e-5)


def test_reference_kernels_agree_with_oracle_and_product():
    """The pin: the reference's own kernels (oracle/_ref, compiled from /root/reference/src/lib/src)
    on this GPU == C oracle == product kernels, bit for bit."""
    from oracle import gen_golden_ref_gpu, ref_gpu

    ref = ref_gpu.load()
    if ref is None:
        pytest.skip("oracle/_ref/pointnet2_cuda.so not built")
    for (B, N, S) in gen_golden_ref_gpu.CASES:
        r = gen_golden_ref_gpu.run_case(ref, B, N, S)
        xyz = _cloud(B, N)
        fps = P.furthest_point_sample(xyz, S)
        assert np.array_equal(r["fps"], fps), (B, N)
        assert np.array_equal(U.furthest_point_sample(_cu(xyz), S).cpu().numpy(), r["fps"])
        new_xyz = np.ascontiguousarray(np.take_along_axis(xyz, fps[..., None].astype(np.int64), 1))
        for rad, ns in ((2.0, 4), (4.0, 8), (8.0, 16), (16.0, 32)):
            assert np.array_equal(r[f"bq_{int(rad)}_{ns}"], P.ball_query(rad, ns, xyz, new_xyz)), (N, rad)
        d2, i3 = P.three_nn_raw(xyz, new_xyz)
        assert np.array_equal(r["nn_idx"], i3) and np.array_equal(r["nn_d2"], d2)
        kd, ki = P.knn_raw(16, new_xyz, xyz)
        assert np.array_equal(r["knn_idx"], ki) and np.array_equal(r["knn_d2"], kd)
        feats = torch.randn((B, 8, S), generator=torch.Generator().manual_seed(1234)).numpy()
        assert np.array_equal(r["interp"], P.three_interpolate(feats, i3, r["interp_w"]))


@pytest.mark.parametrize("n,m,k", [(1024, 1024, 16), (300, 1500, 16), (64, 33, 32), (10, 16, 16)])
def test_knn_expanded_matches_torch_cpu_topk(n, m, k):
    """Cost-volume kNN (model_utils.py:17-39, 85-99): same neighbour SETS as the reference's torch code on the CPU
    wherever the k-th and (k+1)-th expanded-form distances differ; ascending order; ties -> lower index."""
    from oracle import backbone_oracle
    from ratrack_b200.model_utils import knn_point

    q, s = _cloud(2, n, seed=11), _cloud(2, m, seed=12)
    got = knn_point(k, _cu(s), _cu(q)).cpu()
    dist = backbone_oracle.square_distance(torch.from_numpy(q), torch.from_numpy(s))
    want = torch.topk(dist, k, dim=-1, largest=False, sorted=True)
    dg = torch.gather(dist, 2, got)
    assert (dg[..., 1:] >= dg[..., :-1]).all()                 # ascending in the reference's own distance values
    assert torch.equal(dg, want[0])                            # same multiset of distances => same sets up to exact ties
    kth = want[0][..., -1:]
    strict = dist < kth                                        # everything strictly inside the k-th distance must be there
    member = torch.zeros_like(dist, dtype=torch.bool).scatter_(2, got, True)
    assert (member | ~strict).all()


def test_gather_type_gradients_are_bitwise_repeatable():
    """group_points / gather_points / three_interpolate gradients: segmented sums in a fixed order (csrc/segsum.cu) --
    repeated calls agree bit for bit, including ball-query style indices where one point is repeated many times."""
    from ratrack_b200 import pointnet2_cuda as C

    rng = np.random.default_rng(11)
    B, Cc, N, P, S = 3, 37, 700, 512, 32
    idx = rng.integers(0, N, (B, P, S)).astype(np.int32)
    idx[:, :, 8:] = idx[:, :, :1]                                  # padding repeats the first neighbour
    g = rng.normal(size=(B, Cc, P, S)).astype(np.float32)
    outs = []
    for _ in range(3):
        grad = torch.zeros(B, Cc, N, device="cuda")
        C.group_points_grad_wrapper(B, Cc, N, P, S, _cu(g), _cu(idx), grad)
        outs.append(grad)
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2])
    want = P_group_grad(g, idx, N)
    assert np.abs(outs[0].cpu().numpy() - want).max() <= 1e-5 * np.abs(want).max()
    n, m = 1000, 512
    i3 = rng.integers(0, m, (B, n, 3)).astype(np.int32)
    w3 = rng.random((B, n, 3)).astype(np.float32)
    g3 = rng.normal(size=(B, Cc, n)).astype(np.float32)
    outs = []
    for _ in range(3):
        grad = torch.zeros(B, Cc, m, device="cuda")
        C.three_interpolate_grad_wrapper(B, Cc, n, m, _cu(g3), _cu(i3), _cu(w3), grad)
        outs.append(grad)
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2])
    want = np.zeros((B, Cc, m), np.float64)
    for b in range(B):
        for q in range(3):
            np.add.at(want[b].T, i3[b, :, q], (g3[b] * w3[b, :, q]).T.astype(np.float64))
    assert np.abs(outs[0].cpu().numpy() - want).max() <= 1e-5 * np.abs(want).max()


def P_group_grad(g, idx, N):
    B, Cc = g.shape[:2]
    want = np.zeros((B, Cc, N), np.float64)
    for b in range(B):
        np.add.at(want[b].T, idx[b].reshape(-1), g[b].reshape(Cc, -1).T.astype(np.float64))
    return want


@pytest.mark.parametrize("C,N,S,ns,r", [(2, 1024, 512, 4, 2.0), (32, 512, 512, 16, 8.0), (64, 512, 512, 32, 16.0),
                                        (517, 256, 128, 8, 4.0), (5, 100, 33, 3, 3.0)])
def test_query_and_group_rows_layout_equals_the_op_chain(C, N, S, ns, r):
    """QueryAndGroup in the channels-innermost layout (rt_group_rows / rt_group_rows_grad) against the reference's op-by-op
    chain (group_points(xyz) - new_xyz ; group_points(features) ; cat -- reference src/lib/pointnet2_utils.py:269-292) on the
    same kernels' channel-major path: forward bit-identical, feature gradient to 1e-6 of its scale (different, but in both
    cases fixed, summation orders), and bit-repeatable."""
    B = 3
    rng = np.random.default_rng(C + N)
    xyz = _cu(_cloud(B, N, seed=N + C))
    new_xyz = xyz[:, :S].contiguous() if S <= N else torch.cat([xyz, xyz], 1)[:, :S].contiguous()
    feats = _cu(rng.normal(size=(B, C, N)).astype(np.float32))
    grad = _cu(rng.normal(size=(B, 3 + C, S, ns)).astype(np.float32))
    q = U.QueryAndGroup(r, ns)
    outs = []
    for rows_layout in (True, False, True):
        U.QueryAndGroup.rows_layout = rows_layout
        try:
            f = feats.clone().requires_grad_(True)
            y = q(xyz, new_xyz, f)
            y.backward(grad)
        finally:
            U.QueryAndGroup.rows_layout = True
        outs.append((y.detach(), f.grad.detach()))
    (y_rows, g_rows), (y_ref, g_ref), (y_again, g_again) = outs
    assert y_rows.shape == (B, 3 + C, S, ns) and y_rows.stride(1) == 1          # channels innermost
    assert torch.equal(y_rows, y_ref)
    scale = float(g_ref.abs().max())
    assert float((g_rows - g_ref).abs().max()) <= 1e-6 * scale + 1e-6
    assert torch.equal(y_rows, y_again) and torch.equal(g_rows, g_again)
