"""CPU: edge cases of the C oracle (oracle/pointnet2_oracle.c) stated by SURVEY.md Appendix A -- the contract the CUDA
kernels are held to in tests/test_gpu_ops.py.  Everything here is checked against an independent numpy statement of the same rule."""
import numpy as np
import pytest

from oracle import pointnet2_oracle as P


def test_fps_more_samples_than_distinct_points_pads_with_index_zero():
    """A.1: once every temp is 0 the strict '>' keeps besti = 0 -- configs[0] (N=256, npoint=512) relies on it
    (sampling_gpu.cu:129-140,143-203)."""
    rng = np.random.default_rng(3)
    distinct = rng.normal(size=(5, 3)).astype(np.float32)
    xyz = distinct[rng.integers(0, 5, size=64)][None]                    # 64 points at 5 locations
    xyz[0, 0] = distinct[0]
    idx, temp = P.furthest_point_sample(xyz, 16, return_temp=True)
    assert idx[0, 0] == 0
    picked = xyz[0, idx[0, :5]]
    assert len({tuple(p) for p in picked.tolist()}) == 5                 # the first five picks visit every location once
    assert (idx[0, 5:] == 0).all()                                       # nothing is further than 0 any more
    assert (temp == 0).all()


def test_fps_zero_samples_and_single_point():
    xyz = np.zeros((2, 7, 3), np.float32)
    assert P.furthest_point_sample(xyz, 0).shape == (2, 0)               # m <= 0: early return (sampling_gpu.cu:217)
    one = np.ones((1, 1, 3), np.float32)
    assert P.furthest_point_sample(one, 4).tolist() == [[0, 0, 0, 0]]


def _bitrev(v, bits):
    return int(format(v, f"0{bits}b")[::-1], 2) if bits else 0


def test_fps_nan_points_keep_their_initial_distance_and_win():
    """`min(d, temp[k])` is CUDA's fminf: a NaN distance leaves temp[k] untouched (sampling_gpu.cu:133-135), so a NaN point
    keeps the 1e10 it started with and is selected as soon as round 1 -- and once it is the reference point every distance is
    NaN, nothing changes any more, and the same index repeats.  (SURVEY.md A.1 says NaN points never win; the source says
    otherwise, and the oracle and the CUDA kernel follow the source.)"""
    rng = np.random.default_rng(5)
    n, m = 40, 20
    xyz = rng.normal(size=(1, n, 3)).astype(np.float32)
    bad = [7, 19, 33]
    xyz[0, bad] = np.nan
    idx, temp = P.furthest_point_sample(xyz, m, return_temp=True)
    bs = P.opt_n_threads(n)
    bits = bs.bit_length() - 1
    want = min(bad, key=lambda k: (_bitrev(k % bs, bits), k // bs))      # all three tie at 1e10: the tree's rule picks
    assert idx[0].tolist() == [0] + [want] * (m - 1)
    assert (temp[0, bad] == np.float32(1e10)).all()
    d0 = ((xyz[0] - xyz[0, 0]).astype(np.float64) ** 2)
    ok = [k for k in range(n) if k not in bad]
    assert np.allclose(temp[0, ok], d0[ok].sum(-1), rtol=1e-6)           # only round 1 (reference point 0) ever lowered them


def test_ball_query_rejects_nan_and_keeps_index_order():
    """A.2: first nsample hits in index order, strict d2 < r2 (fp32 r*r), NaN distances rejected."""
    rng = np.random.default_rng(11)
    xyz = rng.uniform(-2, 2, size=(3, 50, 3)).astype(np.float32)
    xyz[1, 4] = np.nan
    q = rng.uniform(-2, 2, size=(3, 9, 3)).astype(np.float32)
    r, ns = np.float32(1.3), 6
    got = P.ball_query(float(r), ns, xyz, q)
    r2 = np.float32(r * r)
    for b in range(3):
        for j in range(9):
            d = xyz[b] - q[b, j]
            dx, dy, dz = d[:, 0], d[:, 1], d[:, 2]
            # A.0 rounding order: t = dy*dy (rounded), t = fma(dx, dx, t), d2 = fma(dz, dz, t); in fp64 the products of
            # fp32 values are exact, so rounding once per step to fp32 reproduces each fma
            t = (dy.astype(np.float64) * dy).astype(np.float32)
            t = (dx.astype(np.float64) * dx + t).astype(np.float32)
            d2 = (dz.astype(np.float64) * dz + t).astype(np.float32)
            hits = [k for k in range(50) if d2[k] < r2]                  # NaN < r2 is False
            want = [0] * ns if not hits else (hits[:ns] + [hits[0]] * ns)[:ns]
            assert got[b, j].tolist() == want, (b, j)


def test_group_and_gather_are_plain_index_copies_and_their_grads_scatter_adds():
    """A.5 (group_points_gpu.cu:8-25,47-66; sampling_gpu.cu:8-24,46-63)."""
    rng = np.random.default_rng(2)
    B, C, N, S, ns = 2, 5, 17, 6, 4
    f = rng.normal(size=(B, C, N)).astype(np.float32)
    gi = rng.integers(0, N, size=(B, S, ns)).astype(np.int32)
    out = P.grouping_operation(f, gi)
    for b in range(B):
        assert np.array_equal(out[b], f[b][:, gi[b]])
    go = rng.integers(-3, 4, size=(B, C, S, ns)).astype(np.float32)     # small integers: the sums are exact in any order
    want = np.zeros((B, C, N), np.float32)
    for b in range(B):
        for c in range(C):
            np.add.at(want[b, c], gi[b].ravel(), go[b, c].ravel())
    assert np.array_equal(P.grouping_operation_grad(go, gi, N), want)
    fi = rng.integers(0, N, size=(B, S)).astype(np.int32)
    assert np.array_equal(P.gather_operation(f, fi), np.stack([f[b][:, fi[b]] for b in range(B)]))
    g1 = rng.integers(-3, 4, size=(B, C, S)).astype(np.float32)
    want = np.zeros((B, C, N), np.float32)
    for b in range(B):
        for c in range(C):
            np.add.at(want[b, c], fi[b], g1[b, c])
    assert np.array_equal(P.gather_operation_grad(g1, fi, N), want)


def test_three_nn_ties_keep_the_lower_index_in_the_better_slot():
    """A.3: strict '<' cascade over doubles initialised at 1e40 (interpolate_gpu.cu:81-124)."""
    known = np.array([[[1, 0, 0], [0, 1, 0], [0, 0, 1], [1, 0, 0], [2, 0, 0]]], np.float32)   # 0, 1, 2, 3 all at d2 = 1
    d2, idx = P.three_nn_raw(np.zeros((1, 1, 3), np.float32), known)
    assert idx[0, 0].tolist() == [0, 1, 2] and d2[0, 0].tolist() == [1.0, 1.0, 1.0]
    d2, idx = P.three_nn_raw(np.zeros((1, 1, 3), np.float32), known[:, :1])
    assert idx[0, 0].tolist() == [0, 0, 0] and d2[0, 0, 0] == 1.0 and np.isinf(d2[0, 0, 1:]).all()


def test_three_interpolate_rounding_order_and_grad():
    """A.0: t = w1*p1 (rounded); t = fma(w0, p0, t); out = fma(w2, p2, t)  (interpolate_gpu.cu:149-169); the gradient
    scatters grad * w into the three sources (:192-214)."""
    rng = np.random.default_rng(9)
    B, C, M, n = 2, 4, 11, 13
    f = rng.normal(size=(B, C, M)).astype(np.float32)
    idx = rng.integers(0, M, size=(B, n, 3)).astype(np.int32)
    w = rng.uniform(0.1, 1, size=(B, n, 3)).astype(np.float32)
    w /= w.sum(-1, keepdims=True)
    got = P.three_interpolate(f, idx, w)
    for b in range(B):
        p = f[b][:, idx[b]].astype(np.float64)                           # (C, n, 3)
        ww = w[b].astype(np.float64)
        t = (ww[:, 1] * p[:, :, 1]).astype(np.float32)
        t = (ww[:, 0] * p[:, :, 0] + t).astype(np.float32)
        want = (ww[:, 2] * p[:, :, 2] + t).astype(np.float32)
        assert np.array_equal(got[b], want)
    go = rng.integers(-2, 3, size=(B, C, n)).astype(np.float32)
    wq = (rng.integers(1, 4, size=(B, n, 3)) / 4).astype(np.float32)     # quarter weights: products and sums exact
    got = P.three_interpolate_grad(go, idx, wq, M)
    want = np.zeros((B, C, M), np.float32)
    for b in range(B):
        for c in range(C):
            for j in range(3):
                np.add.at(want[b, c], idx[b, :, j], go[b, c] * wq[b, :, j])
    assert np.array_equal(got, want)


@pytest.mark.parametrize("n,bs", [(1, 1), (2, 2), (255, 128), (256, 256), (322, 256), (512, 512), (1000, 512), (1024, 1024), (3000, 1024)])
def test_opt_n_threads(n, bs):
    """cuda_utils.h:10-14: the largest power of two <= n, clamped to [1, 1024] -- it fixes the FPS tie rule (A.1)."""
    assert P.opt_n_threads(n) == bs


def _d2(p, q):
    """A.0 on the host: t = dy*dy (rounded to fp32), t = fma(dx, dx, t), d2 = fma(dz, dz, t).  Products of fp32 values are exact
    in fp64, so rounding each step's fp64 sum to fp32 reproduces the fused operation (up to double rounding, which the fixed seeds
    below do not hit)."""
    d = (p - q).astype(np.float32)
    dx, dy, dz = (d[..., i].astype(np.float64) for i in range(3))
    t = (dy * dy).astype(np.float32)
    t = (dx * dx + t).astype(np.float32)
    return (dz * dz + t).astype(np.float32)


@pytest.mark.parametrize("n,m", [(7, 7), (100, 40), (322, 64), (1024, 48), (1500, 32)])
def test_fps_random_clouds_against_a_numpy_statement(n, m):
    """Random float clouds (no ties to speak of) with 1 % duplicated points (ties on purpose, as in the VoD frames): running
    minimum, arg-max with the tree's tie rule, A.0 distance arithmetic."""
    rng = np.random.default_rng(n)
    xyz = rng.normal(scale=10.0, size=(2, n, 3)).astype(np.float32)
    dup = rng.integers(0, n, size=(2, max(1, n // 100), 2))
    for b in range(2):
        xyz[b, dup[b, :, 0]] = xyz[b, dup[b, :, 1]]
    got, temp_got = P.furthest_point_sample(xyz, m, return_temp=True)
    bs = P.opt_n_threads(n)
    bits = bs.bit_length() - 1
    rank = np.array([(_bitrev(k % bs, bits), k // bs) for k in range(n)])
    order = np.lexsort((rank[:, 1], rank[:, 0]))                          # positions in the tree's priority order
    for b in range(2):
        temp = np.full(n, 1e10, np.float32)
        old = 0
        for j in range(1, m):
            temp = np.minimum(temp, _d2(xyz[b], xyz[b, old]))
            cands = temp[order]
            old = int(order[int(np.argmax(cands))])                      # first maximum in priority order
            assert got[b, j] == old, (b, j)
        assert np.array_equal(temp_got[b], temp)


def test_three_nn_and_knn_against_a_stable_sort():
    """A.3 / A.4: ascending distances, ties to the lower index -- a stable argsort of the A.0 distances."""
    rng = np.random.default_rng(21)
    known = rng.normal(scale=3.0, size=(2, 60, 3)).astype(np.float32)
    known[:, 30:40] = known[:, 10:20]                                    # duplicates: exact ties
    unknown = rng.normal(scale=3.0, size=(2, 25, 3)).astype(np.float32)
    d2_3, i3 = P.three_nn_raw(unknown, known)
    d2_k, ik = P.knn_raw(9, unknown, known)
    for b in range(2):
        for q in range(25):
            d = _d2(known[b], unknown[b, q])
            o = np.argsort(d, kind="stable")
            assert i3[b, q].tolist() == o[:3].tolist() and np.array_equal(d2_3[b, q], d[o[:3]])
            assert ik[b, q].tolist() == o[:9].tolist() and np.array_equal(d2_k[b, q], d[o[:9]])
