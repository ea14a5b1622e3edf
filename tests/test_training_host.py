"""Host-side logic of the round-2 training path that needs no GPU: dispatch of the dense layers, the switches that select the
reference's op-by-op dataflow, the cluster ordering of Track4D.clustering and the C-ABI argument checks of the new entries."""
import numpy as np
import torch

from ratrack_b200 import _cabi
from ratrack_b200.lib import dense_tc
from ratrack_b200.lib import pointnet2_utils as U
from ratrack_b200.lib.pytorch_utils import PointwiseConv2d
from ratrack_b200.model_utils import FeatureCorrelator, reference_dataflow


def test_dense_dispatch_keeps_cpu_and_narrow_layers_on_torch():
    x = torch.randn(5000, 64)
    w = torch.randn(32, 64)
    assert not dense_tc.covered(x, w)                                   # CPU tensor: torch's library GEMM
    y = dense_tc.linear(x, w, None)
    assert torch.allclose(y, x @ w.t(), atol=1e-5)
    conv = PointwiseConv2d(64, 32, 1)
    xi = torch.randn(2, 64, 50, 4)
    ref = torch.nn.functional.conv2d(xi, conv.weight, conv.bias)
    assert torch.allclose(conv(xi), ref, atol=1e-5)                     # same function as nn.Conv2d, channels-innermost result
    assert conv(xi).stride(1) == 1


def test_rows2d_views_and_copies():
    x = torch.randn(6, 10, 16)
    v, ld = dense_tc._rows2d(x)
    assert v.shape == (60, 16) and ld == 16 and v.data_ptr() == x.data_ptr()
    big = torch.randn(100, 48)
    v, ld = dense_tc._rows2d(big[:, 8:40])                              # strided rows stay a view with their leading dimension
    assert ld == 48 and v.shape == (100, 32) and v.data_ptr() == big[:, 8:40].data_ptr()
    t = torch.randn(16, 30).t()                                         # column stride != 1: one contiguous copy
    v, ld = dense_tc._rows2d(t)
    assert v.is_contiguous() and ld == 16


def test_reference_dataflow_switches_everything_off_and_back():
    before = (PointwiseConv2d.use_gemm, dense_tc.enabled, U.QueryAndGroup.rows_layout, FeatureCorrelator.fused_rows)
    assert all(before)
    with reference_dataflow():
        assert not any((PointwiseConv2d.use_gemm, dense_tc.enabled, U.QueryAndGroup.rows_layout, FeatureCorrelator.fused_rows))
        fc = FeatureCorrelator(16, in_channel=515, mlp=[256, 256, 256])
        assert not fc._rows_path_ok(torch.zeros(1, 3, 8), torch.zeros(1, 256, 8))
    assert (PointwiseConv2d.use_gemm, dense_tc.enabled, U.QueryAndGroup.rows_layout, FeatureCorrelator.fused_rows) == before
    try:
        with reference_dataflow():
            raise RuntimeError("boom")
    except RuntimeError:
        pass
    assert (PointwiseConv2d.use_gemm, dense_tc.enabled, U.QueryAndGroup.rows_layout, FeatureCorrelator.fused_rows) == before


def test_cluster_order_is_first_appearance_then_point_index():
    """The host-side part of Track4D.clustering: clusters in the order they first appear along the point index (the
    reference's defaultdict insertion order, src/models/track4d.py:120-126), points inside a cluster in index order."""
    labels = np.array([2, -1, 0, 2, 1, 0, -1, 1, 2, 0])
    keep = np.nonzero(labels != -1)[0]
    labs = labels[keep]
    uniq, first = np.unique(labs, return_index=True)
    rank = np.empty(int(uniq.max()) + 1, dtype=np.int64)
    rank[uniq[np.argsort(first, kind="stable")]] = np.arange(uniq.size)
    order = np.argsort(rank[labs], kind="stable")
    counts = np.bincount(rank[labs], minlength=uniq.size).tolist()
    perm = keep[order]
    # reference semantics, literally
    from collections import defaultdict
    ref = defaultdict(list)
    for i, lab in enumerate(labels.tolist()):
        if lab != -1:
            ref[lab].append(i)
    want = [v for v in ref.values()]
    got = np.split(perm, np.cumsum(counts)[:-1])
    assert [g.tolist() for g in got] == want and counts == [len(v) for v in want]


def test_new_entries_validate_their_arguments_without_a_gpu():
    L = _cabi.lib()
    assert L.rt_dense_tc_forward(10, 0, 4, 16, 4, 16, 4, 1, None, None, None, 0, 16, 4, None) < 0          # k = 0
    assert L.rt_dense_tc_forward(10, 8, 4, 16, 4, 16, 8, 1, None, None, None, 0, 16, 4, None) < 0          # ldx < k
    assert b"leading" in L.rt_last_error()
    assert L.rt_dense_tc_forward(10, 4, 4, 16, 4, 16, 4, 1, None, None, None, 7, 16, 4, None) < 0          # unknown activation
    assert L.rt_dense_tc_wgrad(10, 4, 4, None, 4, 16, 4, None, None, 16, None) < 0                         # null dy
    assert L.rt_cv1_forward(1, 8, 8, 16, 48, 16, 16, 16, 16, 16, 16, None, 16, None, None) < 0             # channels not a multiple of 32
    assert L.rt_cv1_forward(1, 8, 8, 64, 64, 16, 16, 16, 16, 16, 16, None, 16, None, None) < 0             # more than 32 neighbours
    assert L.rt_wsum_forward(1, 8, 16, 64, None, None, 16, 16, 16, 16, None) < 0                           # null x
    assert L.rt_act_grad(10, 2000, 1, 16, 16, 16, 16, 16, None) < 0                                        # more than 1024 columns
    assert L.rt_group_rows(1, 0, 8, 4, 4, 16, 16, 16, 16, 16, None) < 0                                    # no channels
    assert L.rt_absmax(None, 4, 16, None) < 0
    assert L.rt_object_embeddings(2, 10, None, 16, 16, None) < 0
