"""ratrack_b200.losses against the reference's own loss module: loss values stored in tests/golden/train_step_*.npz were
computed by the reference's src/losses/loss.py on the reference's train-mode outputs (oracle/gen_golden_train.py); here
the mirrors are evaluated on those stored outputs (CPU) and must reproduce them."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from ratrack_b200 import losses, synthetic


@pytest.mark.parametrize("name,batch,n", [("train_step_n256_b1.npz", 1, 256), ("train_step_n512_b2.npz", 2, 512)])
def test_loss_terms_match_reference_loss_module(name, batch, n):
    g = np.load(os.path.join(GOLDEN, name))
    d = synthetic.make_batch(batch, n, seed=1234)
    pc1_wrap = torch.from_numpy(d["pc1"]) + torch.from_numpy(g["flow"])
    sf = losses.flow_loss(pc1_wrap, torch.from_numpy(g["gt_flow"]), reduction="first")
    assert abs(float(sf) - float(g["sf"])) <= 1e-6 * float(g["sf"])
    cls = torch.from_numpy(g["cls"])
    gt = torch.from_numpy(g["gt_cls"])
    seg = sum(losses.motion_seg_loss(cls[b:b + 1], gt[b]) for b in range(batch)) / batch
    assert abs(float(seg) - float(g["seg"])) <= 2e-6 * float(g["seg"])
    if batch == 1:   # one call over the batch == the reference's single call
        assert abs(float(losses.motion_seg_loss(cls, gt)) - float(g["seg"])) <= 2e-6 * float(g["seg"])
        total, items = losses.track_4d_loss(pc1_wrap, cls, torch.from_numpy(g["gt_flow"]), gt)
        assert abs(float(total) - (0.5 * float(g["sf"]) + float(g["seg"]))) <= 1e-5


def test_seg_loss_empty_class_and_pretrain():
    cls = torch.tensor([[0.2, 0.7, 0.9]], requires_grad=True)
    gt = torch.tensor([False, False, False])
    assert torch.isnan(losses.motion_seg_loss(cls, gt))                  # reference: mean over an empty selection
    total, items = losses.track_4d_loss(torch.zeros(1, 3, 3) + cls.unsqueeze(1), cls, torch.ones(1, 3, 3), gt, pretrain=True)
    assert float(items["SegLoss"]) == 0.0 and float(total) == 0.0        # reference replaces the NaN term by 0 (loss.py:18)
    total.backward()
    assert torch.isfinite(cls.grad).all()


def test_affinity_ground_truth_and_loss():
    gt = losses.affinity_ground_truth([3, 5, 9], [5, 3])
    assert gt.tolist() == [0, 1, 1, 0, 0, 0]                             # row-major (prev x curr), loss.py:49-66
    aff = torch.tensor([0.1, 0.8, 0.7, 0.2, 0.3, 0.4])
    ref = torch.nn.functional.binary_cross_entropy(aff, gt)
    assert float(losses.affinity_loss(aff, gt)) == float(ref)
    assert float(losses.affinity_loss(torch.zeros(0), torch.zeros(0))) == 0.0


def test_seg_loss_gradient_is_finite_at_saturated_probabilities():
    """A sigmoid output reaches exactly 0.0 / 1.0 in fp32; nn.BCELoss (the reference's choice, loss.py:131) keeps the gradient
    finite there, and so must the mirror (a -log1p(-p) chain would produce 0 * inf = NaN and poison the whole step)."""
    cls = torch.tensor([[1.0, 0.0, 0.3, 1.0]], requires_grad=True)
    gt = torch.tensor([True, False, True, False])
    loss = losses.motion_seg_loss(cls, gt)
    loss.backward()
    assert torch.isfinite(loss) and torch.isfinite(cls.grad).all()
    ref_pos = torch.nn.BCELoss()(cls.detach()[:, gt], torch.ones(1, 2))
    ref_neg = torch.nn.BCELoss()(cls.detach()[:, ~gt], torch.zeros(1, 2))
    assert abs(float(loss) - float(0.4 * ref_pos + 0.6 * ref_neg)) < 1e-5
