"""EMD approximation (auction algorithm) -- drop-in for the reference's loss_functions/emd/emd_module.py (:29-95).

    dist, assignment = emdModule()(xyz1, xyz2, eps, iters)

xyz1 (prediction) and xyz2 (ground truth) are [batch, points, 3] with equal point counts, points % 256 == 0 and
batch <= 512 (checked with the reference's asserts, :36-39).  `dist` [batch, points] holds squared matched
distances (sqrt -> L2), `assignment` [batch, points] the matched ground-truth index (an approximation, not
necessarily a bijection).  Only xyz1 receives a gradient (:83-87).

Unlike the reference, which hard-codes device="cuda" for every scratch tensor (:41-54), everything lives on the
device the inputs arrive on, and the single persistent kernel runs on the current stream.
"""
import torch
from torch import nn
from torch.autograd import Function

from ... import emd

_COUNTERS = 512  # length of the reference's three counter tensors (:52-54)


def _scratch(batch, n, device):
    """The caller-allocated state of emd.forward with the reference's initial values (:43-54)."""
    f = dict(device=device, dtype=torch.float32)
    i = dict(device=device, dtype=torch.int32)
    return {
        "dist": torch.zeros(batch, n, **f),
        "assignment": torch.full((batch, n), -1, **i),
        "price": torch.zeros(batch, n, **f),
        "assignment_inv": torch.full((batch, n), -1, **i),
        "bid": torch.zeros(batch, n, **i),
        "bid_increments": torch.zeros(batch, n, **f),
        "max_increments": torch.zeros(batch, n, **f),
        "unass_idx": torch.zeros(batch * n, **i),
        "unass_cnt": torch.zeros(_COUNTERS, **i),
        "unass_cnt_sum": torch.zeros(_COUNTERS, **i),
        "cnt_tmp": torch.zeros(_COUNTERS, **i),
        "max_idx": torch.zeros(batch * n, **i),
    }


class emdFunction(Function):
    @staticmethod
    def forward(ctx, xyz1, xyz2, eps, iters):
        batch, n, _ = xyz1.size()
        assert n == xyz2.size()[1]
        assert batch == xyz2.size()[0]
        assert n % 256 == 0
        assert batch <= 512
        pred = xyz1.contiguous().float()
        truth = xyz2.contiguous().float()
        st = _scratch(batch, n, pred.device)
        emd.forward(pred, truth, *st.values(), eps, iters)
        ctx.save_for_backward(pred, truth, st["assignment"])
        ctx.mark_non_differentiable(st["assignment"])
        return st["dist"], st["assignment"]

    @staticmethod
    def backward(ctx, graddist, gradidx):
        pred, truth, assignment = ctx.saved_tensors
        grad_pred = torch.zeros_like(pred)
        emd.backward(pred, truth, grad_pred, graddist.contiguous(), assignment)
        return grad_pred, torch.zeros_like(truth), None, None


class emdModule(nn.Module):
    def forward(self, input1, input2, eps, iters):
        return emdFunction.apply(input1, input2, eps, iters)
