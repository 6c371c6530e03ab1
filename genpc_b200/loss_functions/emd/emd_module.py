"""Drop-in for the reference's loss_functions/emd/emd_module.py (:29-95): EMD approximation (auction algorithm).

Input  xyz1, xyz2: [#batch, #points, 3] (xyz1 predicted, xyz2 ground truth), same size, #points % 256 == 0,
       #batch <= 512; eps balances error rate and convergence speed; iters = number of auction iterations.
Output dist [#batch, #points] (sqrt(dist) -> L2 distance), assignment [#batch, #points] (index of the matched
       ground-truth point; an approximation, not guaranteed to be a bijection).  Gradient for xyz1 only.
Differences from the reference: tensors stay on the device they arrive on (the reference hard-codes
device="cuda", :41-54, which breaks under DataParallel on any GPU but 0) and work goes to the current stream.
"""
import torch
from torch import nn
from torch.autograd import Function

from ... import emd


class emdFunction(Function):
    @staticmethod
    def forward(ctx, xyz1, xyz2, eps, iters):
        batchsize, n, _ = xyz1.size()
        _, m, _ = xyz2.size()

        assert n == m
        assert xyz1.size()[0] == xyz2.size()[0]
        assert n % 256 == 0
        assert batchsize <= 512

        xyz1 = xyz1.contiguous().float()
        xyz2 = xyz2.contiguous().float()
        dev = xyz1.device
        dist = torch.zeros(batchsize, n, device=dev)
        assignment = torch.full((batchsize, n), -1, device=dev, dtype=torch.int32)
        assignment_inv = torch.full((batchsize, m), -1, device=dev, dtype=torch.int32)
        price = torch.zeros(batchsize, m, device=dev)
        bid = torch.zeros(batchsize, n, device=dev, dtype=torch.int32)
        bid_increments = torch.zeros(batchsize, n, device=dev)
        max_increments = torch.zeros(batchsize, m, device=dev)
        unass_idx = torch.zeros(batchsize * n, device=dev, dtype=torch.int32)
        max_idx = torch.zeros(batchsize * m, device=dev, dtype=torch.int32)
        unass_cnt = torch.zeros(512, dtype=torch.int32, device=dev)
        unass_cnt_sum = torch.zeros(512, dtype=torch.int32, device=dev)
        cnt_tmp = torch.zeros(512, dtype=torch.int32, device=dev)

        emd.forward(xyz1, xyz2, dist, assignment, price, assignment_inv, bid, bid_increments, max_increments,
                    unass_idx, unass_cnt, unass_cnt_sum, cnt_tmp, max_idx, eps, iters)

        ctx.save_for_backward(xyz1, xyz2, assignment)
        ctx.mark_non_differentiable(assignment)
        return dist, assignment

    @staticmethod
    def backward(ctx, graddist, gradidx):
        xyz1, xyz2, assignment = ctx.saved_tensors
        graddist = graddist.contiguous()
        gradxyz1 = torch.zeros_like(xyz1)
        gradxyz2 = torch.zeros_like(xyz2)
        emd.backward(xyz1, xyz2, gradxyz1, graddist, assignment)
        return gradxyz1, gradxyz2, None, None


class emdModule(nn.Module):
    def __init__(self):
        super(emdModule, self).__init__()

    def forward(self, input1, input2, eps, iters):
        return emdFunction.apply(input1, input2, eps, iters)
