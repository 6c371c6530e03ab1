"""Drop-in for the reference's utils/loss_util.py (:8-53): Completionloss with the same methods and reductions.

The reference wraps emdModule in nn.DataParallel (:12), which only scatters the batch over the visible GPUs of
ONE process; here the EMD module is called directly (one process per GPU is the multi-GPU model of this build,
see genpc_b200.sharded), results are identical.
"""
import torch

from ..loss_functions import chamfer_3DDist, emdModule


class Completionloss:
    def __init__(self, loss_func='cd_l1'):
        self.loss_func = loss_func
        self.chamfer_dist = chamfer_3DDist()
        self.EMD = emdModule()

        if loss_func == 'cd_l1':
            self.metric = self.chamfer_l1
            self.partial_matching = self.chamfer_partial_l1
        elif loss_func == 'cd_l2':
            self.metric = self.chamfer_l2
            self.partial_matching = self.chamfer_partial_l2
        elif loss_func == 'emd':
            self.metric = self.emd_loss
        else:
            raise Exception('loss function {} not supported yet!'.format(loss_func))

    def chamfer_l1(self, p1, p2):
        d1, d2, _, _ = self.chamfer_dist(p1, p2)
        d1 = torch.mean(torch.sqrt(d1))
        d2 = torch.mean(torch.sqrt(d2))
        return (d1 + d2) / 2

    def chamfer_l2(self, p1, p2):
        d1, d2, _, _ = self.chamfer_dist(p1, p2)
        return torch.mean(d1) + torch.mean(d2)

    def chamfer_partial_l1(self, pcd1, pcd2):
        d1, d2, _, _ = self.chamfer_dist(pcd1, pcd2)
        d1 = torch.mean(torch.sqrt(d1))
        return d1

    def chamfer_partial_l2(self, pcd1, pcd2):
        d1, d2, _, _ = self.chamfer_dist(pcd1, pcd2)
        d1 = torch.mean(d1)
        return d1

    def emd_loss(self, p1, p2):
        d1, _ = self.EMD(p1, p2, eps=0.005, iters=50)
        d = torch.sqrt(d1).mean(1).mean()
        return d

    def get_loss(self, gen, gt):
        loss = self.metric(gen, gt)
        return loss
