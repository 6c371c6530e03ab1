"""CPU oracle of the Chamfer-driven registration loop -- TEST INFRASTRUCTURE ONLY.

Follows optim_registration/diff_obj_pose.py: ObjectPoseOptim.forward (:408-436, transform :419-423),
compute_loss_function's Chamfer term (:323-334: cd = CDp-L1(pts->ref) + 0.5*CDp-L1(ref->pts), weight 3.0) and
the optimiser loop (:518-576: Adam, lr groups {lr, 0.2 lr, 0.1 lr}, init scale 0.75, Ry(k*90 deg) multi-starts,
iters+1 steps, "best" = final params of the start with the lowest ever-seen loss).  The mask / Pulsar-render
terms are out of scope (BASELINE.json).  The ortho regulariser 0.001*|R R^T - I| (:543-545) has an analytically
zero gradient through the Gram-Schmidt parametrisation and is omitted.

Machinery = the reference's own: torch autograd for the gradient and torch.optim.Adam for the update.  The NN
indices come from the C oracle on the fp32-transformed cloud (oracle.transform + oracle.nn_distance), the
differentiable loss is then rebuilt in float64 from those indices, gradients are cast to the fp32 parameters.
pytorch3d is not vendored: rotation_6d_to_matrix is restated from its published definition (parity unpinned).
"""
import math

import numpy as np
import torch

from . import nn_distance, transform


def rotation_6d_to_matrix(d6):
    a1, a2 = d6[..., :3], d6[..., 3:]
    b1 = torch.nn.functional.normalize(a1, dim=-1)
    b2 = a2 - (b1 * a2).sum(-1, keepdim=True) * b1
    b2 = torch.nn.functional.normalize(b2, dim=-1)
    b3 = torch.cross(b1, b2, dim=-1)
    return torch.stack((b1, b2, b3), dim=-2)


def get_init_rot_y(angle_deg):
    """get_init_rot('y', angle) (:470-493): first two rows of Ry(angle)."""
    a = math.radians(angle_deg)
    return np.array([math.cos(a), 0.0, math.sin(a), 0.0, 1.0, 0.0], np.float32)


def init_params(start=0):
    p = np.zeros(10, np.float32)
    p[:6] = get_init_rot_y(start * 90)
    p[9] = np.float32(math.log(0.75))
    return p


def loss_and_grad(params, V, center, ref, w_fwd=1.0, w_inv=0.5, cd_weight=3.0):
    """-> (loss float64, grad[10] float64, (idxA, idxB))."""
    pts32 = transform(V, center, params)
    dA, iA = nn_distance(pts32[None], ref[None])
    dB, iB = nn_distance(ref[None], pts32[None])
    iA, iB = torch.from_numpy(iA[0].astype(np.int64)), torch.from_numpy(iB[0].astype(np.int64))
    p = torch.tensor(params.astype(np.float64), requires_grad=True)
    Vt, ct, rt = (torch.from_numpy(np.asarray(x, np.float64)) for x in (V, center, ref))
    R = rotation_6d_to_matrix(p[:6])
    s = torch.exp(p[9])
    local = (Vt - ct) * s
    pts = (R @ local.T).T + ct + p[6:9]
    da = ((pts - rt[iA]) ** 2).sum(-1)
    db = ((rt - pts[iB]) ** 2).sum(-1)
    loss = cd_weight * (w_fwd * torch.sqrt(da).mean() + w_inv * torch.sqrt(db).mean())
    loss.backward()
    return float(loss.detach()), p.grad.numpy().copy(), (iA.numpy(), iB.numpy())


def run(V, ref, iters, lr=0.01, start=0, w_fwd=1.0, w_inv=0.5, cd_weight=3.0, center=None):
    """`iters` Adam steps -> (params history [iters+1,10] fp32, loss history [iters] fp64)."""
    V, ref = np.asarray(V, np.float32), np.asarray(ref, np.float32)
    center = V.mean(0).astype(np.float32) if center is None else np.asarray(center, np.float32)
    par = torch.tensor(init_params(start))
    rot, trans, ls = (par[:6].clone().requires_grad_(True), par[6:9].clone().requires_grad_(True),
                      par[9:].clone().requires_grad_(True))
    opt = torch.optim.Adam([{"params": [rot], "lr": lr}, {"params": [trans], "lr": lr * 0.2},
                            {"params": [ls], "lr": lr * 0.1}])
    hist = [np.concatenate([rot.detach().numpy(), trans.detach().numpy(), ls.detach().numpy()])]
    losses = []
    for _ in range(iters):
        cur = hist[-1].astype(np.float32)
        loss, g, _ = loss_and_grad(cur, V, center, ref, w_fwd, w_inv, cd_weight)
        opt.zero_grad()
        rot.grad = torch.tensor(g[:6], dtype=torch.float32)
        trans.grad = torch.tensor(g[6:9], dtype=torch.float32)
        ls.grad = torch.tensor(g[9:], dtype=torch.float32)
        opt.step()
        losses.append(loss)
        hist.append(np.concatenate([rot.detach().numpy(), trans.detach().numpy(), ls.detach().numpy()]))
    return np.stack(hist).astype(np.float32), np.array(losses)
