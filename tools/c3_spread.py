"""C3 at the BASELINE length (500 Adam iterations): how far do trajectories of the SAME optimisation drift apart?
  ours        fused kernels (csrc/register.cu), fp32 + deterministic double reductions
  oracle      oracle/registration.py: C-oracle NN indices, float64 autograd gradient, torch.optim.Adam on fp32 parameters
  reference   the reference's machinery on the GPU (its unmodified Chamfer extension inside torch autograd + Adam), run TWICE:
              its backward sums with unordered float atomics, so run-to-run spread is the yardstick for any tolerance.
Prints max |delta param| and relative loss difference at checkpoints.   gpurun -- python tools/c3_spread.py"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import oracle  # noqa: E402
from oracle import registration as OR  # noqa: E402
from genpc_b200.optim_registration.diff_obj_pose import RegistrationBatch, rotation_6d_to_matrix  # noqa: E402
from test_registration import make_pair  # noqa: E402

dev = torch.device("cuda:0")
ITERS = int(os.environ.get("C3_ITERS", 500))
CKPT = [c for c in (1, 10, 25, 50, 100, 250, 500) if c <= ITERS]


def run_reference(V, Rf, center, p0, lr, iters, ext):
    class RefChamfer(torch.autograd.Function):
        @staticmethod
        def forward(ctx, a, b):
            B, n, _ = a.shape
            m = b.shape[1]
            d1 = torch.zeros(B, n, device=a.device); d2 = torch.zeros(B, m, device=a.device)
            i1 = torch.zeros(B, n, dtype=torch.int32, device=a.device); i2 = torch.zeros(B, m, dtype=torch.int32, device=a.device)
            ext.forward(a, b, d1, d2, i1, i2)
            ctx.save_for_backward(a, b, i1, i2)
            return d1, d2, i1, i2

        @staticmethod
        def backward(ctx, g1, g2, _a, _b):
            a, b, i1, i2 = ctx.saved_tensors
            ga, gb = torch.zeros_like(a), torch.zeros_like(b)
            ext.backward(a, b, ga, gb, g1.contiguous(), g2.contiguous(), i1, i2)
            return ga, gb

    def pl1(p, q):
        d1, _, _, _ = RefChamfer.apply(p.contiguous(), q.contiguous())
        return torch.sqrt(d1).mean()

    rot = p0[:6].clone().requires_grad_(True); trans = p0[6:9].clone().requires_grad_(True); ls = p0[9:].clone().requires_grad_(True)
    opt = torch.optim.Adam([{"params": [rot], "lr": lr}, {"params": [trans], "lr": lr * 0.2}, {"params": [ls], "lr": lr * 0.1}])
    hist, losses = [], []
    for _ in range(iters):
        opt.zero_grad()
        R = rotation_6d_to_matrix(rot[None])[0]
        pts = (R @ ((V - center) * torch.exp(ls)[0]).T).T + center + trans
        loss = 3.0 * (pl1(pts[None], Rf[None]) + 0.5 * pl1(Rf[None], pts[None]))
        loss.backward()
        opt.step()
        losses.append(float(loss.detach()))
        hist.append(torch.cat([rot, trans, ls]).detach().cpu().numpy().copy())
    return np.stack(hist), np.array(losses)


def main():
    out = {"iters": ITERS, "cases": []}
    ext = oracle.load_ref_ext("chamfer_3D")
    for seed, nc, nr in [(11, 2048, 1500), (12, 4096, 3000)]:
        comp, part, _ = make_pair(seed, nc, nr)
        V, Rf = torch.from_numpy(comp).to(dev), torch.from_numpy(part).to(dev)
        rb = RegistrationBatch(V[None], Rf[None], n_starts=1, lr=0.01, max_iters=ITERS)
        center, p0 = rb.center[0].clone(), rb.params[0].clone()
        ours = []
        for _ in range(ITERS):
            rb.run(1)
            ours.append(rb.params[0].cpu().numpy().copy())
        ours = np.stack(ours)
        ours_l = rb.losses()[0].cpu().numpy()
        eh, el = OR.run(comp, part, ITERS, lr=0.01, center=center.cpu().numpy())
        eh = eh[1:]
        case = {"seed": seed, "Nc": nc, "Nr": nr, "final_loss_ours": float(ours_l[-1]), "final_loss_oracle": float(el[-1])}
        runs = {}
        if ext is not None:
            r1, l1 = run_reference(V, Rf, center, p0, 0.01, ITERS, ext)
            r2, l2 = run_reference(V, Rf, center, p0, 0.01, ITERS, ext)
            runs = {"ref1": (r1, l1), "ref2": (r2, l2)}
        rows = {}
        for c in CKPT:
            row = {"ours_vs_oracle_param": float(np.abs(ours[c - 1] - eh[c - 1]).max()),
                   "ours_vs_oracle_loss_rel": float(abs(ours_l[c - 1] - el[c - 1]) / abs(el[c - 1]))}
            if runs:
                row["ref_vs_ref_param"] = float(np.abs(r1[c - 1] - r2[c - 1]).max())
                row["ref_vs_ref_loss_rel"] = float(abs(l1[c - 1] - l2[c - 1]) / abs(l1[c - 1]))
                row["ours_vs_ref_param"] = float(np.abs(ours[c - 1] - r1[c - 1]).max())
                row["ours_vs_ref_loss_rel"] = float(abs(ours_l[c - 1] - l1[c - 1]) / abs(l1[c - 1]))
                row["oracle_vs_ref_param"] = float(np.abs(eh[c - 1] - r1[c - 1]).max())
            rows[str(c)] = row
        case["checkpoints"] = rows
        # worst over the whole trajectory
        case["max_over_trajectory"] = {"ours_vs_oracle_param": float(np.abs(ours - eh).max()),
                                       "ours_vs_oracle_loss_rel": float((np.abs(ours_l - el) / np.abs(el)).max())}
        if runs:
            case["max_over_trajectory"].update(ref_vs_ref_param=float(np.abs(r1 - r2).max()),
                                               ref_vs_ref_loss_rel=float((np.abs(l1 - l2) / np.abs(l1)).max()),
                                               ours_vs_ref_param=float(np.abs(ours - r1).max()),
                                               ours_vs_ref_loss_rel=float((np.abs(ours_l - l1) / np.abs(l1)).max()))
        out["cases"].append(case)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
