"""Generate golden vectors by running the UNMODIFIED reference extensions (oracle/_ref) on a B200.

    gpurun -- python tests/golden/make_golden.py      # writes gpurun_out/golden/*.npz
    cp gpurun_out/golden/*.npz tests/golden/           # commit them

The reference has no CPU path and this container has no GPU, so the vectors can only be produced on the
GPU box; inputs are seeded and small enough to commit.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle  # noqa: E402
from util import lattice_cloud, rand_cloud, shape_cloud  # noqa: E402


def main():
    out = os.path.join(ROOT, "gpurun_out", "golden")
    os.makedirs(out, exist_ok=True)
    dev = torch.device("cuda:0")
    ch = oracle.load_ref_ext("chamfer_3D")
    cases = {"rand": (rand_cloud(11, 2, 700), rand_cloud(12, 2, 1900)),
             "shape": (shape_cloud(13, 1, 2048), shape_cloud(14, 1, 4096)),
             "lattice": (lattice_cloud(15, 2, 500), lattice_cloud(16, 2, 1500)),
             # edge shapes: ragged sizes below / across the reference's 512-target chunks, negative and large coordinates
             "ragged": (rand_cloud(17, 3, 37), rand_cloud(18, 3, 513)),
             "negative": (shape_cloud(19, 1, 1025), shape_cloud(20, 1, 255)),
             "chunked": (rand_cloud(21, 1, 600, 10.0, -5.0), rand_cloud(22, 1, 1300, 10.0, -5.0))}
    for name, (a, b) in cases.items():
        B, N, M = a.shape[0], a.shape[1], b.shape[1]
        ta, tb = torch.from_numpy(a).to(dev), torch.from_numpy(b).to(dev)
        d1 = torch.zeros(B, N, device=dev); d2 = torch.zeros(B, M, device=dev)
        i1 = torch.zeros(B, N, dtype=torch.int32, device=dev); i2 = torch.zeros(B, M, dtype=torch.int32, device=dev)
        ch.forward(ta, tb, d1, d2, i1, i2)
        torch.cuda.synchronize()
        np.savez_compressed(os.path.join(out, f"chamfer_ref_{name}.npz"), xyz1=a, xyz2=b, dist1=d1.cpu().numpy(),
                            dist2=d2.cpu().numpy(), idx1=i1.cpu().numpy(), idx2=i2.cpu().numpy())
    em = oracle.load_ref_ext("emd")
    rng = np.random.default_rng(21)
    for name, (B, n, eps, iters) in {"small": (2, 512, 0.005, 50), "n2048": (1, 2048, 0.005, 50),
                                     "n4096": (1, 4096, 0.002, 100), "centered": (1, 1024, 0.005, 50),
                                     "b3_n256": (3, 256, 0.01, 30)}.items():
        x1 = rng.random((B, n, 3), dtype=np.float32); x2 = rng.random((B, n, 3), dtype=np.float32)
        if name == "centered":   # the metric path feeds [-0.5, 0.5] data un-normalised (main.py:26-33)
            x1, x2 = x1 - np.float32(0.5), x2 - np.float32(0.5)
        runs = []
        for rep in range(3):  # the reference is racy (GetMax): record its own run-to-run spread
            t1, t2 = torch.from_numpy(x1).to(dev), torch.from_numpy(x2).to(dev)
            dist = torch.zeros(B, n, device=dev)
            asg = torch.zeros(B, n, device=dev, dtype=torch.int32) - 1
            asg_inv = torch.zeros(B, n, device=dev, dtype=torch.int32) - 1
            price = torch.zeros(B, n, device=dev)
            bid = torch.zeros(B, n, device=dev, dtype=torch.int32)
            binc = torch.zeros(B, n, device=dev)
            minc = torch.zeros(B, n, device=dev)
            uidx = torch.zeros(B * n, device=dev, dtype=torch.int32)
            midx = torch.zeros(B * n, device=dev, dtype=torch.int32)
            ucnt = torch.zeros(512, dtype=torch.int32, device=dev)
            ucs = torch.zeros(512, dtype=torch.int32, device=dev)
            ctmp = torch.zeros(512, dtype=torch.int32, device=dev)
            em.forward(t1, t2, dist, asg, price, asg_inv, bid, binc, minc, uidx, ucnt, ucs, ctmp, midx, eps, iters)
            torch.cuda.synchronize()
            runs.append((dist.cpu().numpy(), asg.cpu().numpy(), price.cpu().numpy()))
        same = all(np.array_equal(runs[0][1], r[1]) for r in runs[1:])
        print(f"emd {name}: reference run-to-run identical assignment: {same}; "
              f"cost {[float(np.sqrt(r[0]).mean()) for r in runs]}")
        np.savez_compressed(os.path.join(out, f"emd_ref_{name}.npz"), xyz1=x1, xyz2=x2, eps=eps, iters=iters,
                            dist=runs[0][0], assignment=runs[0][1], price=runs[0][2],
                            reproducible=np.array(same))


if __name__ == "__main__":
    main()
