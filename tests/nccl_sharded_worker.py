"""Worker of tests/test_sharded.py::test_nccl_two_ranks_*: one process per GPU (torchrun), real NCCL.
Sharded Chamfer (forward + autograd form) on every rank against the single-GPU kernels, plus the batch-sharded helpers."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    dev = torch.device(f"cuda:{local}")
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    from genpc_b200.loss_functions import chamfer_3DDist
    from genpc_b200.sharded import data_parallel_emd, shard_range, sharded_chamfer_3DDist, sharded_chamfer_forward
    from util import lattice_cloud, rand_cloud

    res = {}
    cases = {"rand_300k_x_200k": (rand_cloud(1, 1, 300_000), rand_cloud(2, 1, 200_000)),
             "lattice_ties": (lattice_cloud(3, 1, 20_000, side=9), lattice_cloud(4, 1, 33_333, side=9)),
             "ragged_small": (rand_cloud(5, 1, 1000), rand_cloud(6, 1, 77))}
    for name, (a, b) in cases.items():
        ta, tb = torch.from_numpy(a).to(dev), torch.from_numpy(b).to(dev)
        got = sharded_chamfer_forward(ta, tb)
        exp = chamfer_3DDist()(ta, tb)
        res[name] = all(torch.equal(g, e) for g, e in zip(got, exp))
    # autograd form: gradients equal to the single-GPU module's on every rank
    a, b = cases["rand_300k_x_200k"]
    grads = []
    for mod in (sharded_chamfer_3DDist(), chamfer_3DDist()):
        xa = torch.from_numpy(a[:, :50_000]).to(dev).requires_grad_(True)
        xb = torch.from_numpy(b[:, :40_000]).to(dev).requires_grad_(True)
        d1, d2, _, _ = mod(xa, xb)
        (d1.sqrt().mean() + d2.mean()).backward()
        grads.append((xa.grad, xb.grad))
    res["autograd"] = all(bool((g0 - g1).abs().max() <= 1e-5 * g1.abs().max()) for g0, g1 in zip(grads[0], grads[1]))
    # batch-sharded EMD (the reference's nn.DataParallel(emdModule), utils/loss_util.py:12): every rank owns a batch slice,
    # the gathered result equals the single-GPU call
    g = torch.Generator().manual_seed(0)
    x1, x2 = torch.rand(6, 1024, 3, generator=g).to(dev), torch.rand(6, 1024, 3, generator=g).to(dev)
    from genpc_b200.loss_functions import emdModule
    d_all, a_all = data_parallel_emd(x1, x2, 0.005, 50)
    d_one, a_one = emdModule()(x1, x2, 0.005, 50)
    res["emd_batch_sharded"] = bool(torch.equal(d_all, d_one) and torch.equal(a_all, a_one))
    lo, hi = shard_range(6, rank, world)
    res["emd_slice"] = [lo, hi]
    ok = torch.tensor([1.0 if all(v for k, v in res.items() if isinstance(v, bool)) else 0.0], device=dev)
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    res["all_ranks_ok"] = bool(ok.item() == 1.0)
    print(f"RANK{rank} " + json.dumps(res), flush=True)
    dist.destroy_process_group()
    sys.exit(0 if res["all_ranks_ok"] else 1)


if __name__ == "__main__":
    main()
