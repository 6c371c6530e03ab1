"""Target-sharded Chamfer: host-side logic (slicing, all-reduce-MIN of packed words, unpack) on CPU with gloo,
world_size 2; the CUDA partial scan + merge against the oracle and the single-GPU kernel on the GPU."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle
from util import lattice_cloud, rand_cloud


def test_shard_range_covers_everything():
    from genpc_b200.sharded import shard_range

    for n in (0, 1, 7, 1000, 1000003):
        for w in (1, 2, 3, 8):
            r = [shard_range(n, k, w) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[k][1] == r[k + 1][0] for k in range(w - 1))
            assert max(b - a for a, b in r) - min(b - a for a, b in r) <= 1


def pack_np(d, i):
    return (d.view(np.uint32).astype(np.int64) << 32) | i.astype(np.int64)


def _worker(rank, world, port, a, b, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from genpc_b200.sharded import allreduce_min_packed, shard_range, unpack_packed

    lo, hi = shard_range(b.shape[1], rank, world)
    if hi > lo:
        d, i = oracle.nn_distance(a, np.ascontiguousarray(b[:, lo:hi]))   # the partial scan, done by the checker on CPU
        packed = torch.from_numpy(pack_np(d, i + lo))
    else:
        packed = torch.full(a.shape[:2], -1, dtype=torch.int64)           # empty shard: all-ones words
    allreduce_min_packed(packed)
    dd, ii = unpack_packed(packed)
    out[rank] = (dd.numpy().copy(), ii.numpy().copy())
    dist.destroy_process_group()


@pytest.mark.parametrize("case", ["rand", "lattice", "tiny"])
def test_gloo_world2_merge_equals_full_scan(case):
    if case == "rand":
        a, b = rand_cloud(1, 2, 300), rand_cloud(2, 2, 1001)
    elif case == "lattice":
        a, b = lattice_cloud(3, 1, 400), lattice_cloud(4, 1, 900)     # ties must resolve to the lowest GLOBAL index
    else:
        a, b = rand_cloud(5, 1, 5), rand_cloud(6, 1, 1)               # rank 1 gets an empty shard
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, port, a, b, out), nprocs=2, join=True)
    ed, ei = oracle.nn_distance(a, b)
    for r in range(2):
        d, i = out[r]
        assert np.array_equal(d.view(np.int32), ed.view(np.int32)) and np.array_equal(i, ei)


@pytest.mark.gpu
def test_partial_scans_merge_to_the_single_gpu_result(cuda):
    """Emulate 3 ranks on one GPU: per-shard partial scans + min-merge == chamfer_3DDist == oracle."""
    from genpc_b200.loss_functions import chamfer_3DDist
    from genpc_b200.sharded import nn_partial_packed, nn_unpack, shard_range, sharded_chamfer_forward

    a, b = rand_cloud(7, 1, 50000), lattice_cloud(8, 1, 30011, side=40)
    ta, tb = torch.from_numpy(a).to(cuda), torch.from_numpy(b).to(cuda)
    parts = []
    for r in range(3):
        lo, hi = shard_range(b.shape[1], r, 3)
        parts.append(nn_partial_packed(ta, tb[:, lo:hi], lo))
    merged = torch.stack(parts).min(0).values            # what the all-reduce-MIN computes
    d, i = nn_unpack(merged)
    d1, d2, i1, i2 = chamfer_3DDist()(ta, tb)
    assert torch.equal(d, d1) and torch.equal(i, i1)
    ed, ei = oracle.nn_distance(a, b)
    assert np.array_equal(d.cpu().numpy(), ed) and np.array_equal(i.cpu().numpy(), ei)
    s = sharded_chamfer_forward(ta, tb)                    # world size 1 path (row-sharded symmetric scan)
    for x, y in zip(s, (d1, d2, i1, i2)):
        assert torch.equal(x, y)


@pytest.mark.gpu
def test_row_sharded_symmetric_partials_merge_exactly(cuda):
    """Emulate 3 ranks of the row-sharded symmetric scheme on one GPU (min-merge == all-reduce-MIN)."""
    import ctypes

    from genpc_b200 import _lib
    from genpc_b200.loss_functions import chamfer_3DDist
    from genpc_b200.sharded import EMPTY, nn_unpack, shard_range_aligned

    a, b = lattice_cloud(9, 1, 20011, side=30), rand_cloud(10, 1, 33333) * 30
    ta, tb = torch.from_numpy(a).to(cuda), torch.from_numpy(b).to(cuda)
    N, M = a.shape[1], b.shape[1]
    L = _lib.lib()
    parts = []
    for r in range(3):
        lo, hi = shard_range_aligned(N, r, 3)
        assert lo % 128 == 0
        packed = torch.full((N + M,), EMPTY, dtype=torch.int64, device=cuda)
        rc = L.genpc_chamfer_sym_partial(_lib.ptr(ta[0, lo:hi]), _lib.ptr(tb), ctypes.c_void_p(packed.data_ptr() + lo * 8),
                                         ctypes.c_void_p(packed.data_ptr() + N * 8), 1, hi - lo, M, lo, 0,
                                         _lib.current_stream(cuda))
        assert rc == 0
        parts.append(torch.where(packed == EMPTY, torch.iinfo(torch.int64).max, packed))
    merged = torch.stack(parts).min(0).values
    d1, i1 = nn_unpack(merged[:N].contiguous())
    d2 = torch.empty(M, device=cuda); i2 = torch.empty(M, dtype=torch.int32, device=cuda)
    rc = L.genpc_chamfer_sym_fixup(_lib.ptr(ta), _lib.ptr(tb), ctypes.c_void_p(merged.data_ptr() + N * 8), 1, N, M,
                                   _lib.ptr(d2), _lib.ptr(i2), _lib.current_stream(cuda))
    assert rc == 0
    e = chamfer_3DDist()(ta, tb)
    assert torch.equal(d1, e[0][0]) and torch.equal(i1, e[2][0]) and torch.equal(d2, e[1][0]) and torch.equal(i2, e[3][0])
    ed, ei = oracle.nn_distance(b, a)
    assert np.array_equal(i2.cpu().numpy(), ei[0]) and np.array_equal(d2.cpu().numpy(), ed[0])


@pytest.mark.gpu
def test_sharded_module_is_differentiable_like_chamfer_3DDist(cuda):
    """sharded_chamfer_3DDist (world size 1 here): same outputs as chamfer_3DDist, gradients within the atomics'
    summation-order tolerance (1e-5 relative, both backward passes accumulate with float atomics)."""
    from genpc_b200.loss_functions import chamfer_3DDist
    from genpc_b200.sharded import sharded_chamfer_3DDist

    a, b = rand_cloud(11, 1, 40000), rand_cloud(12, 1, 25000)
    res = []
    for mod in (chamfer_3DDist(), sharded_chamfer_3DDist()):
        ta = torch.from_numpy(a).to(cuda).requires_grad_(True)
        tb = torch.from_numpy(b).to(cuda).requires_grad_(True)
        d1, d2, i1, i2 = mod(ta, tb)
        (d1.sqrt().mean() + 0.5 * d2.mean()).backward()
        res.append((d1.detach(), d2.detach(), i1, i2, ta.grad, tb.grad))
    for x, y in zip(res[0][:4], res[1][:4]):
        assert torch.equal(x, y)
    g1 = (0.5 / np.sqrt(res[0][0].cpu().numpy()) / d1.numel()).astype(np.float32)      # d loss / d dist1
    g2 = np.full(b.shape[:2], 0.5 / d2.numel(), np.float32)
    e1, e2 = oracle.chamfer_backward(a, b, g1, g2, res[0][2].cpu().numpy(), res[0][3].cpu().numpy())
    for k, exp in ((4, e1), (5, e2)):           # both modules against the double-accumulated oracle, 1e-5 of the scale
        scale = np.abs(exp).max()
        for r in res:
            assert np.abs(r[k].cpu().numpy() - exp).max() <= 1e-5 * scale + 1e-12


@pytest.mark.gpu
def test_full_size_c5_million_points_sampled(cuda):
    """BASELINE config C5 at full size (1M x 1M, one rank): the row-sharded symmetric path against the oracle on a
    sample of 400 points per direction (bit-exact dist + idx), and size-independent properties on everything:
    indices in range, every distance equal to the distance recomputed from its index, dist2 consistent with dist1
    (dist2[idx1[j]] <= dist1[j])."""
    from genpc_b200.sharded import sharded_chamfer_forward

    n = 1_000_000
    g = torch.Generator().manual_seed(0)
    a = (torch.rand(1, n, 3, generator=g) * torch.tensor([80.0, 80.0, 3.0])).contiguous()
    b = (a + torch.randn(1, n, 3, generator=g) * 0.05)[:, torch.randperm(n, generator=g)].contiguous()
    ta, tb = a.to(cuda), b.to(cuda)
    d1, d2, i1, i2 = sharded_chamfer_forward(ta, tb)
    assert int(i1.min()) >= 0 and int(i1.max()) < n and int(i2.min()) >= 0 and int(i2.max()) < n
    for q, t, d, i in ((ta, tb, d1, i1), (tb, ta, d2, i2)):
        nn = t[0, i[0].long()]
        dx, dy, dz = (nn - q[0]).double().unbind(-1)
        assert torch.allclose((dx * dx + dy * dy + dz * dz).float(), d[0], rtol=1e-5, atol=1e-12)
    assert bool((d2[0, i1[0].long()] <= d1[0]).all()) and bool((d1[0, i2[0].long()] <= d2[0]).all())
    sel = np.random.default_rng(1).choice(n, 400, replace=False)
    an, bn = a.numpy(), b.numpy()
    ed, ei = oracle.nn_distance(an[:, sel], bn)
    assert np.array_equal(d1[0, sel].cpu().numpy().view(np.int32), ed[0].view(np.int32))
    assert np.array_equal(i1[0, sel].cpu().numpy(), ei[0])
    ed, ei = oracle.nn_distance(bn[:, sel], an)
    assert np.array_equal(d2[0, sel].cpu().numpy().view(np.int32), ed[0].view(np.int32))
    assert np.array_equal(i2[0, sel].cpu().numpy(), ei[0])


@pytest.mark.gpu
def test_nccl_two_ranks_sharded_chamfer_and_batch_sharded_emd(cuda):
    """Real NCCL, one process per GPU (torchrun, 2 ranks): sharded Chamfer forward bit-identical to the single-GPU kernels
    on every rank (random, lattice-tie and ragged clouds), its autograd form within 1e-5, and the batch-sharded EMD helper
    (the reference's nn.DataParallel(emdModule), utils/loss_util.py:12) equal to the single-GPU call.  Skipped on a
    one-GPU box (the driver's round-end suite runs on one GPU; `gpurun --gpus 2` runs it for real)."""
    import subprocess
    import sys

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    worker = os.path.join(os.path.dirname(os.path.abspath(__file__)), "nccl_sharded_worker.py")
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                        "127.0.0.1", "--master-port", str(port), worker], capture_output=True, text=True, timeout=600)
    import re

    # (the two ranks share one pipe: their lines may arrive glued together, so look for the records, not for line starts)
    lines = re.findall(r"RANK\d+ \{[^{}]*\}", p.stdout)
    assert p.returncode == 0 and len(lines) == 2, p.stdout[-2000:] + p.stderr[-2000:]
    for ln in lines:
        assert '"all_ranks_ok": true' in ln, ln


def _dp_worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from genpc_b200.sharded import data_parallel_batch

    g = torch.Generator().manual_seed(0)
    res = {}
    for B in (1, 2, 5, 8):                                  # fewer entries than ranks, uneven and even splits
        x, y = torch.rand(B, 7, 3, generator=g), torch.rand(B, 4, generator=g)
        fn = lambda a, b: ((a * 2).sum(-1), b.cumsum(-1) + a[:, :4, 0])   # noqa: E731
        got = data_parallel_batch(fn, x, y)
        exp = fn(x, y)
        res[B] = all(torch.equal(g_, e_) for g_, e_ in zip(got, exp))
        single = data_parallel_batch(lambda a: a.mean(1), x)
        res[(B, "single")] = torch.equal(single, x.mean(1))
    out[rank] = res
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_gloo_batch_sharding_helper(world):
    """data_parallel_batch (the one-process-per-GPU stand-in for nn.DataParallel, utils/loss_util.py:12): every rank ends up
    with the full-batch result, whatever the split (host logic on CPU, gloo)."""
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_dp_worker, args=(world, port, out), nprocs=world, join=True)
    for r in range(world):
        assert all(out[r].values()), (r, dict(out[r]))
