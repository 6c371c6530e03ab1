"""FPS: oracle invariants on CPU; CUDA kernel bit-exact (indices and distance sequence) against the oracle."""
import numpy as np
import pytest

import oracle
from util import lattice_cloud, rand_cloud, shape_cloud


def test_oracle_fps_invariants():
    a = rand_cloud(0, 2, 1000)
    idx, seq = oracle.fps(a, 200, 0, True)
    assert (idx[:, 0] == 0).all() and np.isinf(seq[:, 0]).all()
    assert (np.diff(seq[:, 1:], axis=1) <= 0).all()          # selected distances never increase
    assert all(len(set(r)) == 200 for r in idx)               # no repeats on distinct points
    # brute-force definition
    p = a[0].astype(np.float64)
    run = np.full(1000, np.inf)
    cur = 0
    for s in range(20):
        assert idx[0, s] == cur
        run = np.minimum(run, ((p - p[cur]) ** 2).sum(1))
        cur = int(np.argmax(run))


def test_oracle_fps_ties_lowest_index():
    a = lattice_cloud(1, 1, 500, side=4)
    idx = oracle.fps(a, 30, 3)
    p = a[0].astype(np.float64)
    run = np.full(500, np.inf)
    cur = 3
    for s in range(30):
        assert idx[0, s] == cur
        run = np.minimum(run, ((p - p[cur]) ** 2).sum(1))  # exact for small integers
        cur = int(np.argmax(run))                            # first max == lowest index


@pytest.mark.gpu
@pytest.mark.parametrize("B,N,K,start", [(1, 1, 1, 0), (2, 100, 100, 5), (3, 1000, 64, 0), (2, 2048, 512, 7),
                                         (1, 16384, 2048, 0), (2, 5000, 300, 4999), (1, 40000, 128, 1)])
def test_fps_gpu_bit_exact(cuda, B, N, K, start):
    import torch

    from genpc_b200.fps import furthest_point_sample

    a = shape_cloud(N, B, N) if N % 2 == 0 else rand_cloud(N, B, N)
    idx, seq = furthest_point_sample(torch.from_numpy(a).to(cuda), K, start, return_seq=True)
    eidx, eseq = oracle.fps(a, K, start, True)
    assert np.array_equal(idx.cpu().numpy(), eidx)
    assert np.array_equal(seq.cpu().numpy().view(np.int32), eseq.view(np.int32))


@pytest.mark.gpu
def test_fps_gpu_ties_and_api(cuda):
    import torch

    from genpc_b200.fps import fps_sampling, furthest_point_sample

    a = lattice_cloud(2, 2, 3000, side=5)
    idx = furthest_point_sample(torch.from_numpy(a).to(cuda), 100, 0)
    assert np.array_equal(idx.cpu().numpy(), oracle.fps(a, 100, 0))
    out = fps_sampling(a[0], 50)           # fpsample-shaped call: numpy in, numpy out
    assert isinstance(out, np.ndarray) and np.array_equal(out.astype(np.int32), oracle.fps(a[:1], 50, 0)[0])


@pytest.mark.gpu
@pytest.mark.parametrize("B,N,K,start", [(1, 16384, 2048, 0), (2, 5000, 300, 4999), (1, 9000, 500, 3), (3, 20000, 64, 1),
                                         (1, 32768, 100, 0), (1, 1500, 200, 7), (1, 71372, 300, 11), (2, 100000, 40, 0),
                                         (1, 147000, 24, 146999)])
def test_fps_cluster_kernel_bit_exact(cuda, B, N, K, start):
    """The thread-block-cluster / DSMEM kernel (forced through GENPC_FPS_MODE) against the oracle, in its 8-CTA form
    (coordinates in registers or shared memory), its 16-CTA form (everything in registers) and with the default choice."""
    import os

    import torch

    from genpc_b200.fps import furthest_point_sample

    a = shape_cloud(N + 1, B, N)
    eidx, eseq = oracle.fps(a, K, start, True)
    from genpc_b200 import _lib

    for c16 in (None, "0", "1"):
        with _lib.tunable(GENPC_FPS_MODE="cluster", GENPC_FPS_CLUSTER16=c16):
            idx, seq = furthest_point_sample(torch.from_numpy(a).to(cuda), K, start, return_seq=True)
            torch.cuda.synchronize()
        assert np.array_equal(idx.cpu().numpy(), eidx), c16
        assert np.array_equal(seq.cpu().numpy().view(np.int32), eseq.view(np.int32)), c16


@pytest.mark.gpu
@pytest.mark.parametrize("B,N,K,start", [(2, 16384, 300, 5), (20, 12000, 50, 3), (1, 8193, 100, 0), (3, 4097, 200, 4096),
                                         (2, 8192, 64, 1), (1, 6001, 6001, 17)])
def test_fps_single_cta_shared_memory_and_l1_forms(cuda, B, N, K, start):
    """One CTA per cloud (batched FPS, forced with GENPC_FPS_MODE=cta) for 4096 < N <= 16384: the shared-memory form
    (LDS.128, the default) and the L1 form (GENPC_FPS_SMEM=0) against the oracle, indices and running distances."""
    import os

    import torch

    from genpc_b200.fps import furthest_point_sample

    a = lattice_cloud(N, B, N, side=12) if N % 2 else shape_cloud(N, B, N)
    eidx, eseq = oracle.fps(a, K, start, True)
    from genpc_b200 import _lib

    for smem in ("1", "0"):
        with _lib.tunable(GENPC_FPS_MODE="cta", GENPC_FPS_SMEM=smem):
            idx, seq = furthest_point_sample(torch.from_numpy(a).to(cuda), K, start, return_seq=True)
            torch.cuda.synchronize()
        assert np.array_equal(idx.cpu().numpy(), eidx), smem
        assert np.array_equal(seq.cpu().numpy().view(np.int32), eseq.view(np.int32)), smem
