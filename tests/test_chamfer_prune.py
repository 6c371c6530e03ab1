"""Spatially pruned exact Chamfer scan (csrc/nn_prune.cuh, GENPC_CHAMFER_PRUNE=1): bit-exact against the oracle and against the
exhaustive kernels on shapes / distributions chosen to break a pruning bound (exact ties, degenerate boxes, far clusters,
ragged block tails), WITH evidence (block counters) that the pruned kernels did the work and skipped most blocks; inputs it
must hand back to the exhaustive kernels (NaN / Inf / huge coordinates) are checked too."""
import numpy as np
import pytest
import torch

import oracle
from util import lattice_cloud, rand_cloud, shape_cloud

pytestmark = pytest.mark.gpu


def run(a, b, dev, prune=True):
    """prune: True / "1" = one-CTA sort + single-level scan (nn_prune.cuh), "2" = multi-CTA sort + two-level scan
    (nn_grid.cuh), "0" = exhaustive kernels even where the large-cloud default would prune, False = library default."""
    from genpc_b200 import _lib
    from genpc_b200.loss_functions import chamfer_3DDist

    stats = torch.zeros(4, dtype=torch.int32, device=dev)
    L = _lib.lib()
    L.genpc_chamfer_prune_stats(_lib.ptr(stats))
    knob = "1" if prune is True else (None if prune is False else prune)
    try:
        with _lib.tunable(GENPC_CHAMFER_PRUNE=knob):
            d1, d2, i1, i2 = chamfer_3DDist()(torch.from_numpy(a).to(dev), torch.from_numpy(b).to(dev))
            torch.cuda.synchronize()
    finally:
        L.genpc_chamfer_prune_stats(None)
    return (d1.cpu().numpy(), d2.cpu().numpy(), i1.cpu().numpy(), i2.cpu().numpy()), stats.cpu().numpy()


def same(got, exp, tag=""):
    for g, e, name in zip(got, exp, ("dist1", "dist2", "idx1", "idx2")):
        assert np.array_equal(g.view(np.int32), e.view(np.int32)), f"{tag} {name}: {(g.view(np.int32) != e.view(np.int32)).sum()} of {g.size} differ"


def blocks_total(B, N, M):
    g = lambda n: (n + 31) // 32
    k = lambda n: (n + 63) // 64
    return B * (g(N) * k(M) + g(M) * k(N))


@pytest.mark.parametrize("B,N,M", [(1, 64, 512), (2, 100, 777), (1, 513, 640), (3, 2049, 2047), (1, 5000, 5000), (2, 700, 4100),
                                   (4, 2048, 16384), (1, 16384, 16384), (1, 32768, 1000), (5, 1025, 4097)])
def test_pruned_scan_bit_exact_on_awkward_shapes(cuda, B, N, M):
    a, b = shape_cloud(B * 3 + N, B, N), shape_cloud(M + 7, B, M)
    got, st = run(a, b, cuda)
    if B * N * M <= 6e7:
        same(got, oracle.chamfer_forward(a, b), f"{B}x{N}x{M} vs oracle")
    ref, st0 = run(a, b, cuda, prune=False)
    same(got, ref, f"{B}x{N}x{M} vs exhaustive")
    assert st[2] > 0 and st0[2] == 0, (st, st0)           # the pruned kernels ran (query groups counted), and only on request


@pytest.mark.parametrize("layout", ["1", "2", "3", "4", "8", "m2", "m3", "m4"])
def test_sort_cluster_layouts_bit_exact(cuda, layout):
    """nn_bin_sort_kernel<CS>: a cloud sorted by a thread-block cluster of CS CTAs (sibling histograms and boxes through distributed
    shared memory), every CS and the mixed layout ("m": the smaller side's clouds take one CTA each, CTAs that round the grid up to
    whole clusters idle) -- the order inside a cell changes, the nearest neighbours must not.  Ragged sizes: slabs that only some
    CTAs of a cluster own, clouds smaller than one slab, 5 clouds (a mixed grid that is not a multiple of the cluster size)."""
    from genpc_b200 import _lib

    for (B, N, M) in [(5, 2048, 16384), (3, 1500, 9000), (2, 5000, 4100), (1, 32768, 777), (7, 200, 2300)]:
        a, b = shape_cloud(B + N, B, N), shape_cloud(M + 3, B, M)
        with _lib.tunable(GENPC_SORT_CLUSTER=layout):
            got, st = run(a, b, cuda)
        ref, _ = run(a, b, cuda, prune="0")
        same(got, ref, f"layout {layout}, {B}x{N}x{M} vs exhaustive")
        if B * N * M <= 6e7:
            same(got, oracle.chamfer_forward(a, b), f"layout {layout}, {B}x{N}x{M} vs oracle")
        assert st[2] > 0, st


def test_c2_full_batch_pruned(cuda):
    """BASELINE C2 (B=32, 2048 x 16384, the bench's batch): bit-exact against the oracle, and most blocks are never visited."""
    from genpc_b200.synthetic import pcn_batch

    part, comp = pcn_batch(0, 32, 2048, 16384)
    got, st = run(part, comp, cuda)
    same(got, oracle.chamfer_forward(part, comp), "C2")
    tot = blocks_total(32, 2048, 16384)
    print("C2 pruned scan: %d of %d (group, block) pairs visited = %.1f %%, %d of %d groups took the tie pass" %
          (st[0], tot, 100.0 * st[0] / tot, st[1], st[2]))
    assert st[2] == 32 * (2048 // 32 + 16384 // 32) and st[0] < 0.5 * tot


def test_exact_ties_and_degenerate_geometry(cuda):
    rng = np.random.default_rng(5)
    cases = {
        "lattice": (lattice_cloud(1, 2, 4096, 16), lattice_cloud(2, 2, 2048, 16)),
        "identical": (rand_cloud(3, 2, 3000), rand_cloud(3, 2, 3000)),
        "duplicates": (rand_cloud(4, 1, 2048), np.concatenate([rand_cloud(5, 1, 1024)] * 2, 1)),
        "planar": (np.concatenate([rng.random((1, 4000, 2)), np.zeros((1, 4000, 1))], 2).astype(np.float32),
                   np.concatenate([rng.random((1, 3000, 2)), np.zeros((1, 3000, 1))], 2).astype(np.float32)),
        "one_point_cloud_repeated": (np.full((1, 256, 3), 0.25, np.float32), rand_cloud(6, 1, 512)),
        "tiny_extent": ((rng.random((1, 4096, 3)) * 1e-4 + 0.5).astype(np.float32), (rng.random((1, 4096, 3)) * 1e-4 + 0.5).astype(np.float32)),
        "far_apart": (rand_cloud(7, 1, 2048), (rand_cloud(8, 1, 2048) + 1000).astype(np.float32)),
        "two_clusters": (np.concatenate([rng.random((1, 1024, 3)), rng.random((1, 1024, 3)) + 40], 1).astype(np.float32),
                         np.concatenate([rng.random((1, 512, 3)), rng.random((1, 3584, 3)) + 40], 1).astype(np.float32)),
        "lidar_scale": ((rng.random((1, 8192, 3)) * 200 - 100).astype(np.float32), (rng.random((1, 8192, 3)) * 200 - 100).astype(np.float32)),
    }
    for name, (a, b) in cases.items():
        a, b = np.ascontiguousarray(a, np.float32), np.ascontiguousarray(b, np.float32)
        got, st = run(a, b, cuda)
        same(got, oracle.chamfer_forward(a, b), name)
        # clouds whose bounding boxes are far apart cannot be pruned (every group would open every block): the sort kernel's
        # overlap test hands them to the exhaustive kernels
        assert (st[2] > 0) == (name not in ("far_apart", "one_point_cloud_repeated")), (name, st)   # (a blob against a spread-out cloud, too)


def test_default_takes_the_pruned_scan_from_2_30_evaluations(cuda):
    """Library default (no knob): 16 x 8192 x 8192 = 2^30 evaluations is pruned, 8 x 8192 x 8192 is not; both bit-identical to the
    exhaustive kernels."""
    for B, expect in ((16, True), (8, False)):
        a, b = shape_cloud(31 + B, B, 8192), shape_cloud(77 + B, B, 8192)
        got, st = run(a, b, cuda, prune=False)
        ref, st0 = run(a, b, cuda, prune="0")
        same(got, ref, f"B={B}")
        assert (st[2] > 0) == expect and st0[2] == 0, (B, st, st0)


def test_out_of_range_inputs_go_to_the_exhaustive_kernels(cuda):
    a, b = rand_cloud(9, 2, 2048), rand_cloud(10, 2, 4096)
    for name, v in (("nan", np.nan), ("inf", np.inf), ("huge", 1e20)):
        a2 = a.copy()
        a2[1, 77, 1] = v
        got, st = run(a2, b, cuda)
        ref, _ = run(a2, b, cuda, prune=False)
        same(got, ref, name)
        assert st[2] == 0, (name, st)                    # the sort kernel's range check sent the work to the exhaustive scan
    got, st = run(a, b, cuda)                            # and the flag does not stick
    same(got, oracle.chamfer_forward(a, b), "clean again")
    assert st[2] > 0


def test_fused_loss_step_with_the_pruned_scan(cuda):
    """Completionloss (fused epilogue: loss, zero fill, re-arm) on top of the pruned scan, several steps on one cached workspace."""
    from genpc_b200 import _lib
    from genpc_b200.utils.loss_util import Completionloss

    a, b = shape_cloud(11, 4, 2048), shape_cloud(12, 4, 8192)
    res = {}
    for prune in (False, True):
        with _lib.tunable(GENPC_CHAMFER_PRUNE="1" if prune else None):
            crit = Completionloss("cd_l2")
            out = []
            for step in range(3):
                ta = torch.from_numpy(a + 0.01 * step).to(cuda).requires_grad_(True)
                tb = torch.from_numpy(b).to(cuda).requires_grad_(True)
                loss = crit.get_loss(ta, tb)
                loss.backward()
                out.append((float(loss), ta.grad.cpu().numpy(), tb.grad.cpu().numpy()))
            res[prune] = out
    for (l0, ga0, gb0), (l1, ga1, gb1) in zip(res[False], res[True]):
        assert l0 == l1                                          # deterministic reduction of identical distances
        assert np.allclose(ga0, ga1, rtol=1e-5, atol=1e-9) and np.allclose(gb0, gb1, rtol=1e-5, atol=1e-9)   # float atomics


# ---- large-cloud form (nn_grid.cuh) ----

@pytest.mark.parametrize("B,N,M", [(1, 64, 512), (2, 100, 777), (3, 2049, 2047), (2, 700, 4100), (4, 2048, 16384), (1, 5000, 40000),
                                   (1, 71372, 16384)])
def test_grid_scan_bit_exact(cuda, B, N, M):
    a, b = shape_cloud(B * 5 + N, B, N), shape_cloud(M + 3, B, M)
    got, st = run(a, b, cuda, prune="2")
    if B * N * M <= 6e7:
        same(got, oracle.chamfer_forward(a, b), f"{B}x{N}x{M} vs oracle")
    ref, st0 = run(a, b, cuda, prune="0")
    same(got, ref, f"{B}x{N}x{M} vs exhaustive")
    assert st[2] > 0 and st0[2] == 0, (st, st0)


def test_grid_scan_ties_degenerate_and_out_of_range(cuda):
    rng = np.random.default_rng(15)
    cases = {
        "lattice": (lattice_cloud(1, 1, 8192, 16), lattice_cloud(2, 1, 4096, 16)),
        "identical": (rand_cloud(3, 1, 9000), rand_cloud(3, 1, 9000)),
        "planar": (np.concatenate([rng.random((1, 6000, 2)), np.zeros((1, 6000, 1))], 2).astype(np.float32),
                   np.concatenate([rng.random((1, 5000, 2)), np.zeros((1, 5000, 1))], 2).astype(np.float32)),
        "far_apart": (rand_cloud(7, 1, 5000), (rand_cloud(8, 1, 5000) + 1000).astype(np.float32)),
        "two_clusters": (np.concatenate([rng.random((1, 3000, 3)), rng.random((1, 3000, 3)) + 40], 1).astype(np.float32),
                         np.concatenate([rng.random((1, 500, 3)), rng.random((1, 6000, 3)) + 40], 1).astype(np.float32)),
        "all_points_equal": (np.full((1, 4500, 3), 0.5, np.float32), np.full((1, 4200, 3), 0.5, np.float32)),
    }
    for name, (a, b) in cases.items():
        a, b = np.ascontiguousarray(a, np.float32), np.ascontiguousarray(b, np.float32)
        got, st = run(a, b, cuda, prune="2")
        same(got, oracle.chamfer_forward(a, b), name)
        assert st[2] > 0, name
    a, b = rand_cloud(9, 1, 6000), rand_cloud(10, 1, 7000)
    a2 = a.copy()
    a2[0, 4321, 2] = np.nan
    got, st = run(a2, b, cuda, prune="2")
    ref, _ = run(a2, b, cuda, prune="0")
    same(got, ref, "nan")
    assert st[2] == 0


def test_large_clouds_take_the_grid_scan_by_default(cuda):
    """300 000 x 250 000 LiDAR-like scene pair (the C5 generator): the library default is the pruned scan, bit-identical to the
    exhaustive kernels (forced with the knob), visiting a small fraction of the blocks."""
    from genpc_b200.synthetic import lidar_scene_pair

    a, b = lidar_scene_pair(300000, 0)
    a, b = a.numpy()[None, :300000], b.numpy()[None, :250000]
    got, st = run(np.ascontiguousarray(a), np.ascontiguousarray(b), cuda, prune=False)
    ref, st0 = run(np.ascontiguousarray(a), np.ascontiguousarray(b), cuda, prune="0")
    same(got, ref, "300k x 250k")
    tot = blocks_total(1, 300000, 250000)
    print("grid scan: %d of %d (group, block) pairs visited = %.3f %%, %d superblocks opened by %d groups" %
          (st[0], tot, 100.0 * st[0] / tot, st[3], st[2]))
    assert st[2] > 0 and st0[2] == 0 and st[0] < 0.02 * tot


def test_probe_hands_non_overlapping_large_clouds_to_the_exhaustive_kernels(cuda):
    """Default path at 70 000 x 70 000 (>= 2^32 evaluations): overlapping clouds take the pruned scan; clouds that do not
    overlap would make every query group open every block -- the sampled probe sees that and the exhaustive kernels run."""
    a = rand_cloud(21, 1, 70000)
    b_near = (a[:, ::-1] + 0.003 * rand_cloud(22, 1, 70000)).astype(np.float32).copy()
    b_far = (rand_cloud(23, 1, 70000) * np.array([1, 1, 0.01], np.float32) + np.array([0, 0, 5], np.float32)).astype(np.float32)
    got, st = run(a, b_near, cuda, prune=False)
    ref, _ = run(a, b_near, cuda, prune="0")
    same(got, ref, "overlapping")
    assert st[2] > 0, st
    got, st = run(a, b_far, cuda, prune=False)
    ref, _ = run(a, b_far, cuda, prune="0")
    same(got, ref, "apart")
    assert st[2] == 0, st


def test_host_fed_chunks_take_the_pruned_scan_and_a_late_bad_chunk_starts_over(cuda):
    """Host-fed C2 batch (forward_from_host, 6 chunks): sort + pruned scan are queued per chunk behind the chunk's copy event.
    A NaN in the LAST chunk is only seen after earlier chunks have stored exact words: the words are re-armed and the
    exhaustive kernels redo the batch -- same outputs as the device-resident exhaustive call."""
    from genpc_b200 import _lib
    from genpc_b200.loss_functions import chamfer_3DDist

    a, b = shape_cloud(41, 32, 2048), shape_cloud(42, 32, 16384)

    def host(x, y):
        stats = torch.zeros(4, dtype=torch.int32, device=cuda)
        _lib.lib().genpc_chamfer_prune_stats(_lib.ptr(stats))
        try:
            with _lib.tunable(GENPC_HOST_PRUNE="1"):   # the default for a batch of this size; forced here
                out = chamfer_3DDist().forward_from_host(torch.from_numpy(x).pin_memory(), torch.from_numpy(y).pin_memory(), device=cuda, chunks=6)
                torch.cuda.synchronize()
        finally:
            _lib.lib().genpc_chamfer_prune_stats(None)
        return tuple(t.detach().cpu().numpy() for t in out[:4]), stats.cpu().numpy()

    got, st = host(a, b)
    same(got, oracle.chamfer_forward(a, b), "host-fed C2")
    assert st[2] == 32 * (2048 // 32 + 16384 // 32), st
    a2 = a.copy()
    a2[31, 100, 0] = np.nan
    got, st = host(a2, b)
    ref, _ = run(a2, b, cuda, prune="0")
    same(got, ref, "host-fed, NaN in the last chunk")
    got, st = host(a, b)                                     # the flag does not stick
    same(got, oracle.chamfer_forward(a, b), "host-fed C2 again")
    assert st[2] > 0
