"""Tensor-core filter path of the Chamfer forward (csrc/nn_tc.cuh): bit-exact against the oracle on every shape / point
distribution below, WITH evidence (filter statistics) that the tcgen05 path did the work rather than the FP32 fall-back;
the error assumption behind the filter margin pinned against float64."""
import numpy as np
import pytest
import torch

import oracle
from util import lattice_cloud, rand_cloud, shape_cloud

pytestmark = pytest.mark.gpu


def run(a, b, dev, force=True, limit=None):
    from genpc_b200 import _lib
    from genpc_b200.loss_functions import chamfer_3DDist

    stats = torch.zeros(4, dtype=torch.int32, device=dev)
    L = _lib.lib()
    L.genpc_chamfer_tc_stats(_lib.ptr(stats))
    try:
        with _lib.tunable(GENPC_CHAMFER_TC="1" if force else None, GENPC_TC_LIMIT=limit):
            d1, d2, i1, i2 = chamfer_3DDist()(torch.from_numpy(a).to(dev), torch.from_numpy(b).to(dev))
            torch.cuda.synchronize()
    finally:
        L.genpc_chamfer_tc_stats(None)
    return (d1.cpu().numpy(), d2.cpu().numpy(), i1.cpu().numpy(), i2.cpu().numpy()), stats.cpu().numpy()


def check(got, a, b, tag=""):
    exp = oracle.chamfer_forward(a, b)
    for g, e, name in zip(got, exp, ("dist1", "dist2", "idx1", "idx2")):
        assert np.array_equal(g.view(np.int32), e.view(np.int32)), f"{tag} {name}: {(g != e).sum()} of {g.size} differ"


@pytest.mark.parametrize("B,N,M", [(1, 640, 1), (1, 513, 31), (2, 1000, 257), (1, 4096, 2048), (3, 2049, 2047), (1, 5000, 5000),
                                   (2, 700, 4100), (4, 2048, 16384), (1, 20000, 300), (1, 1025, 4097)])
def test_forced_filter_bit_exact_on_awkward_shapes(cuda, B, N, M):
    """Row / column counts that are not multiples of 128 / 256 / 32, more than one column span (M > 2048 on the column side),
    items of one to eight row blocks, tiny column clouds -- shape clouds in [-0.5, 0.5]^3."""
    a, b = shape_cloud(B * 3 + N, B, N), shape_cloud(M + 7, B, M)
    got, st = run(a, b, cuda)
    check(got, a, b, f"{B}x{N}x{M}")
    assert st[3] > 0 and st[2] == 0, st               # the filter ran (items > 0), nothing degenerate


def test_c2_full_batch_through_the_filter(cuda):
    """BASELINE C2 (B=32, 2048 x 16384, the bench's batch) through the filter: bit-exact; whole-tile fall-backs stay rare.
    (The filter is opt-in: it measured slower than the FP32 scan on this shape)"""
    from genpc_b200.synthetic import pcn_batch

    part, comp = pcn_batch(0, 32, 2048, 16384)
    got, st = run(part, comp, cuda, force=True)
    check(got, part, comp, "C2")
    n_points = 32 * 16384 + 32 * 2048 * 16        # (point, item) resolutions: every row once, every column once per row tile
    assert st[3] >= 148 and st[2] == 0, st
    assert st[1] < 0.03 * n_points, f"whole-tile exact scans: {st[1]} of {n_points} point resolutions"
    print("C2 filter statistics: runner-up re-evaluations %d, whole-tile scans %d, items %d" % (st[0], st[1], st[3]))


@pytest.mark.parametrize("kind", ["lattice", "duplicates", "clustered", "tiny_scale", "coincident"])
def test_forced_filter_on_tie_heavy_and_near_degenerate_data(cuda, kind):
    """Data built to defeat the filter: exact ties everywhere (lattice, duplicated points, identical clouds), points packed
    far below the margin (clusters of 1e-4 extent around unit-scale centres, a whole cloud of 1e-3 extent): every point ends
    up in the exact re-evaluation / whole-tile scan, the result must still be the oracle's bit for bit."""
    B, N, M = 2, 3000, 2500
    if kind == "lattice":
        a, b = lattice_cloud(1, B, N, side=5) / 5 - 0.4, lattice_cloud(2, B, M, side=5) / 5 - 0.4
        a, b = a.astype(np.float32), b.astype(np.float32)
    elif kind == "duplicates":
        a, b = rand_cloud(3, B, N, 1.0, -0.5), rand_cloud(4, B, M, 1.0, -0.5)
        a[:, N // 2:] = a[:, :N - N // 2]
        b[:, M // 3:2 * (M // 3)] = b[:, :M // 3]
    elif kind == "clustered":
        rng = np.random.default_rng(5)
        ca, cb = rng.random((B, 20, 3), dtype=np.float32) - 0.5, rng.random((B, 20, 3), dtype=np.float32) - 0.5
        a = (ca[:, rng.integers(0, 20, N)] + 1e-4 * rng.standard_normal((B, N, 3))).astype(np.float32)
        b = (ca[:, rng.integers(0, 20, M)] + 1e-4 * rng.standard_normal((B, M, 3))).astype(np.float32)
    elif kind == "tiny_scale":
        a, b = rand_cloud(6, B, N, 1e-3, 0.3), rand_cloud(7, B, M, 1e-3, 0.3)
    else:
        a = rand_cloud(8, B, N, 1.0, -0.5)
        b = a[:, :M].copy()
    got, st = run(a, b, cuda)
    check(got, a, b, kind)
    assert st[3] > 0, st


def test_out_of_range_coordinates_fall_back_to_fp32_and_poison_is_contained(cuda):
    """The device-side precheck: coordinates beyond the filter's range (LiDAR-scale scenes) or NaN / inf send the launch to
    the FP32 kernel (no filter items); inside the range, huge-norm points make their item degenerate (exact scans)."""
    a, b = rand_cloud(11, 2, 3000, 80.0, -40.0), rand_cloud(12, 2, 2600, 80.0, -40.0)
    got, st = run(a, b, cuda)
    check(got, a, b, "lidar scale")
    assert st[3] == 0, st                                       # precheck said no: nn_tc_kernel returned at once
    a, b = rand_cloud(13, 2, 3000, 1.0, -0.5), rand_cloud(14, 2, 2600, 1.0, -0.5)
    a[0, 17] = np.nan
    got, st = run(a, b, cuda)
    assert st[3] == 0 and int(got[2].min()) >= 0 and int(got[2].max()) < 2600 and int(got[3].min()) >= 0 and int(got[3].max()) < 3000
    # a raised limit lets 1e15-scale points into the filter: their items are degenerate, results still exact
    a, b = rand_cloud(15, 1, 2000, 1.0, -0.5), rand_cloud(16, 1, 1500, 1.0, -0.5)
    a[0, 5] = [3e14, -1e15, 2e14]
    b[0, 1400] = [1e15, 1e15, -1e15]
    got, st = run(a, b, cuda, limit="1e20")
    check(got, a, b, "degenerate items")
    assert st[3] > 0 and st[2] > 0, st


def test_error_assumption_of_the_filter_margin(cuda):
    """nn_tc.cuh's margin assumes |e - |x-y|^2| <= TC_KAPPA (|x|+|y|)^2 with TC_KAPPA = 2.5e-6, of which 1.9e-6 is an ASSUMED
    bound on the tensor core's fp32 accumulation (its order is undocumented).  Measured here on 600 random 128 x 256 tiles
    (2e7 pairs) at scales 1e-3 .. 1e3, centred and offset, against float64: the worst ratio must stay below a QUARTER of
    the budget -- a hardware / driver change that eats the safety factor fails this test before it can flip an index."""
    from genpc_b200 import _lib

    L = _lib.lib()
    rng = np.random.default_rng(0)
    worst = 0.0
    e = torch.empty(128, 256, device=cuda)
    for trial in range(600):
        scale = 10.0 ** rng.uniform(-3, 3)
        off = rng.standard_normal(3) * scale * rng.choice([0.0, 1.0, 30.0])
        r = (rng.standard_normal((128, 3)) * scale + off).astype(np.float32)
        c = (rng.standard_normal((256, 3)) * scale + off).astype(np.float32)
        if trial % 7 == 0:
            c[:128] = r + (rng.standard_normal((128, 3)) * scale * 1e-4).astype(np.float32)   # near-coincident pairs
        tr, tc = torch.from_numpy(r).to(cuda), torch.from_numpy(c).to(cuda)
        _lib.check(L.genpc_tc_probe(_lib.ptr(tr), _lib.ptr(tc), _lib.ptr(e), _lib.current_stream(cuda)), "probe")
        got = e.cpu().numpy().astype(np.float64)
        rd, cd = r.astype(np.float64), c.astype(np.float64)
        d = ((rd[:, None, :] - cd[None]) ** 2).sum(-1)
        P = (np.linalg.norm(rd, axis=1)[:, None] + np.linalg.norm(cd, axis=1)[None]) ** 2
        worst = max(worst, float((np.abs(got - d) / P).max()))
    print(f"worst |e - d| / (|x|+|y|)^2 over 2e7 pairs: {worst:.3e}  (budget 2.5e-6)")
    assert worst <= 2.5e-6 / 4, worst
