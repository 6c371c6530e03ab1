"""CPU: property-based checks of the oracle itself (hypothesis): the invariants the reference's semantics imply,
on arbitrary small clouds -- so the checker is not only pinned on a few golden vectors."""
import numpy as np
from hypothesis import given, settings
from hypothesis import strategies as st
from hypothesis.extra import numpy as hnp

import oracle

coords = st.floats(min_value=-4.0, max_value=4.0, allow_nan=False, width=32)


def clouds(min_pts=1, max_pts=40):
    return hnp.arrays(np.float32, st.tuples(st.integers(1, 3), st.integers(min_pts, max_pts), st.just(3)), elements=coords)


@settings(max_examples=60, deadline=None)
@given(clouds(), clouds())
def test_chamfer_is_symmetric_and_self_consistent(a, b):
    B = min(a.shape[0], b.shape[0])
    a, b = np.ascontiguousarray(a[:B]), np.ascontiguousarray(b[:B])
    d1, d2, i1, i2 = oracle.chamfer_forward(a, b)
    e2, e1, j2, j1 = oracle.chamfer_forward(b, a)             # swapping the inputs swaps the outputs, bit for bit
    assert np.array_equal(d1, e1) and np.array_equal(d2, e2) and np.array_equal(i1, j1) and np.array_equal(i2, j2)
    assert (d1 >= 0).all() and (i1 >= 0).all() and (i1 < b.shape[1]).all() and (i2 < a.shape[1]).all()
    for bb in range(B):                                       # the reported distance is the distance to the reported index
        t = b[bb][i1[bb]]
        dx, dy, dz = [(t[:, k] - a[bb][:, k]).astype(np.float32) for k in range(3)]
        s = (dy.astype(np.float64) * dy).astype(np.float32)
        s = (dx.astype(np.float64) * dx + s).astype(np.float32)
        s = (dz.astype(np.float64) * dz + s).astype(np.float32)
        assert np.array_equal(s, d1[bb])
        # no target is strictly closer, and among equally close ones the lowest index was reported
        D = ((a[bb][:, None, :].astype(np.float64) - b[bb][None].astype(np.float64)) ** 2).sum(-1)
        assert (D.min(1) >= d1[bb].astype(np.float64) * (1 - 1e-6) - 1e-12).all()


@settings(max_examples=40, deadline=None)
@given(clouds(min_pts=2, max_pts=60), st.integers(0, 1))
def test_fps_prefix_and_monotone(a, start):
    N = a.shape[1]
    K = max(1, N // 2)
    idx, seq = oracle.fps(a, K, start, True)
    idx2 = oracle.fps(a, max(1, K // 2), start)
    assert np.array_equal(idx[:, :idx2.shape[1]], idx2)        # the first picks do not depend on how many follow
    assert (idx[:, 0] == start).all() and np.isinf(seq[:, 0]).all()
    assert (np.diff(seq[:, 1:], axis=1) <= 0).all()
    for b in range(a.shape[0]):                                # a repeated index only once everything left is a duplicate
        if len(set(idx[b])) < K:
            assert seq[b, len(set(idx[b]))] == 0 or len(np.unique(a[b], axis=0)) < K


@settings(max_examples=25, deadline=None)
@given(st.integers(0, 2**31 - 1), st.sampled_from([256, 512]), st.sampled_from([0.005, 0.05]), st.integers(1, 60))
def test_emd_outputs_are_consistent(seed, n, eps, iters):
    rng = np.random.default_rng(seed)
    x1, x2 = rng.random((1, n, 3), dtype=np.float32), rng.random((1, n, 3), dtype=np.float32)
    d, a, stt = oracle.emd_forward(x1, x2, eps, iters, return_state=True)
    assert (a >= 0).all() and (a < n).all()                    # the last iteration assigns everyone (emd_cuda.cu:201-205)
    g = x2[0][a[0]]
    dx, dy, dz = [(x1[0][:, k] - g[:, k]).astype(np.float32) for k in range(3)]
    s = (dy.astype(np.float64) * dy).astype(np.float32)
    s = (dx.astype(np.float64) * dx + s).astype(np.float32)
    s = (dz.astype(np.float64) * dz + s).astype(np.float32)
    assert np.array_equal(s, d[0])                             # CalcDist (:217-226)
    assert (stt["price"] >= 0).all()
    if iters >= 2:                                             # prices only ever rise by at least eps per win
        won = stt["price"][0] > 0
        assert (stt["price"][0][won] >= np.float32(eps) * (1 - 1e-6)).all()


@settings(max_examples=30, deadline=None)
@given(clouds(min_pts=3, max_pts=50), st.integers(8, 40), st.integers(1, 3))
def test_zbuffer_owner_is_nearest_of_its_pixel(a, res, ps):
    from genpc_b200 import depth as D

    pts = np.ascontiguousarray(a[0]) * np.float32(0.1)
    eyes = D.fibonacci_sphere(3, 1.6)
    cams = np.stack([D.make_camera(e, np.zeros(3), D.calculate_up_vector(e, np.zeros(3)), 0.857, res, res) for e in eyes])
    ndc, uv, bounds = oracle.project_uv(cams, pts, True, 0.15)
    if not np.isfinite(uv).all():                              # degenerate bbox (all points project to one pixel column)
        return
    zb = oracle.zbuffer(uv, ndc, res, ps)
    zb1 = oracle.zbuffer(uv, ndc, res, 1)
    filled, filled1 = zb != np.uint64(0xFFFFFFFFFFFFFFFF), zb1 != np.uint64(0xFFFFFFFFFFFFFFFF)
    assert (filled | ~filled1).all()                           # a bigger splat only adds pixels
    own = (zb & np.uint64(0xFFFFFFFF)).astype(np.int64)
    assert (own[filled] < pts.shape[0]).all()
    for v in range(3):                                         # point_size 1: the owner is the nearest point of that pixel
        col = np.clip((uv[v, :, 0] * np.float32(res)).astype(np.int64), 0, res - 1)
        row = np.clip((uv[v, :, 1] * np.float32(res)).astype(np.int64), 0, res - 1)
        o1 = (zb1[v] & np.uint64(0xFFFFFFFF)).astype(np.int64)
        for p in range(pts.shape[0]):
            w = o1[res - 1 - row[p], col[p]]
            assert ndc[v, w, 2] < ndc[v, p, 2] or (ndc[v, w, 2] == ndc[v, p, 2] and w <= p)
