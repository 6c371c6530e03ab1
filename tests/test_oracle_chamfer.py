"""CPU: the oracle's Chamfer restatement against float64 brute force and the reference's invariants
(Chamfer(X,X) => d=0, idx=arange; lowest index on ties; dist recomputed from idx)."""
import numpy as np
import pytest

import oracle
from util import lattice_cloud, rand_cloud


def brute(a, b):
    D = ((a[:, :, None, :].astype(np.float64) - b[:, None, :, :].astype(np.float64)) ** 2).sum(-1)
    return D


@pytest.mark.parametrize("B,N,M", [(1, 1, 1), (2, 37, 513), (3, 600, 64), (1, 1025, 1023)])
def test_forward_matches_float64(B, N, M):
    a, b = rand_cloud(1, B, N), rand_cloud(2, B, M)
    d1, d2, i1, i2 = oracle.chamfer_forward(a, b)
    D = brute(a, b)
    # fp32 rounding can flip near-ties vs float64, so compare the distance at the returned index
    np.testing.assert_allclose(np.take_along_axis(D, i1[..., None].astype(np.int64), 2)[..., 0], D.min(2), rtol=1e-5, atol=1e-9)
    np.testing.assert_allclose(d1, D.min(2), rtol=1e-5, atol=1e-9)
    np.testing.assert_allclose(d2, D.min(1), rtol=1e-5, atol=1e-9)
    assert i1.dtype == np.int32 and i2.dtype == np.int32


def test_self_distance_is_zero_and_identity():
    a = rand_cloud(3, 2, 777)
    d1, d2, i1, i2 = oracle.chamfer_forward(a, a)
    assert (d1 == 0).all() and (d2 == 0).all()
    assert (i1 == np.arange(777)).all() and (i2 == np.arange(777)).all()


def test_lowest_index_wins_ties():
    a, b = lattice_cloud(4, 2, 300), lattice_cloud(5, 2, 2000)
    d1, d2, i1, i2 = oracle.chamfer_forward(a, b)
    D = brute(a, b)  # exact in float64 AND float32 for small integers
    assert (i1 == D.argmin(2)).all()  # numpy argmin returns the first (lowest) index
    assert (i2 == D.argmin(1)).all()
    assert (d1 == D.min(2).astype(np.float32)).all()


def test_distance_is_fma_order():
    # d must equal fma(dz,dz,fma(dx,dx,dy*dy)) evaluated in fp32, emulated here through float64
    a, b = rand_cloud(6, 1, 50), rand_cloud(7, 1, 60)
    d1, _, i1, _ = oracle.chamfer_forward(a, b)
    t = b[0][i1[0]]
    dx, dy, dz = [(t[:, k] - a[0][:, k]).astype(np.float32) for k in range(3)]
    f64 = np.float64
    s = (f64(dy) * f64(dy)).astype(np.float32)                  # rounded product
    s = (f64(dx) * f64(dx) + f64(s)).astype(np.float32)          # fma: exact product + add, one rounding
    s = (f64(dz) * f64(dz) + f64(s)).astype(np.float32)
    assert (s == d1[0]).all()


def test_backward_matches_float64():
    a, b = rand_cloud(8, 2, 90), rand_cloud(9, 2, 150)
    d1, d2, i1, i2 = oracle.chamfer_forward(a, b)
    rng = np.random.default_rng(0)
    g1 = rng.standard_normal((2, 90)).astype(np.float32)
    g2 = rng.standard_normal((2, 150)).astype(np.float32)
    gx1, gx2 = oracle.chamfer_backward(a, b, g1, g2, i1, i2)
    e1 = np.zeros((2, 90, 3))
    e2 = np.zeros((2, 150, 3))
    for bb in range(2):
        for j in range(90):
            t = 2 * g1[bb, j] * (a[bb, j].astype(np.float64) - b[bb, i1[bb, j]])
            e1[bb, j] += t
            e2[bb, i1[bb, j]] -= t
        for k in range(150):
            t = 2 * g2[bb, k] * (b[bb, k].astype(np.float64) - a[bb, i2[bb, k]])
            e2[bb, k] += t
            e1[bb, i2[bb, k]] -= t
    np.testing.assert_allclose(gx1, e1, rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(gx2, e2, rtol=1e-5, atol=1e-6)
