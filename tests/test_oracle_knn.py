"""CPU: the oracle's k-NN statistic (fusion tail, reg_xyz.py:219) against a float64 brute force."""
import numpy as np


def test_oracle_knn_mean_distance_vs_float64_bruteforce():
    """CPU: the oracle's k-NN statistic against a float64 brute force (values agree to fp32 rounding), both
    self-inclusion modes, a cloud smaller than k, duplicates."""
    import oracle

    rng = np.random.default_rng(3)
    x = rng.random((700, 3), dtype=np.float32)
    x[10] = x[11]                                              # an exact duplicate: distance 0 besides self
    d = np.sqrt(((x[:, None, :].astype(np.float64) - x[None, :, :]) ** 2).sum(-1))
    d.sort(1)
    for k in (1, 5, 20, 32):
        assert np.abs(oracle.knn_mean_distance(x, k, True) - d[:, :k].mean(1)).max() < 1e-6
        assert np.abs(oracle.knn_mean_distance(x, k, False) - d[:, 1:k + 1].mean(1)).max() < 1e-6
    small = x[:7]
    ds = np.sort(np.sqrt(((small[:, None, :].astype(np.float64) - small[None, :, :]) ** 2).sum(-1)), 1)
    assert np.abs(oracle.knn_mean_distance(small, 20, True) - ds.mean(1)).max() < 1e-6      # averages what exists
    assert (oracle.knn_mean_distance(x[:1], 5, False) == -1).all()                           # nothing found
    m = oracle.knn_mean_distance(x, 20, True)
    keep = oracle.statistical_outlier_mask(m, 2.5)
    assert keep.sum() > 0.9 * len(x) and not keep[m.argmax()]
