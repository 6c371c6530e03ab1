"""IO / format glue (reference utils/dataUtils.py:174-189, 561-581 and Open3D's voxel_down_sample as
the reference calls it, reg_xyz.py:154-155): restated in numpy here and compared with genpc_b200.utils.dataUtils."""
import numpy as np
import torch

from genpc_b200.utils.dataUtils import normalize_numpy, voxel_down_sample


def voxel_down_sample_numpy(p, voxel):
    """Open3D semantics: voxel index = floor((p - (min_bound - voxel/2)) / voxel); output = mean of each occupied voxel."""
    lo = p.min(0) - voxel * 0.5
    key = np.floor((p - lo) / voxel).astype(np.int64)
    out = {}
    for k, q in zip(map(tuple, key), p.astype(np.float64)):
        s = out.setdefault(k, [np.zeros(3), 0])
        s[0] += q
        s[1] += 1
    keys = sorted(out)
    return np.array([out[k][0] / out[k][1] for k in keys]), keys


def test_voxel_down_sample_matches_numpy_restatement():
    rng = np.random.default_rng(0)
    p = rng.random((4000, 3)).astype(np.float32) * np.array([1.0, 0.5, 0.25], np.float32)
    for voxel in (0.02, 0.03, 0.11):
        got = voxel_down_sample(torch.from_numpy(p), voxel).numpy()
        want, keys = voxel_down_sample_numpy(p, np.float32(voxel))
        assert got.shape == want.shape
        assert np.allclose(got, want, atol=2e-6)          # fp32 sums of <= a few hundred points vs float64
        # every output point lies inside its voxel
        lo = p.min(0) - np.float32(voxel) * 0.5
        assert np.array_equal(np.floor((got - lo) / voxel).astype(np.int64), np.array(keys))


def test_voxel_down_sample_degenerate():
    p = torch.tensor([[0.1, 0.2, 0.3]] * 7)
    assert voxel_down_sample(p, 0.05).shape == (1, 3)
    q = torch.rand(100, 3)
    assert voxel_down_sample(q, 10.0).shape == (1, 3) and torch.allclose(voxel_down_sample(q, 10.0)[0], q.mean(0), atol=1e-6)
    assert voxel_down_sample(q, 1e-4).shape[0] == 100     # one point per voxel: a permutation of the input
    

def test_normalize_numpy_is_the_reference_formula():
    rng = np.random.default_rng(1)
    xyz = rng.standard_normal((500, 3)) * np.array([3.0, 1.0, 0.2]) + 5.0
    for rg in (0.5, 1.0):
        out, c, s = normalize_numpy(xyz, range=rg)
        vmin, vmax = xyz.min(0), xyz.max(0)                # utils/dataUtils.py:561-581
        assert np.allclose(c, (vmin + vmax) / 2) and np.isclose(s, (vmax - vmin).max())
        assert np.allclose(out, (xyz - c) / s * (rg / 0.5))
        assert np.isclose(out.max() - out.min(), 2 * rg) or np.isclose((out.max(0) - out.min(0)).max(), 2 * rg)
