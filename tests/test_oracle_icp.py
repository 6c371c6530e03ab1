"""ICP oracle (oracle/icp.py): Kabsch/SVD self-checks on CPU; csrc/icp.cu (Horn's closed form) against
it on the GPU, step by step and over whole registrations."""
import math

import numpy as np
import pytest

import oracle
from oracle import icp as OI


def rot(axis, ang):
    axis = np.asarray(axis, np.float64) / np.linalg.norm(axis)
    K = np.array([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]])
    return np.eye(3) + math.sin(ang) * K + (1 - math.cos(ang)) * K @ K


def test_kabsch_recovers_rigid_motion_and_rejects_reflections():
    rng = np.random.default_rng(0)
    p = rng.standard_normal((200, 3))
    R, t = rot([1, 2, 3], 0.7), np.array([0.3, -0.2, 1.0])
    Rg, tg = OI.kabsch(p, p @ R.T + t)
    assert np.allclose(Rg, R, atol=1e-12) and np.allclose(tg, t, atol=1e-12)
    q = p * np.array([1, 1, -1.0])                     # a mirrored copy: the best PROPER rotation, det = +1
    Rm, _ = OI.kabsch(p, q)
    assert abs(np.linalg.det(Rm) - 1) < 1e-12
    flat = p * np.array([1, 1, 0.0])                   # planar cloud: still a rotation
    Rf, tf = OI.kabsch(flat, flat @ R.T + t)
    assert np.allclose(flat @ Rf.T + tf, flat @ R.T + t, atol=1e-10)


def test_oracle_icp_converges_on_a_small_problem():
    from genpc_b200.synthetic import superquadric

    tgt = superquadric(3, 600)
    R, t = rot([0, 0, 1], math.radians(4)), np.array([0.01, -0.015, 0.02])
    src = ((tgt[::2] - t) @ R).astype(np.float32)      # R src + t == tgt[::2]
    T, fit, rmse, its = OI.icp(src, tgt, 0.075)
    assert fit > 0.99 and rmse < 2e-3 and its <= 31
    assert np.allclose(T[:3, :3], R, atol=5e-3) and np.allclose(T[:3, 3], t, atol=5e-3)


@pytest.mark.gpu
def test_icp_step_kernel_vs_oracle_step_by_step(cuda):
    """Every call of genpc_icp_step against oracle.icp.icp_step on the same (cur, dist, idx): statistics to float32
    rounding, the updated transform within 1e-6 (Horn vs Kabsch, both float64), convergence flags identical; candidates
    with no inliers / fewer than three / already converged stay untouched."""
    import torch

    from genpc_b200 import _lib
    from genpc_b200.loss_functions import chamfer_3DDist
    from genpc_b200.synthetic import superquadric

    tgt = superquadric(9, 2500)
    K, Ns = 6, 700
    rng = np.random.default_rng(5)
    src = []
    for k in range(K):
        R, t = rot(rng.standard_normal(3), 0.02 * (k + 1)), (rng.random(3) - 0.5) * 0.04
        src.append(((tgt[k::3][:Ns] - t) @ R * (1.0 + 0.03 * (k - 3))).astype(np.float32))
    src[-1] = src[-1] + 5.0                            # never any inlier
    src[-2][3:] += 7.0                                 # exactly three inliers at most
    source = torch.from_numpy(np.stack(src)).to(cuda)
    target = torch.from_numpy(tgt).to(cuda)[None].expand(K, -1, -1).contiguous()
    T = torch.eye(4, device=cuda).repeat(K, 1, 1).contiguous()
    state = torch.zeros(K, 4, device=cuda)
    To = [np.eye(4, dtype=np.float32) for _ in range(K)]
    so = [np.zeros(4, np.float32) for _ in range(K)]
    cd, L = chamfer_3DDist(), _lib.lib()
    for it in range(12):
        cur = (source @ T[:, :3, :3].transpose(1, 2) + T[:, None, :3, 3]).contiguous()
        d1, _, i1, _ = cd(cur, target)
        rc = L.genpc_icp_step(_lib.ptr(cur), _lib.ptr(target), _lib.ptr(d1), _lib.ptr(i1), _lib.ptr(T), _lib.ptr(state), K, Ns, K,
                              tgt.shape[0], 0.075 ** 2, 1e-6, 1e-6, int(it < 11), _lib.current_stream(cuda))
        assert rc == 0
        torch.cuda.synchronize()
        curn, dn, inn = cur.cpu().numpy(), d1.cpu().numpy(), i1.cpu().numpy()
        for k in range(K):
            # the oracle steps from the KERNEL's previous transform (cur was built from it), so errors do not compound
            To[k], so[k] = OI.icp_step(curn[k], tgt, dn[k], inn[k], To[k], so[k], 0.075 ** 2, 1e-6, 1e-6, it < 11)
            gs = state[k].cpu().numpy()
            assert np.allclose(gs[:2], so[k][:2], rtol=2e-7, atol=1e-9), (it, k, gs, so[k])
            assert gs[2] == so[k][2] and gs[3] == so[k][3], (it, k, gs, so[k])
            assert np.abs(T[k].cpu().numpy() - To[k]).max() <= 1e-6, (it, k)
            To[k] = T[k].cpu().numpy().copy()
    assert np.array_equal(T[-1].cpu().numpy(), np.eye(4, dtype=np.float32)) and float(state[-1, 0]) == 0.0


@pytest.mark.gpu
def test_batched_icp_vs_oracle_loop(cuda):
    """icp_point_to_point (batched, on-device convergence) against the oracle's loop run candidate by candidate with the
    C oracle's nearest neighbours: same fitness / rmse to 1e-5, transforms to 1e-4, same iteration of convergence."""
    import torch

    from genpc_b200.reg_xyz import icp_point_to_point
    from genpc_b200.synthetic import superquadric

    tgt = superquadric(4, 1800)
    K = 4
    rng = np.random.default_rng(2)
    src = np.stack([((tgt[k::2][:800] - (rng.random(3) - 0.5) * 0.03) @ rot(rng.standard_normal(3), 0.03 * (k + 1))).astype(np.float32)
                    for k in range(K)])
    Tg, fg, rg = icp_point_to_point(torch.from_numpy(src).to(cuda), torch.from_numpy(tgt).to(cuda)[None], 0.075)

    def nn(cur, target):
        d, i = oracle.nn_distance(cur[None], target[None])
        return d[0], i[0]

    for k in range(K):
        To, fo, ro, _ = OI.icp(src[k], tgt, 0.075, nn=nn)
        assert abs(float(fg[k]) - fo) <= 1e-5 and abs(float(rg[k]) - ro) <= 1e-5, (k, float(fg[k]), fo, float(rg[k]), ro)
        assert np.abs(Tg[k].cpu().numpy() - To).max() <= 1e-4, (k, np.abs(Tg[k].cpu().numpy() - To).max())
