"""Scale / ICP search (reg_xyz mirror): batched candidate scoring equals one-at-a-time Completionloss calls bit for
bit; the batched point-to-point ICP recovers a known rigid motion; the fusion tail behaves as documented."""
import math

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_batched_scoring_equals_individual_calls(cuda):
    import torch

    from genpc_b200.reg_xyz import chamfer_partial_l1_batched
    from genpc_b200.utils.loss_util import Completionloss

    g = torch.Generator().manual_seed(0)
    src = torch.rand(12, 900, 3, generator=g).to(cuda)
    tgt = torch.rand(1, 1300, 3, generator=g).to(cuda).expand(12, -1, -1)
    got = chamfer_partial_l1_batched(src, tgt, 0.5)
    cl = Completionloss("cd_l1")
    for k in range(12):
        one = cl.chamfer_partial_l1(src[k:k + 1], tgt[k:k + 1].contiguous()) + 0.5 * cl.chamfer_partial_l1(tgt[k:k + 1].contiguous(), src[k:k + 1])
        assert torch.allclose(got[k], one, rtol=1e-6)


def test_icp_recovers_rigid_motion_and_scale_search_picks_true_scale(cuda):
    import torch

    from genpc_b200.reg_xyz import icp_point_to_point, iterative_scale_search
    from genpc_b200.synthetic import superquadric

    tgt = torch.from_numpy(superquadric(3, 3000)).to(cuda)
    ang = math.radians(4.0)
    R = torch.tensor([[math.cos(ang), -math.sin(ang), 0], [math.sin(ang), math.cos(ang), 0], [0, 0, 1.0]], device=cuda)
    t = torch.tensor([0.01, -0.015, 0.02], device=cuda)
    src = (tgt[::2] - t) @ R                     # so that R src + t == tgt[::2]
    T, fit, rmse = icp_point_to_point(src[None], tgt[None], 0.075)
    assert float(fit[0]) > 0.99 and float(rmse[0]) < 2e-3
    assert torch.allclose(T[0, :3, :3], R, atol=5e-3) and torch.allclose(T[0, :3, 3], t, atol=5e-3)
    # anisotropic scale search: source = target squeezed by (1/1.1, 1, 1/0.9) -> best scales ~ (1.1, 1.0, 0.9)
    src2 = tgt[::2] / torch.tensor([1.1, 1.0, 0.9], device=cuda)
    S, loss, Tb = iterative_scale_search(src2, tgt, [(0.8, 1.2)] * 3, 5, None, 0.5)
    assert np.allclose(np.diag(S)[:3], [1.1, 1.0, 0.9], atol=1e-6) and loss < 0.01


def test_remove_close_points_and_reg_points(cuda):
    import torch

    from genpc_b200.reg_xyz import reg_points, remove_close_points
    from genpc_b200.synthetic import partial_view, superquadric

    comp = superquadric(5, 6000)
    part = partial_view(comp, 5, 2500)
    tc, tp = torch.from_numpy(comp).to(cuda), torch.from_numpy(part).to(cuda)
    keep = remove_close_points(tp, tc, 1e-4)
    d = torch.cdist(tc, tp).min(1).values ** 2
    assert torch.equal(keep, d >= 1e-4) or (keep != (d >= 1e-4)).float().mean() < 1e-3   # fp rounding at the threshold
    out = reg_points(tp, tc * 1.2, cd_inv_weight=0.5, diff_init=False, reg_fine_xyz=False, n_fused=3000)
    assert out["fused"].shape[0] <= 3000 and out["fused"].shape[1] == 3 and torch.isfinite(out["fused"]).all()
    assert 0.8 <= out["best_scale"] <= 1.5


@pytest.mark.gpu
@pytest.mark.parametrize("n,k", [(1, 3), (7, 20), (300, 1), (5000, 20), (20000, 20), (4099, 32), (2048, 7)])
def test_knn_mean_distance_gpu_bit_exact(cuda, n, k):
    """genpc_knn_mean_distance against the oracle, bit for bit (same rounding order, ascending fp32 sum), on uniform,
    lattice (many ties) and shape clouds; then remove_statistical_outlier against the oracle's mask."""
    import oracle
    import torch

    from genpc_b200.reg_xyz import knn_mean_distance, remove_statistical_outlier
    from util import lattice_cloud, rand_cloud, shape_cloud

    for kind, x in (("rand", rand_cloud(n, 1, n)[0]), ("lattice", lattice_cloud(n + 1, 1, n, side=9)[0]),
                    ("shape", shape_cloud(n + 2, 1, n)[0])):
        t = torch.from_numpy(x).to(cuda)
        for inc in (True, False):
            got = knn_mean_distance(t, k, inc).cpu().numpy()
            exp = oracle.knn_mean_distance(x, k, inc)
            assert np.array_equal(got.view(np.int32), exp.view(np.int32)), (kind, inc)
        if k == 20:
            keep = remove_statistical_outlier(t, 20, 2.5).cpu().numpy()
            assert np.array_equal(keep, oracle.statistical_outlier_mask(oracle.knn_mean_distance(x, 20, True), 2.5)), kind


def test_icp_step_kernel_matches_the_torch_formulation(cuda):
    """genpc_icp_step (Horn closed form, on-device convergence) against the torch formulation it replaced (batched
    float64 Kabsch/SVD): transforms within 1e-4, fitness / rmse within 1e-5, over candidates that converge at different
    iterations, one that never gets 3 inliers, and a shared ([1,Nt,3]) as well as a per-candidate target."""
    import os

    import torch

    from genpc_b200.reg_xyz import icp_point_to_point
    from genpc_b200.synthetic import superquadric

    tgt = torch.from_numpy(superquadric(9, 2500)).to(cuda)
    K = 7
    g = torch.Generator().manual_seed(5)
    src = []
    for k in range(K):
        ang = 0.02 * (k + 1)
        R = torch.tensor([[math.cos(ang), -math.sin(ang), 0], [math.sin(ang), math.cos(ang), 0], [0, 0, 1.0]])
        t = (torch.rand(3, generator=g) - 0.5) * 0.04
        s = (tgt.cpu()[k::3][:700] - t) @ R * (1.0 + 0.03 * (k - 3))
        src.append(s)
    src[-1] = src[-1] + 5.0                                    # far away: no inliers, must stay at the identity
    source = torch.stack(src).to(cuda)
    res = {}
    for mode in ("kernel", "torch"):
        for shared in (True, False):
            if mode == "torch":
                os.environ["GENPC_ICP_TORCH"] = "1"
            try:
                tg = tgt[None] if shared else tgt[None].expand(K, -1, -1).contiguous()
                res[mode, shared] = icp_point_to_point(source, tg, 0.075)
            finally:
                os.environ.pop("GENPC_ICP_TORCH", None)
    for shared in (True, False):
        Tk, fk, rk = res["kernel", shared]
        Tt, ft, rt = res["torch", shared]
        assert torch.allclose(Tk, Tt, atol=1e-4), (Tk - Tt).abs().max()
        assert torch.allclose(fk, ft, atol=1e-5) and torch.allclose(rk, rt, atol=1e-5)
        assert torch.equal(Tk[-1], torch.eye(4, device=cuda)) and float(fk[-1]) == 0.0
        assert float(fk[:-1].min()) > 0.9
    assert torch.equal(res["kernel", True][0], res["kernel", False][0])     # deterministic, same arithmetic
