"""DepthPrompting geometry: oracle invariants on CPU, CUDA kernels bit-exact against the oracle on the GPU."""
import math

import numpy as np
import pytest

import oracle
from util import rand_cloud, shape_cloud


def cams_np(V, res=256):
    from genpc_b200 import depth as D

    eyes = D.fibonacci_sphere(V, 1.6)
    fov = math.pi * 49.1 / 180
    return np.stack([D.make_camera(e, np.zeros(3), D.calculate_up_vector(e, np.zeros(3)), fov, res, res) for e in eyes])


def test_camera_record_matches_lookat_semantics():
    c = cams_np(4)
    R = c[:, :9].reshape(-1, 3, 3)
    assert np.allclose(R @ R.transpose(0, 2, 1), np.eye(3), atol=1e-6)
    ndc, uv, b = oracle.project_uv(c, np.zeros((1, 3), np.float32), rescale=False)
    assert np.allclose(ndc[:, 0, :2], 0, atol=1e-6)          # the look-at target projects to the image centre
    assert np.allclose(uv[:, 0], 0.5, atol=1e-6)


def test_oracle_uv_range_and_zbuffer_depth_test():
    pts = shape_cloud(0, 1, 5000)[0]
    c = cams_np(3)
    ndc, uv, b = oracle.project_uv(c, pts, True, 0.15)
    assert uv.min() >= 0.15 - 1e-5 and uv.max() <= 0.85 + 1e-5   # DepthPrompting.py:259-261 comment
    zb = oracle.zbuffer(uv, ndc, 128, 1)
    idx, dep = oracle.zbuffer_resolve(zb, ndc, np.stack([ndc[..., 2].min(1), ndc[..., 2].max(1)], 1))
    # every painted pixel holds the nearest of the points that map to it
    for v in range(3):
        col = np.clip((uv[v, :, 0] * np.float32(128)).astype(np.int64), 0, 127)
        row = np.clip((uv[v, :, 1] * np.float32(128)).astype(np.int64), 0, 127)
        for p in range(0, 5000, 97):
            owner = idx[v, 127 - row[p], col[p]]
            assert owner >= 0 and ndc[v, owner, 2] <= ndc[v, p, 2]
    assert dep.max() <= 0.9 + 1e-6 and dep[idx >= 0].min() >= 0.1 - 1e-6


def test_oracle_render_unproject_render_is_idempotent():
    pts = shape_cloud(1, 1, 20000)[0]
    c = cams_np(2)
    ndc, uv, b = oracle.project_uv(c, pts, True, 0.15)
    zb = oracle.zbuffer(uv, ndc, 64, 1)
    out, own, counts = oracle.unproject(c, b, zb, ndc, True)
    for v in range(2):
        n = counts[v]
        assert n == (zb[v] != np.uint64(0xFFFFFFFFFFFFFFFF)).sum()
        ndc2, _, _ = oracle.project_uv(c[v:v + 1], out[v, :n], rescale=False)
        # re-project with the ORIGINAL rescale bounds: same pixels are painted
        uv2 = ((ndc2[0, :, :2] - b[v, :2]) / b[v, 2]) * b[v, 3] + 0.5
        col = (uv2[:, 0] * 64).astype(np.int64)
        row = (uv2[:, 1] * 64).astype(np.int64)
        painted = np.zeros((64, 64), bool)
        painted[63 - row, col] = True
        assert np.array_equal(painted, zb[v] != np.uint64(0xFFFFFFFFFFFFFFFF))


@pytest.mark.gpu
@pytest.mark.parametrize("V,N,res,ps,rescale", [(8, 16384, 512, 1, True), (8, 16384, 512, 2, True), (3, 71372, 256, 3, True),
                                                (2, 1000, 64, 1, False), (1, 1, 16, 2, True)])
def test_depth_gpu_bit_exact(cuda, V, N, res, ps, rescale):
    import torch

    from genpc_b200 import depth as D

    pts = shape_cloud(N, 1, N)[0] if N > 1 else np.array([[0.1, 0.2, 0.05]], np.float32)
    c = cams_np(max(V, 2), res)[:V]
    tc, tp = torch.from_numpy(c).to(cuda), torch.from_numpy(pts).to(cuda)
    ndc, uv, bounds = D.project_uv(tc, tp, rescale, 0.15)
    endc, euv, eb = oracle.project_uv(c, pts, rescale, 0.15)
    assert np.array_equal(ndc.cpu().numpy().view(np.int32), endc.view(np.int32))
    if rescale:
        assert np.array_equal(bounds.cpu().numpy().view(np.int32), eb.view(np.int32))
    if N > 1 or not rescale:
        assert np.array_equal(uv.cpu().numpy().view(np.int32), euv.view(np.int32))
    rng = np.random.default_rng(0)
    valid = (rng.random((V, N)) < 0.8).astype(np.uint8)
    colors = rng.random((N, 3), dtype=np.float32)
    r = D.zbuffer_render(uv, ndc, res, ps, torch.from_numpy(valid).to(cuda), torch.from_numpy(colors).to(cuda))
    ezb = oracle.zbuffer(euv, endc, res, ps, valid)
    assert np.array_equal(r["zbuf"].cpu().numpy().view(np.uint64), ezb)
    z = np.where(valid.astype(bool), endc[..., 2], np.nan)
    zmm = np.stack([np.nanmin(z, 1), np.nanmax(z, 1)], 1).astype(np.float32) if N > 1 else None
    if zmm is not None and np.isfinite(zmm).all():
        assert np.array_equal(r["zminmax"].cpu().numpy(), zmm)
        eidx, edep = oracle.zbuffer_resolve(ezb, endc, zmm)
        assert np.array_equal(r["idx"].cpu().numpy(), eidx)
        assert np.array_equal(r["depth"].cpu().numpy().view(np.int32), edep.view(np.int32))
        ci = r["color"].cpu().numpy()
        assert np.array_equal(ci[:, :, eidx[0] >= 0][0].T, colors[eidx[0][eidx[0] >= 0]])
    if rescale:
        out, own, counts = D.unproject(tc, bounds, r["zbuf"], ndc, rescale)
        eout, eown, ecounts = oracle.unproject(c, eb, ezb, endc, rescale)
        assert np.array_equal(counts.cpu().numpy(), ecounts)
        for v in range(V):
            n = ecounts[v]
            assert np.array_equal(own[v, :n].cpu().numpy(), eown[v, :n])
            assert np.array_equal(out[v, :n].cpu().numpy().view(np.int32), eout[v, :n].view(np.int32))


@pytest.mark.gpu
def test_depthprompting_class_runs_and_is_deterministic(cuda):
    import torch

    from genpc_b200.DepthPrompting import DepthPrompting

    pts = torch.from_numpy(shape_cloud(5, 1, 30000)[0]).to(cuda)
    dp = DepthPrompting(dict(view_num=16, res=128, cam_res=128, downsample_num=4000))
    a = dp.getDepth(pts, torch.rand(30000, 3, device=cuda, generator=torch.Generator(device="cuda").manual_seed(0)))
    b = dp.getDepth(pts, torch.rand(30000, 3, device=cuda, generator=torch.Generator(device="cuda").manual_seed(0)))
    assert a[0] == b[0]
    for x, y in zip(a[1:], b[1:]):
        assert torch.equal(x, y)
    assert a[2].shape == (3, 128, 128) and 0.1 <= float(a[2][a[2] > 0].min()) and float(a[2].max()) <= 0.9 + 1e-6


def _oracle_visible(cams, pts, res, rescale=True, padding=0.15):
    """Z-buffer visibility restated on the oracle (a point is visible in a view iff it owns a pixel)."""
    ndc, uv, _ = oracle.project_uv(cams, pts, rescale, padding)
    zb = oracle.zbuffer(uv, ndc, res, 1)
    V, N = uv.shape[0], uv.shape[1]
    vis = np.zeros((V, N), bool)
    for v in range(V):
        w = zb[v][zb[v] != np.uint64(0xFFFFFFFFFFFFFFFF)]
        vis[v, (w & np.uint64(0xFFFFFFFF)).astype(np.int64)] = True
    return vis


@pytest.mark.gpu
@pytest.mark.parametrize("V,N,res", [(4, 3000, 64), (16, 20000, 128), (64, 10000, 256)])
def test_get_visible_points_vs_oracle(cuda, V, N, res):
    """DepthPrompting.getVisiblePoints (reference :273-290 uses Open3D hidden_point_removal; here z-buffer ownership)
    against the oracle's z-buffer: the boolean [V,N] mask bit for bit, and the properties the stage relies on -- a visible
    point is the nearest of its pixel, points on the far side of a closed surface are mostly hidden."""
    import torch

    from genpc_b200.DepthPrompting import DepthPrompting

    pts = shape_cloud(7, 1, N)[0]
    dp = DepthPrompting(dict(view_num=V, res=res, cam_res=res))
    vis = dp.getVisiblePoints(torch.from_numpy(pts).to(cuda)).cpu().numpy()
    exp = _oracle_visible(dp.cameras.cpu().numpy(), pts, res)
    assert vis.shape == (V, N) and np.array_equal(vis, exp)
    # at most one owner per pixel; a closed surface seen from outside hides its far side (back-facing points are
    # visible through sampling gaps only)
    assert (vis.sum(1) <= res * res).all() and (vis.sum(1) > 0).all()
    eyes = np.asarray(dp.viewpoints, np.float32)
    facing = np.einsum("vk,nk->vn", eyes / np.linalg.norm(eyes, axis=1, keepdims=True), pts / np.linalg.norm(pts, axis=1, keepdims=True))
    if N >= 10000 and res <= 128:
        assert (vis & (facing > 0.3)).sum() > 3 * (vis & (facing < -0.3)).sum()


@pytest.mark.gpu
def test_viewpoint_select_vs_oracle(cuda):
    """viewpoint_select (:87-98: FPS down-sample, visibility per view, argmax of the visible count) against the oracle
    pipeline (oracle.fps -> oracle z-buffer -> first maximum): same view, same counts."""
    import torch

    from genpc_b200.DepthPrompting import DepthPrompting

    rng = np.random.default_rng(3)
    pts = shape_cloud(11, 1, 25000)[0]
    pts = pts[pts @ np.array([0.3, 0.8, 0.5], np.float32) > -0.05]          # a partial scan: one side missing
    rng.shuffle(pts)
    dp = DepthPrompting(dict(view_num=32, res=128, cam_res=128, downsample_num=5000))
    t = torch.from_numpy(pts).to(cuda)
    best = int(dp.viewpoint_select(t))
    sel = np.asarray(oracle.fps(pts[None], 5000, 0))[0].astype(np.int64)
    counts = _oracle_visible(dp.cameras.cpu().numpy(), pts[sel], 128).sum(1)
    assert best == int(np.argmax(counts))
    got = dp.getVisiblePoints(t[torch.from_numpy(sel).to(cuda)]).sum(1).cpu().numpy()
    assert np.array_equal(got, counts)
