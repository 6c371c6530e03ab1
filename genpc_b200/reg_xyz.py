"""Scale / ICP search of the Geometric-Preserving-Fusion stage on the GPU (reference: reg_xyz.py).

The reference scores 11 isotropic + 1000 anisotropic scale candidates one at a time: deepcopy -> Open3D ICP on the
CPU -> H2D -> two Chamfer extension calls (4 NN scans) -> `cd < best_loss` D2H sync (reg_xyz.py:60-96, 146-173).
Here the candidates ARE the batch dimension: one batched point-to-point ICP (NN from the Chamfer kernel + one
genpc_icp_step launch per iteration: inliers, covariance, Horn's closed-form rotation, convergence) and one Chamfer call
score all of them, everything stays on the device.

Open3D is not vendored: `registration_icp` (point-to-point, default criteria: 30 iterations, relative fitness /
RMSE 1e-6), voxel down-sampling and `remove_statistical_outlier` are restated from their documented behaviour
(parity unpinned).  Function names and argument meaning follow the reference.
"""
import os

import numpy as np
import torch

from . import _lib
from .fps import furthest_point_sample
from .loss_functions import chamfer_3DDist
from .optim_registration.diff_obj_pose import object_pose_optimization_points
from .utils.dataUtils import normalize_numpy, voxel_down_sample  # noqa: F401

_cd = chamfer_3DDist()


def _apply(T, pts):
    """T [K,4,4], pts [K,N,3] -> transformed points."""
    return pts @ T[:, :3, :3].transpose(1, 2) + T[:, None, :3, 3]


def chamfer_partial_l1_batched(src, tgt, cd_inv_weight=0.0):
    """Per-candidate score of the reference's sweep (reg_xyz.py:81-84, 167-170):
    CDp-L1(src->tgt) + cd_inv_weight * CDp-L1(tgt->src), src [K,Ns,3], tgt [K,Nt,3] -> [K]."""
    d1, d2, _, _ = _cd(src.contiguous(), tgt.contiguous())
    return torch.sqrt(d1).mean(1) + cd_inv_weight * torch.sqrt(d2).mean(1)


def icp_point_to_point(source, target, max_correspondence_distance=0.05, init_transform=None, max_iteration=30,
                       relative_fitness=1e-6, relative_rmse=1e-6):
    """Batched Open3D-style registration_icp (TransformationEstimationPointToPoint, no scaling), reg_xyz.py:9-38.
    source [K,Ns,3], target [K,Nt,3] (or [1,Nt,3]) -> (T [K,4,4], fitness [K], inlier_rmse [K]).
    Per iteration: transform (one bmm), nearest neighbours of all candidates (one Chamfer launch pair) and ONE
    genpc_icp_step launch (inliers, fitness / rmse / convergence, covariance, Horn rotation, T update) -- no host
    synchronisation except a convergence poll every 4 iterations.  GENPC_ICP_TORCH=1 selects the torch formulation
    (batched float64 SVD, ~25 launches and a host sync per iteration) the kernel replaced; both agree to rounding."""
    if os.environ.get("GENPC_ICP_TORCH") == "1":
        return _icp_point_to_point_torch(source, target, max_correspondence_distance, init_transform, max_iteration,
                                         relative_fitness, relative_rmse)
    _lib.require_cuda(source, target)
    K, Ns, _ = source.shape
    dev = source.device
    source = source.contiguous().float()
    tgt = target.float()
    if tgt.shape[0] == 1 and K > 1:
        tgt = tgt.expand(K, -1, -1)
    tgt = tgt.contiguous()
    Nt = tgt.shape[1]
    T = torch.eye(4, device=dev).repeat(K, 1, 1) if init_transform is None else \
        torch.as_tensor(init_transform, dtype=torch.float32, device=dev).expand(K, 4, 4).clone()
    T = T.contiguous()
    state = torch.zeros(K, 4, device=dev)
    thr2 = float(max_correspondence_distance) ** 2
    L = _lib.lib()
    with torch.no_grad(), torch.cuda.device(dev):
        for it in range(max_iteration + 1):
            cur = _apply(T, source).contiguous()
            d1, _, i1, _ = _cd(cur, tgt)
            rc = L.genpc_icp_step(_lib.ptr(cur), _lib.ptr(tgt), _lib.ptr(d1), _lib.ptr(i1), _lib.ptr(T), _lib.ptr(state), K, Ns,
                                  K, Nt, thr2, float(relative_fitness), float(relative_rmse), int(it < max_iteration),
                                  _lib.current_stream(dev))
            _lib.check(rc, "genpc_icp_step")
            if it % 4 == 3 and bool(state[:, 2].all()):   # every candidate converged: the remaining iterations are no-ops
                break
    return T, state[:, 0].clone(), state[:, 1].clone()


def _icp_point_to_point_torch(source, target, max_correspondence_distance=0.05, init_transform=None, max_iteration=30,
                              relative_fitness=1e-6, relative_rmse=1e-6):
    """The torch formulation of icp_point_to_point (kept for A/B checks: tests/test_reg_xyz.py)."""
    K, Ns, _ = source.shape
    dev = source.device
    if target.shape[0] == 1 and K > 1:
        target = target.expand(K, -1, -1)
    target = target.contiguous()
    T = torch.eye(4, device=dev).repeat(K, 1, 1) if init_transform is None else \
        torch.as_tensor(init_transform, dtype=torch.float32, device=dev).expand(K, 4, 4).clone()
    thr2 = float(max_correspondence_distance) ** 2
    active = torch.ones(K, dtype=torch.bool, device=dev)
    fit_prev = torch.zeros(K, device=dev)
    rmse_prev = torch.zeros(K, device=dev)
    fitness = torch.zeros(K, device=dev)
    rmse = torch.zeros(K, device=dev)
    for it in range(max_iteration + 1):
        cur = _apply(T, source)
        d1, _, i1, _ = _cd(cur.contiguous(), target)
        inl = d1 < thr2
        cnt = inl.sum(1).clamp(min=1).float()
        fitness = inl.sum(1).float() / Ns
        rmse = torch.sqrt((d1 * inl).sum(1) / cnt)
        if it > 0:
            conv = ((fitness - fit_prev).abs() < relative_fitness) & ((rmse - rmse_prev).abs() < relative_rmse)
            active = active & ~conv
        if it == max_iteration or not bool(active.any()):
            break
        fit_prev, rmse_prev = fitness, rmse
        # Kabsch on the inlier correspondences (cur -> target[nn])
        nn = torch.gather(target, 1, i1.long()[..., None].expand(-1, -1, 3))
        w = inl.float()[..., None]
        mu_s = (cur * w).sum(1) / cnt[:, None]
        mu_t = (nn * w).sum(1) / cnt[:, None]
        H = ((cur - mu_s[:, None]) * w).transpose(1, 2) @ (nn - mu_t[:, None])
        U, _, Vh = torch.linalg.svd(H.double())
        V, Ut = Vh.transpose(1, 2), U.transpose(1, 2)
        D = torch.eye(3, dtype=torch.float64, device=dev).repeat(K, 1, 1)
        D[:, 2, 2] = torch.sign(torch.linalg.det(V @ Ut))
        R = (V @ D @ Ut).float()
        t = mu_t - (R @ mu_s[..., None])[..., 0]
        upd = torch.eye(4, device=dev).repeat(K, 1, 1)
        upd[:, :3, :3], upd[:, :3, 3] = R, t
        ok = (active & (inl.sum(1) >= 3))[:, None, None]
        T = torch.where(ok, upd @ T, T)
    return T, fitness, rmse


def icp_with_scaling_xyz(source, target, scales, max_correspondence_distance=0.05, init_transform=None):
    """reg_xyz.py:9-21 batched: scale the source per axis (scales [K,3]) then ICP.  Returns (scaled source, T, fitness, rmse)."""
    scales = torch.as_tensor(scales, dtype=torch.float32, device=source.device)
    src = source[None] * scales[:, None, :] if source.dim() == 2 else source * scales[:, None, :]
    T, f, r = icp_point_to_point(src, target if target.dim() == 3 else target[None], max_correspondence_distance,
                                 init_transform)
    return src, T, f, r


def icp_with_scaling(source, target, scale, max_correspondence_distance=0.05, init_transform=None):
    """reg_xyz.py:24-38 batched over `scale` [K]: ICP, then ICP again from result @ diag(scale)."""
    scale = torch.as_tensor(scale, dtype=torch.float32, device=source.device).reshape(-1)
    K = scale.shape[0]
    src = source[None].expand(K, -1, -1).contiguous()
    tgt = target[None].contiguous()
    T0, _, _ = icp_point_to_point(src, tgt, max_correspondence_distance, init_transform)
    S = torch.eye(4, device=source.device).repeat(K, 1, 1)
    S[:, 0, 0] = S[:, 1, 1] = S[:, 2, 2] = scale
    return icp_point_to_point(src, tgt, max_correspondence_distance, T0 @ S)


def iterative_scale_search(source_xyz, target_xyz, scale_ranges, scale_steps, init_transform=None, cd_inv_weight=0.0,
                           chunk=250):
    """reg_xyz.py:60-96: scale_steps^3 anisotropic candidates (z outer, x, y inner -- the reference's loop order, so
    `cd < best_loss` ties resolve identically: first candidate wins), each ICP-refined and scored by partial CD-L1.
    Returns (best_scales_transformation 4x4 np, best_loss float, best_transformation 4x4 np)."""
    dev = source_xyz.device
    xs = np.linspace(scale_ranges[0][0], scale_ranges[0][1], scale_steps)
    ys = np.linspace(scale_ranges[1][0], scale_ranges[1][1], scale_steps)
    zs = np.linspace(scale_ranges[2][0], scale_ranges[2][1], scale_steps)
    cand = np.array([[x, y, z] for z in zs for x in xs for y in ys], dtype=np.float32)
    losses, Ts = [], []
    for c0 in range(0, len(cand), chunk):
        sc = torch.from_numpy(cand[c0:c0 + chunk]).to(dev)
        src, T, _, _ = icp_with_scaling_xyz(source_xyz, target_xyz, sc, 0.075, init_transform)
        tgt = target_xyz[None].expand(src.shape[0], -1, -1)
        # the reference scores the SCALED (not ICP-transformed) source copy against the target (:79-84)
        losses.append(chamfer_partial_l1_batched(src, tgt, cd_inv_weight))
        Ts.append(T)
    losses, Ts = torch.cat(losses), torch.cat(Ts)
    k = int(torch.argmin(losses))  # argmin returns the first minimum == the reference's strict `<` update
    best = np.eye(4)
    best[0, 0], best[1, 1], best[2, 2] = cand[k]
    return best, float(losses[k]), Ts[k].double().cpu().numpy()


def remove_close_points(source_xyz, target_xyz, distance_threshold=0.0001):
    """reg_xyz.py:41-57: keep the target points whose nearest source point is at squared distance >= threshold
    (Open3D's KD-tree returns squared distances).  One NN launch instead of a Python loop of KD-tree queries."""
    d1, _, _, _ = _cd(target_xyz[None].contiguous(), source_xyz[None].contiguous())
    return d1[0] >= distance_threshold


def knn_mean_distance(xyz, k, include_self=True):
    """Mean distance from each point of xyz [N,3] (CUDA) to its k nearest points of the same cloud
    (genpc_knn_mean_distance; include_self counts the point itself, as a KD-tree query of a cloud point does)."""
    _lib.require_cuda(xyz)
    x = xyz.contiguous().float()
    out = torch.empty(x.shape[0], dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        rc = _lib.lib().genpc_knn_mean_distance(_lib.ptr(x), x.shape[0], int(k), int(bool(include_self)), _lib.ptr(out),
                                                _lib.current_stream(x.device))
    _lib.check(rc, "genpc_knn_mean_distance")
    return out


def remove_statistical_outlier(xyz, nb_neighbors=20, std_ratio=2.5):
    """Open3D remove_statistical_outlier as the reference uses it (reg_xyz.py:219, utils/dataUtils.py:652-666),
    restated on the k-NN kernel: per point the mean distance to its nb_neighbors nearest points -- the KD-tree query of
    a cloud point returns the point itself first, so it is one of them -- then keep 0 < mean < cloud_mean + std_ratio *
    std_dev (float64 statistics; means of exactly 0 -- duplicated points -- stay out of both sums, as in Open3D).  Returns the boolean keep mask.  (Open3D is not vendored: semantics from its
    documented behaviour, parity unpinned; the oracle defines them.)"""
    md = knn_mean_distance(xyz, nb_neighbors, include_self=True).double()
    nv = int((md >= 0).sum())
    if nv == 0:
        return torch.zeros_like(md, dtype=torch.bool)
    pos = md > 0            # Open3D: only positive means enter the sums, the divisors count every answered query
    mu = md[pos].sum() / nv
    sd = torch.sqrt(((md[pos] - mu) ** 2).sum() / (nv - 1)) if nv > 1 else torch.zeros((), dtype=torch.float64, device=md.device)
    return pos & (md < mu + std_ratio * sd)


def get_rotate_matrix(axis, angle):
    """utils/dataUtils.py:455-471 (degrees)."""
    a = angle * np.pi / 180
    c, s_ = np.cos(a), np.sin(a)
    if axis == "x":
        return np.array([[1, 0, 0], [0, c, -s_], [0, s_, c]])
    if axis == "y":
        return np.array([[c, 0, s_], [0, 1, 0], [-s_, 0, c]])
    if axis == "z":
        return np.array([[c, -s_, 0], [s_, c, 0], [0, 0, 1]])
    raise ValueError("axis should be x,y,z")


GENERATIVE_MODELS = ("instantmesh", "trellis", "sf3d")


def reg_points(partial_xyz, complete_xyz, cd_inv_weight=0.5, diff_init=True, reg_fine_xyz=False, dataset="redwood",
               n_fused=20000, lr=0.01, iters=200, partial_rgb=None, complete_rgb=None, generative_model="trellis",
               diff_complete_xyz=None):
    """In-memory form of reg() (reg_xyz.py:99-223): differentiable init -> generator-specific frame fix-up -> 11-scale
    coarse ICP sweep -> optional per-axis scale grid -> fuse (drop generated points that coincide with scan points, FPS to
    n_fused, outlier filter).  partial_xyz = the scan (`color_point.ply`), complete_xyz = the generated shape (163 840
    surface samples in the reference, :125); colours ride along through every selection (:211-219).  diff_complete_xyz:
    the cloud the differentiable init registers (the reference samples the mesh a second time, 120 000 points,
    diff_obj_pose.py:504); default = complete_xyz.  Returns a dict of device tensors.
    Deviations from the reference, all forced by unseeded / unvendored third-party calls: FPS
    starts at index 0 (fpsample: random start) and is skipped when the fused cloud already has <= n_fused points."""
    if generative_model not in GENERATIVE_MODELS:
        raise ValueError(f"generative_model {generative_model!r} not supported (reg_xyz.py:133-140 knows {GENERATIVE_MODELS})")
    dev = partial_xyz.device
    source, target = partial_xyz.float(), complete_xyz.float()
    src_rgb = None if partial_rgb is None else partial_rgb.float().to(dev)
    tgt_rgb = None if complete_rgb is None else complete_rgb.float().to(dev)
    diff_T = np.eye(4)
    if diff_init:
        dc = target if diff_complete_xyz is None else diff_complete_xyz.float()
        T = object_pose_optimization_points(voxel_down_sample(dc, 0.02), voxel_down_sample(source, 0.02), lr=lr,
                                            iters=iters, device=dev)
        diff_T = np.linalg.inv(T)                                   # reg_xyz.py:122
    dT = torch.as_tensor(diff_T, dtype=torch.float32, device=dev)
    source = source @ dT[:3, :3].T + dT[:3, 3]                      # partial into the generated shape's frame (:126)
    tn, _, _ = normalize_numpy(target.cpu().numpy(), range=0.5)     # :131
    if generative_model == "instantmesh":                           # :133-138
        keep0 = remove_statistical_outlier(source, 20, 1.5)         # remove_noise_from_point_cloud(source_pcd)
        source = source[keep0]
        src_rgb = None if src_rgb is None else src_rgb[keep0]
        tn = np.dot(np.dot(tn, get_rotate_matrix("x", 90).T), get_rotate_matrix("y", 90).T)
    target = torch.as_tensor(tn, dtype=torch.float32, device=dev)
    # coarse sweep (:146-173), all 11 scales at once
    scales = np.linspace(1.5, 0.8, 11)
    s_down, t_down = voxel_down_sample(source, 0.03), voxel_down_sample(target, 0.03)
    Ts, _, _ = icp_with_scaling(s_down, t_down, scales, 0.075)
    inv = torch.linalg.inv(Ts.double()).float()
    t_moved = _apply(inv, t_down[None].expand(len(scales), -1, -1))
    cd = chamfer_partial_l1_batched(s_down[None].expand(len(scales), -1, -1), t_moved, cd_inv_weight)
    k = int(torch.argmin(cd))
    coarse = Ts[k]
    out = {"best_scale": float(scales[k]), "coarse_loss": float(cd[k]), "coarse_transformation": coarse,
           "diff_transform": diff_T}
    tgt_full = target
    if reg_fine_xyz:
        source = source @ coarse[:3, :3].T + coarse[:3, 3]          # :176
        s_in = source if dataset in ("pcn", "kitti") else voxel_down_sample(source, 0.03)
        t_in = voxel_down_sample(target, 0.04 if dataset in ("pcn", "kitti") else 0.03)
        Sx, loss_xyz, Txyz = iterative_scale_search(s_in, t_in, [(0.8, 1.2)] * 3, 10, None, cd_inv_weight)
        for Mx in (np.linalg.inv(Sx), np.linalg.inv(Txyz)):         # :194-197
            Mt = torch.as_tensor(Mx, dtype=torch.float32, device=dev)
            tgt_full = tgt_full @ Mt[:3, :3].T + Mt[:3, 3]
        ci = torch.linalg.inv(coarse.double()).float()
        source = source @ ci[:3, :3].T + ci[:3, 3]                  # :199-200
        out.update(best_scales=np.diag(Sx)[:3].copy(), fine_loss=loss_xyz)
    for Mt in (torch.linalg.inv(coarse.double()).float(), torch.as_tensor(np.linalg.inv(diff_T), dtype=torch.float32, device=dev)):
        tgt_full = tgt_full @ Mt[:3, :3].T + Mt[:3, 3]              # :202-205
    di = torch.as_tensor(np.linalg.inv(diff_T), dtype=torch.float32, device=dev)
    source = source @ di[:3, :3].T + di[:3, 3]                      # :206
    keep = remove_close_points(source, tgt_full, 0.0001)            # :210
    fused = torch.cat([source, tgt_full[keep]])                     # :211 fused = source + filtered target
    with_rgb = src_rgb is not None and tgt_rgb is not None
    fused_rgb = torch.cat([src_rgb, tgt_rgb[keep]]) if with_rgb else None
    if fused.shape[0] > n_fused:
        idx = furthest_point_sample(fused[None].contiguous(), n_fused, 0)[0].long()   # :215-217
        fused = fused[idx]
        fused_rgb = fused_rgb[idx] if with_rgb else None
    ok = remove_statistical_outlier(fused, std_ratio=2.5)                             # :219
    fused = fused[ok]
    fused_rgb = fused_rgb[ok] if with_rgb else None
    out.update(fused=fused, fused_rgb=fused_rgb, source=source, target=tgt_full)
    return out
