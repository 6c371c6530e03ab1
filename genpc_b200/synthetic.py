"""Seeded synthetic shapes standing in for the generators' outputs (InstantMesh / TRELLIS meshes are out of
scope, BASELINE.json north_star).  numpy only; used by bench.py, smoke() and the tests -- not a hot path."""
import numpy as np


def _superquadric_params(seed, exps=None):
    rng = np.random.default_rng(seed)
    e1, e2 = exps if exps is not None else rng.uniform(0.3, 1.6, size=2)
    ax = rng.uniform(0.4, 1.0, size=3)
    return rng, (e1, e2, ax)


def _superquadric_raw(params, rng, n):
    e1, e2, ax = params
    eta = rng.uniform(-np.pi / 2, np.pi / 2, n)
    om = rng.uniform(-np.pi, np.pi, n)

    def f(w, e):
        return np.sign(w) * np.abs(w) ** e

    x = ax[0] * f(np.cos(eta), e1) * f(np.cos(om), e2)
    y = ax[1] * f(np.cos(eta), e1) * f(np.sin(om), e2)
    z = ax[2] * f(np.sin(eta), e1)
    return np.stack([x, y, z], 1)


def superquadric(seed, n, exps=None):
    """n points on a superquadric surface, normalised like normalize_numpy(range=0.5)
    (reference utils/dataUtils.py:561-581): centred, longest bbox side = 1 -> coords in [-0.5, 0.5]."""
    rng, params = _superquadric_params(seed, exps)
    p = _superquadric_raw(params, rng, n)
    lo, hi = p.min(0), p.max(0)
    p = (p - (lo + hi) / 2) / (hi - lo).max()
    return p.astype(np.float32)


def superquadric_pair(seed, n, n_second):
    """Two INDEPENDENT samplings of the same surface in the same frame (the first one defines the normalisation): a
    complete cloud and the raw material of a partial scan that shares no sample with it."""
    rng, params = _superquadric_params(seed)
    p = _superquadric_raw(params, rng, n)
    q = _superquadric_raw(params, np.random.default_rng(seed + 15485863), n_second)
    lo, hi = p.min(0), p.max(0)
    c, s = (lo + hi) / 2, (hi - lo).max()
    return ((p - c) / s).astype(np.float32), ((q - c) / s).astype(np.float32)


def partial_view(points, seed, n_out):
    """Subset visible from a random direction (front half along the view axis), resampled to n_out."""
    rng = np.random.default_rng(seed + 7919)
    v = rng.standard_normal(3)
    v /= np.linalg.norm(v)
    depth = points @ v
    vis = np.nonzero(depth > np.quantile(depth, 0.45))[0]
    sel = rng.choice(vis, size=n_out, replace=len(vis) < n_out)
    return points[sel].astype(np.float32)


def rigid_perturb(points, seed, max_rot_deg=30.0, max_t=0.1, scale_range=(0.7, 1.3)):
    """Known ground-truth pose for the registration workload: p' = s R p + t."""
    rng = np.random.default_rng(seed + 104729)
    axis = rng.standard_normal(3)
    axis /= np.linalg.norm(axis)
    ang = np.deg2rad(rng.uniform(-max_rot_deg, max_rot_deg))
    K = np.array([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]])
    R = np.eye(3) + np.sin(ang) * K + (1 - np.cos(ang)) * K @ K
    t = rng.standard_normal(3)
    t *= rng.uniform(0, max_t) / np.linalg.norm(t)
    s = rng.uniform(*scale_range)
    return (s * points @ R.T + t).astype(np.float32), (R.astype(np.float32), t.astype(np.float32), np.float32(s))


def pcn_batch(seed0, B, n_partial=2048, n_complete=16384):
    """BASELINE config C2: B shapes, complete = n_complete surface samples, partial = n_partial points of a SEPARATE
    sampling of the same surface seen from a random direction.  The two clouds share no sample, so
    neither direction's distances are trivially zero (r01 drew the partial cloud from the complete one: dist1 == 0,
    which would flatter any filtering / pruning kernel)."""
    comp, part = [], []
    for b in range(B):
        c, raw = superquadric_pair(seed0 + b, n_complete, 3 * n_partial)
        comp.append(c)
        part.append(partial_view(raw, seed0 + b, n_partial))
    return np.stack(part), np.stack(comp)


def superquadric_mesh(seed, n_eta=48, n_om=96):
    """Triangle mesh of the same surface family (a stand-in for the generators' .glb output): a latitude / longitude
    grid over the parametric domain, two triangles per cell, normalised like `superquadric`.  -> (verts f32 [V,3],
    faces i32 [F,3], vertex colours f32 [V,3] = min-max normalised coordinates)."""
    _, (e1, e2, ax) = _superquadric_params(seed)
    eta = np.linspace(-np.pi / 2, np.pi / 2, n_eta)
    om = np.linspace(-np.pi, np.pi, n_om, endpoint=False)
    E, O = np.meshgrid(eta, om, indexing="ij")

    def f(w, e):
        return np.sign(w) * np.abs(w) ** e

    p = np.stack([ax[0] * f(np.cos(E), e1) * f(np.cos(O), e2), ax[1] * f(np.cos(E), e1) * f(np.sin(O), e2),
                  ax[2] * f(np.sin(E), e1)], -1).reshape(-1, 3)
    lo, hi = p.min(0), p.max(0)
    p = (p - (lo + hi) / 2) / (hi - lo).max()
    i, j = np.meshgrid(np.arange(n_eta - 1), np.arange(n_om), indexing="ij")
    a = (i * n_om + j).reshape(-1)
    b = (i * n_om + (j + 1) % n_om).reshape(-1)
    c = ((i + 1) * n_om + j).reshape(-1)
    d = ((i + 1) * n_om + (j + 1) % n_om).reshape(-1)
    faces = np.concatenate([np.stack([a, b, c], 1), np.stack([b, d, c], 1)]).astype(np.int32)
    col = (p - p.min(0)) / (p.max(0) - p.min(0) + 1e-8)
    return p.astype(np.float32), faces, col.astype(np.float32)


def lidar_scene_pair(n, seed=0):
    """BASELINE config C5: two n-point LiDAR-like clouds -- ground plane + ~100 object blobs, shuffled;
    the second = the first rigidly perturbed (1 cm-scale motion) + 2 cm jitter.  torch CPU tensors [n,3] float32."""
    import torch

    g = torch.Generator().manual_seed(seed)
    ground = torch.rand(n // 2, 3, generator=g) * torch.tensor([80.0, 80.0, 0.05]) - torch.tensor([40.0, 40.0, 0.0])
    centers = torch.rand(100, 3, generator=g) * torch.tensor([70.0, 70.0, 0.0]) - torch.tensor([35.0, 35.0, -1.0])
    objs = centers[torch.randint(0, 100, (n - n // 2,), generator=g)] + \
        torch.randn(n - n // 2, 3, generator=g) * torch.tensor([1.5, 0.8, 0.7])
    a = torch.cat([ground, objs])[torch.randperm(n, generator=g)].contiguous()
    ang = 0.01
    R = torch.tensor([[np.cos(ang), -np.sin(ang), 0], [np.sin(ang), np.cos(ang), 0], [0, 0, 1]], dtype=torch.float32)
    b = (a @ R.T + torch.tensor([0.05, -0.03, 0.01]) + torch.randn(n, 3, generator=g) * 0.02).contiguous()
    return a, b
