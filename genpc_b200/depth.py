"""Functional API over the DepthPrompting geometry kernels (csrc/depth.cu): cameras, projection,
z-buffer render and unprojection.  All tensors are CUDA float32 unless stated; nothing leaves the device."""
import math

import numpy as np
import torch

from . import _lib


# ---- cameras (host side, tiny): restated from utils/camera_utils.py:86-147 + kaolin's look-at/perspective ----
def fibonacci_sphere(samples, radius):
    """utils/camera_utils.py:86-102."""
    pts = []
    phi = math.pi * (3.0 - math.sqrt(5.0))
    for i in range(samples):
        y = 1 - (i / float(samples - 1)) * 2
        radius_y = math.sqrt(1 - y * y)
        theta = phi * i
        pts.append((math.cos(theta) * radius_y * radius, y * radius, math.sin(theta) * radius_y * radius))
    return np.array(pts)


def calculate_up_vector(eye_position, target_position):
    """utils/camera_utils.py:104-113."""
    gaze = np.asarray(target_position, dtype=np.float64) - np.asarray(eye_position, dtype=np.float64)
    world_up = np.array([0.0, 1.0, 0.0])
    if np.allclose(np.cross(gaze, world_up), 0):
        return np.array([0.0, 0.0, 1.0])
    side = np.cross(gaze, world_up)
    up = np.cross(side, gaze)
    return up / np.linalg.norm(up)


def make_camera(eye, at, up, fov, width=1, height=1, near=1e-2, far=1e2):
    """16-float camera record with the semantics of kaolin Camera.from_args(eye, at, up, fov, width, height)
    (look-at view matrix looking down -z, vertical-fov pinhole, OpenGL projection)."""
    eye, at, up = (np.asarray(a, dtype=np.float64) for a in (eye, at, up))
    back = eye - at
    back /= np.linalg.norm(back)
    right = np.cross(up, back)
    right /= np.linalg.norm(right)
    upv = np.cross(back, right)
    R = np.stack([right, upv, back])
    t = -R @ eye
    fy = 1.0 / math.tan(fov / 2.0)
    fx = fy / (float(width) / float(height))
    A = (far + near) / (far - near)
    Bc = 2.0 * far * near / (far - near)
    return np.concatenate([R.reshape(9), t, [fx, fy, A, Bc]]).astype(np.float32)


def create_cameras(num_views=8, distance=1.6, fovy=49.1, res=512, device="cuda"):
    """utils/camera_utils.py:115-147 (fibonacci distribution): returns (cams [V,16] tensor, eye_positions [V,3])."""
    eyes = fibonacci_sphere(num_views, distance)
    fov = math.pi * fovy / 180
    cams = np.stack([make_camera(e, np.zeros(3), calculate_up_vector(e, np.zeros(3)), fov, res, res) for e in eyes])
    return torch.from_numpy(cams).to(device), eyes


# ---- kernels -----------------------------------------------------------------------------------------
def _ws(V, device):
    n = _lib.lib().genpc_depth_workspace_bytes(V)
    return torch.empty(max(n, 8), dtype=torch.uint8, device=device), n


def project_uv(cams, xyz, rescale=True, padding=0.15):
    """cams [V,16], xyz [N,3] -> (ndc [V,N,3], uv [V,N,2], bounds [V,4]).  DepthPrompting.py:239-271."""
    _lib.require_cuda(cams, xyz)
    cams, xyz = cams.contiguous().float(), xyz.contiguous().float()
    V, N = cams.shape[0], xyz.shape[0]
    dev = xyz.device
    ndc = torch.empty(V, N, 3, device=dev)
    uv = torch.empty(V, N, 2, device=dev)
    bounds = torch.empty(V, 4, device=dev)
    with torch.cuda.device(dev):
        ws, n = _ws(V, dev)
        rc = _lib.lib().genpc_project_uv(_lib.ptr(cams), _lib.ptr(xyz), V, N, int(bool(rescale)), float(padding),
                                         _lib.ptr(ndc), _lib.ptr(uv), _lib.ptr(bounds), _lib.ptr(ws), n,
                                         _lib.current_stream(dev))
    _lib.check(rc, "genpc_project_uv")
    return ndc, uv, bounds


def zbuffer_render(uv, ndc, res, point_size=1, valid=None, colors=None):
    """-> dict(zbuf [V,res,res] int64 view of the packed words, idx [V,res,res] i32, depth [V,res,res],
    color [V,3,res,res] or None, zminmax [V,2])."""
    _lib.require_cuda(uv, ndc)
    uv, ndc = uv.contiguous(), ndc.contiguous()
    V, N = uv.shape[0], uv.shape[1]
    dev = uv.device
    zbuf = torch.empty(V, res, res, dtype=torch.int64, device=dev)
    idx = torch.empty(V, res, res, dtype=torch.int32, device=dev)
    dep = torch.empty(V, res, res, device=dev)
    zmm = torch.empty(V, 2, device=dev)
    col = None
    if colors is not None:
        colors = colors.contiguous().float()
        col = torch.empty(V, 3, res, res, device=dev)
    if valid is not None:
        valid = valid.to(torch.uint8).contiguous()
    with torch.cuda.device(dev):
        ws, n = _ws(V, dev)
        rc = _lib.lib().genpc_zbuffer_render(_lib.ptr(uv), _lib.ptr(ndc), _lib.ptr(valid), _lib.ptr(colors), V, N,
                                             int(res), int(point_size), _lib.ptr(zbuf), _lib.ptr(idx),
                                             _lib.ptr(dep), _lib.ptr(col), _lib.ptr(zmm), _lib.ptr(ws), n,
                                             _lib.current_stream(dev))
    _lib.check(rc, "genpc_zbuffer_render")
    return dict(zbuf=zbuf, idx=idx, depth=dep, color=col, zminmax=zmm)


def unproject(cams, bounds, zbuf, ndc, rescale=True):
    """Every non-empty z-buffer pixel back to 3-D (pixel centre, owner's depth), raster order.
    -> (points [V,res*res,3] (first counts[v] rows valid), owner [V,res*res] i32, counts [V] i32)."""
    _lib.require_cuda(cams, bounds, zbuf, ndc)
    V, res = zbuf.shape[0], zbuf.shape[1]
    N = ndc.shape[1]
    dev = zbuf.device
    out = torch.zeros(V, res * res, 3, device=dev)
    own = torch.full((V, res * res), -1, dtype=torch.int32, device=dev)
    counts = torch.zeros(V, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        rc = _lib.lib().genpc_unproject(_lib.ptr(cams.contiguous()), _lib.ptr(bounds.contiguous()),
                                        int(bool(rescale)), _lib.ptr(zbuf.contiguous()), _lib.ptr(ndc.contiguous()),
                                        V, N, res, _lib.ptr(out), _lib.ptr(own), _lib.ptr(counts),
                                        _lib.current_stream(dev))
    _lib.check(rc, "genpc_unproject")
    return out, own, counts
