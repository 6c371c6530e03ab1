"""Geometry of the reference's DepthPrompting stage (DepthPrompting.py:87-98, 239-391) on the GPU.

Same method names and argument meaning as the reference class; the generators (depth inpainting, depth->image
diffusion, :20-85) are out of scope (BASELINE.json north_star).  Differences, all deliberate:
  * cameras are [V,16] float records (depth.make_camera) instead of kaolin Camera objects (kaolin is not vendored);
  * visibility is decided by the z-buffer (a point is visible in a view iff it owns a pixel) instead of Open3D's
    CPU hidden_point_removal (:273-290, third-party, adjacent -- SURVEY.md R8);
  * painting has a depth test (nearest wins, lowest index on ties) where the reference's index_put lets an
    arbitrary point win (:305,:336).
"""
import math
from types import SimpleNamespace

import numpy as np
import torch

from . import depth as D
from .fps import furthest_point_sample

_DEFAULTS = dict(device="cuda", distance=1.6, fovy=49.1, point_size=1, mask_pixel_rate=3, downsample_num=10000,
                 cam_res=256, view_num=8, res=256, rescale=True, padding=0.15, dataset="redwood")


class DepthPrompting:
    def __init__(self, cfg=None):
        c = dict(_DEFAULTS)
        if cfg is not None:
            c.update(cfg if isinstance(cfg, dict) else {k: getattr(cfg, k) for k in _DEFAULTS if hasattr(cfg, k)})
        self.cfg = SimpleNamespace(**c)
        self.device = torch.device(self.cfg.device)
        # create_cameras (utils/camera_utils.py:115-147), fibonacci distribution
        self.cameras, self.viewpoints = D.create_cameras(self.cfg.view_num, self.cfg.distance, self.cfg.fovy,
                                                         self.cfg.cam_res, self.device)
        self._bounds = None

    # DepthPrompting.py:239-271
    def getUvs(self, cams, points, rescale=True, padding=0.15):
        ndc, uv, bounds = D.project_uv(cams, points, rescale, padding)
        self._bounds = bounds
        return uv, ndc[:, :, 2], ndc

    # DepthPrompting.py:273-290 -- z-buffer visibility instead of hidden_point_removal
    def getVisiblePoints(self, points, cams=None, res=None):
        cams = self.cameras if cams is None else cams
        res = res or self.cfg.res
        ndc, uv, _ = D.project_uv(cams, points, self.cfg.rescale, self.cfg.padding)
        r = D.zbuffer_render(uv, ndc, res, 1)
        V, N = uv.shape[0], uv.shape[1]
        # a point is visible in a view iff it owns at least one pixel of that view's z-buffer
        idx = r["idx"].reshape(V, -1)
        flat = (idx.to(torch.int64) + torch.arange(V, device=idx.device, dtype=torch.int64)[:, None] * (N + 1))
        flat = flat[idx >= 0]
        vis = torch.zeros(V * (N + 1), dtype=torch.bool, device=points.device)
        vis[flat] = True
        return vis.view(V, N + 1)[:, :N]

    # DepthPrompting.py:87-98
    def viewpoint_select(self, xyz):
        k = min(self.cfg.downsample_num, xyz.shape[0])
        idx = furthest_point_sample(xyz[None].contiguous(), k, 0)[0].long()
        visible = self.getVisiblePoints(xyz[idx])
        return torch.argmax(visible.sum(dim=1))

    # DepthPrompting.py:341-391 (+ paintPixels :292-339); takes uv/ndc of ONE view instead of pre-clipped pixels
    def getRawDepth(self, point_uv, point_ndc, colors=None, res=None, point_size=None, mask_pixel_rate=None,
                    valid=None):
        res = res or self.cfg.res
        ps = point_size or self.cfg.point_size
        mpr = mask_pixel_rate or self.cfg.mask_pixel_rate
        uv, ndc = point_uv[None], point_ndc[None]
        v = None if valid is None else valid[None]
        r = D.zbuffer_render(uv, ndc, res, ps, v, colors)
        big = D.zbuffer_render(uv, ndc, res, ps * mpr, v, None)
        sparse_depth = r["depth"].expand(3, -1, -1).contiguous()
        sparse_img = r["color"][0] if colors is not None else (r["idx"] >= 0).float().expand(3, -1, -1).contiguous()
        all_front_mask = (big["idx"] >= 0).float().expand(3, -1, -1)
        all_back_mask = 1 - all_front_mask
        front_mask = (r["idx"] >= 0).float().expand(3, -1, -1) if colors is None else (sparse_img != 0).float()
        back_mask = 1 - front_mask
        hole_mask1 = ((all_back_mask * 255).int() ^ (back_mask * 255).int()).float() / 255
        hole_mask2 = ((all_front_mask * 255).int() ^ (back_mask * 255).int()).float() / 255
        self._last_render = r
        return sparse_img, sparse_depth, hole_mask1, hole_mask2

    # the "depth -> point" half of the stage (no reference counterpart)
    def unproject(self, cam, bounds, zbuf, ndc):
        pts, own, counts = D.unproject(cam, bounds, zbuf, ndc, self.cfg.rescale)
        return pts, own, counts

    # DepthPrompting.py:100-237 geometry only: best view, uv/depth of that view, sparse depth + masks
    def getDepth(self, xyz, rgb=None):
        with torch.no_grad():
            # the reference projects ALL points through ALL view_num cameras first (877 MB at 1024 views x 71 372
            # points, :102) and then uses one view; the result is identical when the view is chosen first
            best = int(self.viewpoint_select(xyz))
            cam = self.cameras[best:best + 1]
            point_uvs, point_depths, ndc = self.getUvs(cam, xyz, self.cfg.rescale, self.cfg.padding)
            vis = self.getVisiblePoints(xyz, cam)[0]
            out = self.getRawDepth(point_uvs[0], ndc[0], rgb, valid=vis)
            self.point_uv = point_uvs[0]
            self.view = self.viewpoints[best]
            self.cam = self.cameras[best]
            return (best,) + tuple(out)
