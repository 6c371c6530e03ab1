"""Stage-2 façade of the reference (ScaleAdapter.py:15-86) for the geometric hot path.

Same class and method names.  `scaleReg` / `reg` run entirely on the GPU (genpc_b200.reg_xyz); the generator-facing
methods (`remove_bg`, `img2shape`: background removal and image-to-3D, ScaleAdapter.py:19-44,70-72) are out of
scope (BASELINE.json north_star) and raise.  File hand-off mirrors the reference's workspace layout: the generated shape is
`<flag>_<generative_model>.glb` (reg_xyz.py:105-107), parsed and surface-sampled here (utils/glb.py, csrc/mesh.cu).
"""
import os
from types import SimpleNamespace

import numpy as np
import torch

from .reg_xyz import reg_points
from .utils.dataUtils import load_xyz, write_ply_xyz
from .utils.glb import read_glb, sample_mesh


def reg(cfg, flag, cd_inv_weight=0.5, diff_init=True, reg_fine_xyz=False, seed=0):
    """reg_xyz.reg (reg_xyz.py:99-223) on the reference's workspace layout:
        <output_path>/<flag>/color_point.ply                      the coloured partial scan
        <output_path>/<flag>/<flag>_<generative_model>.glb        the generated mesh
     -> <output_path>/<flag>/<flag>_fused.ply                     (:220)
    The mesh is sampled on the GPU (utils/glb.glb2point: 163 840 points for the fusion :125, 120 000 for the
    differentiable init diff_obj_pose.py:504; seeded -- trimesh's sampler is not).  A `<flag>_<generative_model>.ply`
    point cloud is accepted in place of the .glb (r01 layout)."""
    path = cfg.output_path
    scan = f"{path}/{flag}/color_point.ply"
    glb = f"{path}/{flag}/{flag}_{cfg.generative_model}.glb"
    ply = f"{path}/{flag}/{flag}_{cfg.generative_model}.ply"
    if not os.path.exists(scan):
        print(f"Path {scan} does not exist.")
        raise FileNotFoundError(f"Path {scan} does not exist.")
    if not os.path.exists(glb) and not os.path.exists(ply):
        print(f"Path {glb} does not exist.")
        raise FileNotFoundError(f"Path {glb} does not exist.")
    dev = torch.device(cfg.device)
    partial, partial_rgb = load_xyz(scan)
    diff_complete = None
    if os.path.exists(glb):
        verts, faces, vcol = read_glb(glb)
        tv, tf = torch.from_numpy(verts).to(dev), torch.from_numpy(faces).to(dev)
        tc = None if vcol is None else torch.from_numpy(vcol).to(dev)
        complete, complete_rgb = sample_mesh(tv, tf, 163840, seed, tc)
        if diff_init:
            diff_complete, _ = sample_mesh(tv, tf, 120000, seed + 1, tc)
    else:
        c, crgb = load_xyz(ply)
        complete, complete_rgb = torch.from_numpy(c).to(dev), torch.from_numpy(crgb).to(dev)
    out = reg_points(torch.from_numpy(partial).to(dev), complete, cd_inv_weight, diff_init, reg_fine_xyz,
                     getattr(cfg, "dataset", "redwood"), partial_rgb=torch.from_numpy(partial_rgb).to(dev),
                     complete_rgb=complete_rgb, generative_model=cfg.generative_model, diff_complete_xyz=diff_complete)
    write_ply_xyz(f"{path}/{flag}/{flag}_fused.ply", out["fused"].cpu().numpy(), out["fused_rgb"].cpu().numpy())
    return out


class ScaleAdapter:
    def __init__(self, cfg):
        self.cfg = cfg if not isinstance(cfg, dict) else SimpleNamespace(**cfg)
        self.device = self.cfg.device

    def remove_bg(self, flag, img_resource):
        raise NotImplementedError("background removal (rembg / RMBG) is a pretrained generator: out of scope")

    def img2shape(self, flag):
        raise NotImplementedError("image-to-3D generation (InstantMesh / TRELLIS / SF3D) is out of scope")

    def colorPoint(self, flag, xyz, gt, rgb, img_resource, img=None, point_uv=None):
        """ScaleAdapter.py:46-68: gather image colours back onto the points through the saved uv (one device gather
        instead of a Python loop over points); `img` [3,H,W] float tensor, `point_uv` [N,2]."""
        out = f"{self.cfg.output_path}/{flag}/color_point.ply"
        os.makedirs(os.path.dirname(out), exist_ok=True)
        if img_resource == "obj":
            write_ply_xyz(out, xyz.detach().cpu().numpy(), rgb.detach().cpu().numpy())
            return
        if point_uv is None:
            point_uv = torch.as_tensor(np.load(f"{self.cfg.output_path}/{flag}/point_uv.npy"))
        point_uv = point_uv.to(xyz.device)
        img = torch.flip(img.to(xyz.device), dims=[1])                      # Image.FLIP_TOP_BOTTOM (:58)
        res = img.shape[1]
        px = (point_uv * res).long()                                        # (:61-64, hard-coded 1024 in the reference)
        row, col = px[:, 1].clip(0, res - 1), px[:, 0].clip(0, res - 1)
        colors = img[:, row, col].T
        write_ply_xyz(out, xyz.detach().cpu().numpy(), colors.cpu().numpy())
        return colors

    def scaleReg(self, flag):
        return reg(self.cfg, flag, cd_inv_weight=0.5, diff_init=True, reg_fine_xyz=True)   # ScaleAdapter.py:74-75

    def scaleAdapter(self, xyz, flag, rgb=None):
        """ScaleAdapter.py:78-86 minus the generators: colour the scan; the caller provides the generated shape."""
        self.colorPoint(flag, xyz, xyz, rgb if rgb is not None else torch.ones_like(xyz), img_resource="obj")
