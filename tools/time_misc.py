"""Secondary workloads of BASELINE.json (C3 registration, C4 FPS + z-buffer) timed with CUDA events."""
import json
import math
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from genpc_b200 import depth as D  # noqa: E402
from genpc_b200.fps import furthest_point_sample  # noqa: E402
from genpc_b200.optim_registration.diff_obj_pose import RegistrationBatch  # noqa: E402
from genpc_b200.synthetic import partial_view, rigid_perturb, superquadric  # noqa: E402

dev = torch.device("cuda:0")
out = {}


def ev_time(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts), sum(ts) / len(ts)


# ---- C1: B=1, 71372 x 16384 Chamfer forward (the reference's own CPU-runnable case, here on the GPU) ----
from genpc_b200.loss_functions import chamfer_3DDist  # noqa: E402
a1 = torch.from_numpy(superquadric(1, 71372)[None]).to(dev)
b1 = torch.from_numpy(superquadric(2, 16384)[None]).to(dev)
mn, av = ev_time(lambda: chamfer_3DDist()(a1, b1))
out["c1_chamfer_fwd_71372x16384_ms"] = mn
out["c1_pairs_per_s"] = 2.0 * 71372 * 16384 / (mn * 1e-3)

# ---- C4: FPS 16384 -> 2048 (B=1 and B=32) ----
for B in (1, 32, 148):
    x = torch.rand(B, 16384, 3, device=dev)
    mn, av = ev_time(lambda: furthest_point_sample(x, 2048, 0))
    out[f"fps_16384_to_2048_B{B}_ms"] = mn
x = torch.rand(1, 4096, 3, device=dev)
out["fps_4096_to_1024_B1_ms"] = ev_time(lambda: furthest_point_sample(x, 1024, 0))[0]
x = torch.rand(1, 71372, 3, device=dev)
out["fps_71372_to_10000_B1_ms"] = ev_time(lambda: furthest_point_sample(x, 10000, 0), reps=2, warm=1)[0]

# ---- C4: 8 views, 512^2, project + z-buffer + unproject on 71372 points ----
pts = torch.from_numpy(superquadric(0, 71372)).to(dev)
cams, _ = D.create_cameras(8, 1.6, 49.1, 512, dev)
def c4():
    ndc, uv, b = D.project_uv(cams, pts, True, 0.15)
    r = D.zbuffer_render(uv, ndc, 512, 1)
    D.unproject(cams, b, r["zbuf"], ndc, True)
out["c4_project_zbuffer_unproject_8x512_71372pts_ms"] = ev_time(c4)[0]
cams1k, _ = D.create_cameras(1024, 1.6, 49.1, 256, dev)
p10k = pts[:10000].contiguous()
def views1k():
    ndc, uv, b = D.project_uv(cams1k, p10k, True, 0.15)
    D.zbuffer_render(uv, ndc, 256, 1)
out["visibility_1024views_256_10000pts_ms"] = ev_time(views1k)[0]

# ---- the reference's default stage-1 geometry: 1024 views, res 256, 71 372 points (DepthPrompting.getDepth) ----
from genpc_b200.DepthPrompting import DepthPrompting  # noqa: E402
dp = DepthPrompting(dict(view_num=1024, res=256, cam_res=256, downsample_num=10000))
rgb = torch.rand(71372, 3, device=dev)
dp.getDepth(pts, rgb); torch.cuda.synchronize()
ts = []
for _ in range(3):   # wall clock (the call synchronises on the selected view): best of three
    t0 = time.perf_counter(); dp.getDepth(pts, rgb); torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
out["depthprompting_getDepth_1024views_res256_71372pts_ms"] = min(ts)

# ---- C3: registration, 64 scans x 16384 pts, 1 start each (scan-iters/s) ----
S = int(os.environ.get("REG_SCANS", 64))
comp = np.stack([superquadric(s, 16384) for s in range(S)])
part = np.stack([rigid_perturb(partial_view(comp[s], s, 16384), s)[0] for s in range(S)])
tc, tp = torch.from_numpy(comp).to(dev), torch.from_numpy(part).to(dev)
iters = int(os.environ.get("REG_ITERS", 20))
rb = RegistrationBatch(tc, tp, n_starts=1, lr=0.01, max_iters=iters + 8)
rb.run(3); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); rb.run(iters); e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
out["c3_registration"] = {"scans": S, "pts": 16384, "iters_timed": iters, "ms_per_iter_all_scans": ms / iters,
                          "scan_iters_per_s": S * iters / (ms * 1e-3),
                          "pairs_per_s": S * iters * 2.0 * 16384 * 16384 / (ms * 1e-3),
                          "loss_first_last": [float(rb.losses()[0, 0]), float(rb.losses()[0, rb.t - 1])]}
# small-cloud regime of the real pipeline (about 1-3 K points after voxel down-sampling), 4 starts
comp_s, part_s = tc[:1, :2500].contiguous(), tp[:1, :1000].contiguous()
rb2 = RegistrationBatch(comp_s, part_s, n_starts=4, lr=0.01, max_iters=300)
rb2.run(10); torch.cuda.synchronize()
e0.record(); rb2.run(201); e1.record(); torch.cuda.synchronize()
out["registration_small_2500x1000_4starts_201iters_ms"] = e0.elapsed_time(e1)
# ---- section 8f rows: scale / ICP candidate search and the fusion tail (reg_xyz.py) ----
from genpc_b200.reg_xyz import iterative_scale_search, knn_mean_distance, remove_close_points, remove_statistical_outlier  # noqa: E402
from genpc_b200.utils.dataUtils import voxel_down_sample  # noqa: E402
tgt_s = voxel_down_sample(tc[0], 0.03)
src_s = voxel_down_sample(tc[0] / torch.tensor([1.1, 1.0, 0.9], device=dev), 0.03)
iterative_scale_search(src_s, tgt_s, [(0.8, 1.2)] * 3, 4, None, 0.5); torch.cuda.synchronize()
t0 = time.perf_counter(); iterative_scale_search(src_s, tgt_s, [(0.8, 1.2)] * 3, 10, None, 0.5); torch.cuda.synchronize()
out["scale_search_1000_candidates_icp30_ms"] = {"ms": (time.perf_counter() - t0) * 1e3, "src_pts": int(src_s.shape[0]),
                                                "tgt_pts": int(tgt_s.shape[0])}
fused = torch.cat([tc[0], tp[0]])[:20000].contiguous()
out["knn20_mean_distance_20000pts_ms"] = ev_time(lambda: knn_mean_distance(fused, 20, True))[0]
out["remove_statistical_outlier_20000pts_ms"] = ev_time(lambda: remove_statistical_outlier(fused, 20, 2.5))[0]
out["remove_close_points_16384x16384_ms"] = ev_time(lambda: remove_close_points(tp[0], tc[0], 1e-4))[0]
def outlier_torch(x, k=20, chunk=4096):   # the library formulation this kernel replaces (torch.cdist + topk)
    md = torch.empty(x.shape[0], device=x.device)
    for c0 in range(0, x.shape[0], chunk):
        md[c0:c0 + chunk] = torch.cdist(x[c0:c0 + chunk], x).topk(k, largest=False).values.mean(1)
    return md
out["knn20_torch_cdist_topk_20000pts_ms"] = ev_time(lambda: outlier_torch(fused))[0]
print(json.dumps(out, indent=1))
