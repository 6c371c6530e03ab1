import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from genpc_b200.DepthPrompting import DepthPrompting
from genpc_b200.synthetic import superquadric
from genpc_b200.fps import furthest_point_sample
dev = torch.device("cuda:0")
pts = torch.from_numpy(superquadric(0, 71372)).to(dev)
rgb = torch.rand(71372, 3, device=dev)
dp = DepthPrompting(dict(view_num=1024, res=256, cam_res=256, downsample_num=10000))
def T(fn, n=3):
    fn(); torch.cuda.synchronize(); ts = []
    for _ in range(n):
        t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
    return round(min(ts), 2), round(max(ts), 2)
for c16 in (None, "0"):
    if c16 is None: os.environ.pop("GENPC_FPS_CLUSTER16", None)
    else: os.environ["GENPC_FPS_CLUSTER16"] = c16
    print("cluster16 =", c16)
    print(" getDepth", T(lambda: dp.getDepth(pts, rgb)))
    print(" fps", T(lambda: furthest_point_sample(pts[None].contiguous(), 10000, 0)))
    idx = furthest_point_sample(pts[None].contiguous(), 10000, 0)[0].long()
    print(" getVisiblePoints(10000 pts, 1024 views)", T(lambda: dp.getVisiblePoints(pts[idx])))
    print(" viewpoint_select", T(lambda: dp.viewpoint_select(pts)))
    cam = dp.cameras[3:4]
    print(" getUvs+vis+raw", T(lambda: (dp.getUvs(cam, pts, True, 0.15), dp.getVisiblePoints(pts, cam))))
