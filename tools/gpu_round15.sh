#!/bin/bash
set +e
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_capi_from_c.py tests/test_depth.py -m gpu -x -q --timeout 300 > gpurun_out/pytest_b.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_b.log
tail -6 gpurun_out/pytest_b.log
timeout 300 python - <<'PY' 2>&1 | tail -6
import sys, time, torch
sys.path.insert(0, '.')
from genpc_b200.DepthPrompting import DepthPrompting
from genpc_b200.synthetic import superquadric
dev = torch.device('cuda:0')
pts = torch.from_numpy(superquadric(0, 71372)).to(dev); rgb = torch.rand(71372, 3, device=dev)
dp = DepthPrompting(dict(view_num=1024, res=256, cam_res=256, downsample_num=10000))
dp.getDepth(pts, rgb); torch.cuda.synchronize()
for _ in range(2):
    t0 = time.perf_counter(); r = dp.getDepth(pts, rgb); torch.cuda.synchronize(); print("getDepth 1024 views res256 71372 pts: %.2f ms, best view %d" % ((time.perf_counter() - t0) * 1e3, r[0]))
t0 = time.perf_counter(); b = dp.viewpoint_select(pts); torch.cuda.synchronize(); print("viewpoint_select: %.2f ms" % ((time.perf_counter() - t0) * 1e3))
PY
