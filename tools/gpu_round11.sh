#!/bin/bash
set +e
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_fps.py tests/test_chamfer_gpu.py tests/test_emd.py -m gpu -x -q --timeout 300 > gpurun_out/pytest_a.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_a.log
tail -12 gpurun_out/pytest_a.log
timeout 300 python - <<'PY' 2>&1 | tail -8
import os, sys, torch
sys.path.insert(0, '.')
from genpc_b200.fps import furthest_point_sample
dev = torch.device('cuda:0')
for mode in ('cta', 'cluster'):
    os.environ['GENPC_FPS_MODE'] = mode
    for (B, N, K) in [(1, 71372, 10000), (1, 139138, 16384), (1, 45164, 10000)]:
        x = torch.rand(B, N, 3, device=dev)
        furthest_point_sample(x, 64, 0); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); furthest_point_sample(x, K, 0); e1.record(); torch.cuda.synchronize()
        print(mode, B, N, K, f"{e0.elapsed_time(e1):.3f} ms  {1e3*e0.elapsed_time(e1)/K:.3f} us/pick")
PY
