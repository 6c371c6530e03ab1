#!/bin/bash
set +e
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_fps.py -m gpu -x -q --timeout 120 > gpurun_out/pytest_fps.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_fps.log
tail -15 gpurun_out/pytest_fps.log
timeout 300 python - <<'PY' 2>&1 | tail -14
import os, sys, torch
sys.path.insert(0, '.')
from genpc_b200.fps import furthest_point_sample
dev = torch.device('cuda:0')
for mode in ('cta', 'cluster'):
    os.environ['GENPC_FPS_MODE'] = mode
    for (B, N, K) in [(1, 16384, 2048), (18, 16384, 2048), (1, 32768, 4096), (1, 8192, 1024), (1, 4096, 1024), (1, 1024, 512)]:
        x = torch.rand(B, N, 3, device=dev)
        furthest_point_sample(x, K, 0); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); furthest_point_sample(x, K, 0); e1.record(); torch.cuda.synchronize()
        print(mode, B, N, K, f"{e0.elapsed_time(e1):.3f} ms  {1e3*e0.elapsed_time(e1)/K:.3f} us/pick")
PY
