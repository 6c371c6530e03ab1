#!/bin/bash
set +e
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_fps.py -m gpu -x -q --timeout 300 > gpurun_out/pytest_fps.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_fps.log
tail -15 gpurun_out/pytest_fps.log
timeout 300 python - <<'PY' 2>&1 | tail -8
import os, sys, torch
sys.path.insert(0, '.')
from genpc_b200.fps import furthest_point_sample
dev = torch.device('cuda:0')
for mode in ('cta', 'cluster'):
    os.environ['GENPC_FPS_MODE'] = mode
    for (B, N, K) in [(1, 16384, 2048), (8, 16384, 2048), (18, 16384, 2048), (1, 32768, 4096), (1, 8192, 1024)]:
        x = torch.rand(B, N, 3, device=dev)
        furthest_point_sample(x, K, 0); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); furthest_point_sample(x, K, 0); e1.record(); torch.cuda.synchronize()
        print(mode, B, N, K, f"{e0.elapsed_time(e1):.3f} ms")
PY
bash tools/sym_variants.sh run 2>&1 | tee gpurun_out/sym_variants2.jsonl | python -c "
import sys, json
for l in sys.stdin:
    try: j=json.loads(l)
    except Exception: continue
    print(j['variant'][:34], j['shape'], j['best_ms'], '%.3e'%j['pairs_per_s'], j['checksum'])
"
