#!/bin/bash
set +e
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_registration.py tests/test_chamfer_gpu.py tests/test_sharded.py -m gpu -x -q --timeout 600 > gpurun_out/pytest_reg.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_reg.log
tail -15 gpurun_out/pytest_reg.log
timeout 600 python tools/time_misc.py > gpurun_out/time_misc_sym.json 2> gpurun_out/time_misc.err
GENPC_REGISTER_MODE=scan timeout 600 python tools/time_misc.py > gpurun_out/time_misc_scan.json 2>> gpurun_out/time_misc.err
python - <<'PY'
import json
for f in ("sym","scan"):
    j=json.load(open(f"gpurun_out/time_misc_{f}.json")); print(f, json.dumps(j["c3_registration"]), j["registration_small_2500x1000_4starts_201iters_ms"])
PY
tail -3 gpurun_out/time_misc.err
