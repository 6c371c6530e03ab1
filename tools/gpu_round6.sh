#!/bin/bash
set +e
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_reg_xyz.py -m gpu -x -q --timeout 600 > gpurun_out/pytest_regxyz.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_regxyz.log
tail -25 gpurun_out/pytest_regxyz.log
timeout 1200 compute-sanitizer --tool memcheck --print-limit 20 python tools/sanitize_smoke.py > gpurun_out/sanitizer_memcheck.log 2>&1; tail -6 gpurun_out/sanitizer_memcheck.log
timeout 1200 compute-sanitizer --tool racecheck --print-limit 20 python tools/sanitize_smoke.py > gpurun_out/sanitizer_racecheck.log 2>&1; tail -6 gpurun_out/sanitizer_racecheck.log
