#!/bin/bash
set +e
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -25 gpurun_out/pytest_gpu.log
timeout 600 python tools/time_misc.py > gpurun_out/time_misc.json 2> gpurun_out/time_misc.err
cat gpurun_out/time_misc.json; tail -5 gpurun_out/time_misc.err
