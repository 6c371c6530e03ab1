#!/bin/bash
set +e
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 600 python tools/time_misc.py > gpurun_out/time_misc.json 2> gpurun_out/time_misc.err; cat gpurun_out/time_misc.json; tail -3 gpurun_out/time_misc.err
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -2
