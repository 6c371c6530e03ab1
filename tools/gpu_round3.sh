#!/bin/bash
set +e
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_emd.py -m gpu -x -q --timeout 300 > gpurun_out/pytest_emd.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_emd.log
tail -30 gpurun_out/pytest_emd.log
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 --deselect tests/test_emd.py > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 300 python tools/time_emd.py > gpurun_out/time_emd.json 2> gpurun_out/time_emd.err; cat gpurun_out/time_emd.json; tail -3 gpurun_out/time_emd.err
