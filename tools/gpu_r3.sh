#!/bin/bash
set +e
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_emd.py -m gpu -x -q > gpurun_out/pytest_emd.log 2>&1; tail -3 gpurun_out/pytest_emd.log
timeout 300 python tools/emd_iters.py > gpurun_out/emd_iters.json 2> gpurun_out/emd_iters.err; cat gpurun_out/emd_iters.json | tr -d '\n ' | sed 's/},/},\n/g'; tail -3 gpurun_out/emd_iters.err
timeout 300 python tools/time_emd.py > gpurun_out/time_emd.json 2> gpurun_out/time_emd.err; grep -E "ours_ms|reference_ext_ms|B" gpurun_out/time_emd.json
timeout 100 ./tools/fp32_peak 300 4 > gpurun_out/fp32_peak.json
