#!/bin/bash
set +e
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -5 gpurun_out/pytest_gpu.log
tail -3 gpurun_out/bench.err
python -c "
import json
j=json.loads(open('gpurun_out/bench.json').read().strip().splitlines()[-1]); print('ours', j['value'], j['ms_per_step'], j['e2e']['value'], j['e2e']['ms_per_step'], j['roofline']['ms'], j['roofline']['frac'], j['roofline_bwd']['ms'], j['registration']['value'])"
./tools/gpu_r2.sh
