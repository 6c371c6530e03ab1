#!/bin/bash
set +e
mkdir -p gpurun_out
timeout 300 ./tools/tail_variants.sh > gpurun_out/tail_variants.jsonl 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
timeout 300 python tools/time_misc.py > gpurun_out/time_misc.json 2> gpurun_out/time_misc.err
cat gpurun_out/tail_variants.jsonl | cut -c1-200
tail -5 gpurun_out/pytest_gpu.log
python -c "
import json
j=json.loads(open('gpurun_out/bench.json').read().strip().splitlines()[-1]); print('ours', j['value'], j['ms_per_step'], j['e2e']['value'], j['e2e']['ms_per_step'], j['roofline']['ms'], j['roofline']['frac'], j['roofline_bwd']['ms'], j['registration']['value'])"
