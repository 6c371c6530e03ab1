#!/bin/bash
set +e
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu.log
tail -8 gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
timeout 600 python bench.py --impl reference --steps 30 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
tail -3 gpurun_out/bench.err gpurun_out/bench_ref.err
python -c "
import json
l=open('gpurun_out/bench.json').read().strip().splitlines(); print('ours lines', len(l)); j=json.loads(l[-1]); print('ours', j['value'], j['ms_per_step'], j['e2e']['value'], j['e2e']['ms_per_step'], j['roofline']['ms'], j['roofline']['frac'], j['roofline_bwd']['ms'], j['registration']['value'], j['cpu_baseline']['value'])
r=open('gpurun_out/bench_ref.json').read().strip().splitlines(); print('ref lines', len(r)); j=json.loads(r[-1]); print('ref', j['value'], j['e2e']['value'], j['note'][:120])"
