#!/bin/bash
set +e
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu.log
tail -8 gpurun_out/pytest_gpu.log
timeout 600 python tools/time_misc.py > gpurun_out/time_misc.json 2> gpurun_out/time_misc.err
python -c "
import json; j=json.load(open('gpurun_out/time_misc.json')); print(json.dumps(j['c3_registration'])); print(j['registration_small_2500x1000_4starts_201iters_ms'])"
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err
timeout 600 python bench.py --impl reference --steps 20 --no-cpu-baseline > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
python -c "
import json
j=json.loads(open('gpurun_out/bench.json').read().strip().splitlines()[-1]); print('ours', j['value'], j['e2e']['value'], j['roofline']['ms'], j['roofline']['frac'], j['roofline_bwd']['ms'], j['registration']['value'])
r=open('gpurun_out/bench_ref.json').read().strip().splitlines(); print('ref lines', len(r)); j=json.loads(r[-1]); print('ref', j['value'], j['e2e']['value'])"
