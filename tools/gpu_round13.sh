#!/bin/bash
set +e
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_emd.py -m gpu -x -q --timeout 300 > gpurun_out/pytest_emd.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_emd.log
tail -6 gpurun_out/pytest_emd.log
timeout 300 python tools/time_emd.py > gpurun_out/time_emd.json 2> gpurun_out/time_emd.err; python -c "
import json; j=json.load(open('gpurun_out/time_emd.json'))
for k,v in j.items(): print(k, round(v['ours_ms'],3), round(v.get('reference_ext_ms',0),3), round(v.get('speedup',0),2))"
