#!/bin/bash
# Same-box A/B of library builds on the C3 registration metric (and the C2 step) through bench.py.
mkdir -p gpurun_out
for r in 1 2; do
for v in default "$@"; do
  if [ $v = default ]; then unset GENPC_LIB; else export GENPC_LIB=tools/bin/libgenpc_$v.so; fi
  timeout 200 python bench.py --no-extras --no-cpu-baseline --steps 50 > gpurun_out/c3ab_${v}_$r.json 2> gpurun_out/c3ab_${v}_$r.err
  python -c "
import json; j=json.loads(open('gpurun_out/c3ab_${v}_$r.json').read().strip().splitlines()[-1]); print('$v', $r, 'C3', round(j['registration']['value']), j['registration']['ms_per_iter'], 'C2 step', j['ms_per_step'])"
done; done
