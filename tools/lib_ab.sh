#!/bin/bash
# Same-box A/B of library builds (GENPC_LIB=tools/bin/libgenpc_<tag>.so): pruned forward / fused step timings, two rounds.
mkdir -p gpurun_out
for r in 1 2; do
for v in default "$@"; do
  if [ $v = default ]; then unset GENPC_LIB; else export GENPC_LIB=tools/bin/libgenpc_$v.so; fi
  timeout 120 python tools/time_prune.py 32x2048x16384 32x8192x8192 > gpurun_out/libab_${v}_$r.json 2> gpurun_out/libab_${v}_$r.err
  echo "$v r$r rc=$?"; tail -2 gpurun_out/libab_${v}_$r.err
  python -c "
import json
j=json.load(open('gpurun_out/libab_${v}_$r.json'))
print({k:(v['pruned']['forward']['median_ms'], v['pruned']['loss_step']['median_ms']) for k,v in j.items()})"
done; done
