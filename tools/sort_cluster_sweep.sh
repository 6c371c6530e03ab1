#!/bin/bash
# Same-box A/B of the sort kernel's cluster layout (GENPC_SORT_CLUSTER): parity tests of the pruned scan, then timings.
mkdir -p gpurun_out
for cs in ${PARITY:-m2 m3 m4 3 8}; do
  GENPC_SORT_CLUSTER=$cs timeout 300 python -m pytest tests/test_chamfer_prune.py tests/test_chamfer_gpu.py -m gpu -x -q 2>&1 | tail -2
done
for cs in ${TIMED:-default 1 2 m2 m3 m4}; do
  if [ $cs = default ]; then unset GENPC_SORT_CLUSTER; else export GENPC_SORT_CLUSTER=$cs; fi
  timeout 120 python tools/time_prune.py 32x2048x16384 32x8192x8192 16x16384x16384 > gpurun_out/sortcs_$cs.json 2> gpurun_out/sortcs_$cs.err
  echo "cs=$cs rc=$?"; tail -2 gpurun_out/sortcs_$cs.err
  python -c "
import json
j=json.load(open('gpurun_out/sortcs_$cs.json'))
print({k:(v['pruned']['forward']['median_ms'], v['pruned']['loss_step']['median_ms']) for k,v in j.items()})"
done
