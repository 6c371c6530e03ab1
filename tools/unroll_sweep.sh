mkdir -p gpurun_out
for r in 1 2; do
for u in default U1 U4 U8; do
  if [ $u = default ]; then unset GENPC_LIB; else export GENPC_LIB=tools/bin/libgenpc_$u.so; fi
  timeout 120 python tools/time_prune.py 32x2048x16384 32x8192x8192 > gpurun_out/usweep_${u}_$r.json 2> gpurun_out/usweep_${u}_$r.err
  echo "$u r$r rc=$?"; tail -2 gpurun_out/usweep_${u}_$r.err
  python -c "
import json,sys
j=json.load(open('gpurun_out/usweep_${u}_$r.json'))
print({k:(v['pruned']['forward']['median_ms'], v['pruned']['loss_step']['median_ms']) for k,v in j.items()})"
done; done
