#!/bin/bash
# One GPU-box session: microbench, parity tests, goldens, bench, ncu.  Everything lands in gpurun_out/.
set +e
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,power.limit --format=csv > gpurun_out/smi.txt 2>&1
nproc > gpurun_out/nproc.txt
STAGES=${STAGES:-"peak tests golden bench ref ncu"}
for s in $STAGES; do
case $s in
peak)
  timeout 300 ./tools/fp32_peak 300 4 > gpurun_out/fp32_peak.json 2> gpurun_out/fp32_peak.err
  timeout 300 ./tools/fp32_peak 300 2 > gpurun_out/fp32_peak_occ2.json 2>> gpurun_out/fp32_peak.err ;;
tests)
  timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log ;;
golden)
  timeout 600 python tests/golden/make_golden.py > gpurun_out/golden.log 2>&1 ;;
bench)
  timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err ;;
ref)
  timeout 900 python bench.py --impl reference --steps 20 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err ;;
ncu)
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
      python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"${NCU_KERNEL:-nn_sym_kernel}" -s 3 -c 2 -f -o gpurun_out/prof_nn \
      python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1 ;;
smoke)
  timeout 600 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1 ;;
esac
done
tail -3 gpurun_out/pytest_gpu.log 2>/dev/null
cat gpurun_out/bench.json 2>/dev/null | head -c 3000
