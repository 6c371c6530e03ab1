#!/bin/bash
# One profiling session on the GPU box: tests, both bench arms, the ncu launch list and one --set full capture per kernel family.
set +e
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
timeout 600 python bench.py --impl reference > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:nn_sym_kernel -s 3 -c 1 -f -o gpurun_out/prof_nn_sym python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/ncu_full.log 2>&1
timeout 300 ncu --set full --clock-control none -k regex:"nn_sym_epilogue|chamfer_grad_kernel" -s 6 -c 2 -f -o gpurun_out/prof_fix_grad python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/ncu_full2.log 2>&1
timeout 300 ncu --set full --clock-control none -k regex:nn_sym_gated -s 2 -c 1 -f -o gpurun_out/prof_nn_gated python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/ncu_full3.log 2>&1
timeout 300 ncu --set full --clock-control none -k regex:"register_sym_scan|register_finish" -s 2 -c 2 -f -o gpurun_out/prof_register python tools/prof_targets.py register > gpurun_out/ncu_reg.log 2>&1
timeout 300 ncu --set full --clock-control none -k regex:emd_auction -s 1 -c 1 -f -o gpurun_out/prof_emd python tools/prof_targets.py emd > gpurun_out/ncu_emd.log 2>&1
timeout 300 ncu --set full --clock-control none -k regex:fps_ -s 1 -c 2 -f -o gpurun_out/prof_fps python tools/prof_targets.py fps > gpurun_out/ncu_fps.log 2>&1
timeout 300 ncu --set full --clock-control none -k regex:knn_mean -s 1 -c 1 -f -o gpurun_out/prof_knn python tools/prof_targets.py knn > gpurun_out/ncu_knn.log 2>&1
timeout 300 ncu --set full --clock-control none -k regex:icp_step -s 1 -c 1 -f -o gpurun_out/prof_icp python tools/prof_targets.py icp > gpurun_out/ncu_icp.log 2>&1
timeout 300 ncu --set full --clock-control none -k regex:"project_kernel|uv_kernel|zsplat|zresolve|unproject" -s 5 -c 5 -f -o gpurun_out/prof_depth python tools/prof_targets.py depth > gpurun_out/ncu_depth.log 2>&1
head -c 2500 gpurun_out/bench.json; echo; for f in ncu_emd ncu_fps ncu_reg ncu_knn; do tail -n 2 gpurun_out/$f.log; done; exit 0
