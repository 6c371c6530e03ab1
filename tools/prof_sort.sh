mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"nn_bin_sort|nn_prune_kernel" -s 6 -c 2 -f -o gpurun_out/prof_sort_w python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/ncu_sort.log 2>&1
tail -3 gpurun_out/ncu_sort.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches_r02w.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/ncu_bench.log 2>&1
tail -2 gpurun_out/ncu_bench.log
