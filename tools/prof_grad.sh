mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"nn_sym_epilogue|chamfer_grad_kernel" -s 6 -c 2 -f -o gpurun_out/prof_grad_w python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/ncu_grad.log 2>&1
tail -c 300 gpurun_out/ncu_grad.log
